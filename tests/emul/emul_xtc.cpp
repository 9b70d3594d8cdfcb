// emul_xtc.cpp — TEST INFRASTRUCTURE: xtc.cu (k_xtc_scan: one warp per frame walking the compressed stream with speculation; k_xtc_decode:
// one thread per group) compiled by g++ and run through emul_launch with the buffers and launch shapes of mdgpu_xtc_decode_frames.
// GPU-validated kernels as a CPU regression net.
#include "cuda_emul.h"
#include "xtc_nolaunch.cu"
#include <vector>

namespace mdg { void note_launch(const char*, cudaStream_t) {} }

extern "C" int emul_xtc_decode(const uint8_t* blob, const uint64_t* frame_offsets, uint32_t count, uint32_t num_atoms, float* out /* [count][3][num_atoms] */) {
    using namespace mdg;
    const uint64_t beg = frame_offsets[0], end = frame_offsets[count];
    std::vector<uint8_t> b((size_t)(end - beg) + 32, 0); memcpy(b.data(), blob + beg, (size_t)(end - beg));
    std::vector<unsigned long long> off(count + 1); for (uint32_t i = 0; i <= count; ++i) off[i] = frame_offsets[i] - beg;
    std::vector<XtcFrameInfo> info(count); std::vector<uint2> rec((size_t)count * num_atoms + 1); std::vector<uint16_t> state((size_t)count * num_atoms + 1, 0);
    int err = 0;
    emul_launch(dim3(count), dim3(32), [&]() { k_xtc_scan(b.data(), off.data(), num_atoms, (int)count, info.data(), rec.data(), state.data(), num_atoms); });
    const uint32_t per_frame = num_atoms <= 9 ? 1u : std::min((num_atoms + 255u) / 256u, 64u);
    emul_launch(dim3(per_frame, count), dim3(256), [&]() { k_xtc_decode(b.data(), off.data(), num_atoms, info.data(), rec.data(), state.data(), num_atoms, out, 3 * (size_t)num_atoms, num_atoms, &err); });
    return err;
}
