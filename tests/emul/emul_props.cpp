// emul_props.cpp — TEST INFRASTRUCTURE: viamd_b200/csrc/props.cu compiled by g++ (cuda_emul.h); k_distance_pair (one thread per pair) runs on
// the CPU exactly as written, thread by thread.
#include "cuda_emul.h"
#include "props_nolaunch.cu"   // viamd_b200/csrc/props.cu with its <<<>>> launch statements blanked (build_emul.py)

namespace mdg { void note_launch(const char*, cudaStream_t) {} }   // the launchers are compiled (their launch statements blanked) but never called

extern "C" int emul_distance_pair(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const mdgpu_unitcell_t* cells,
                                  const int32_t* ia, uint32_t na, const int32_t* ib, uint32_t nb, float* out) {
    mdg::BatchFrames fr{}; fr.xyz = frames; fr.frame_stride = frame_stride; fr.axis_stride = axis_stride; fr.count = num_frames;
    const unsigned long long npairs = (unsigned long long)na * nb;
    if (!npairs) return 0;
    blockDim = dim3(256, 1, 1); gridDim = dim3((unsigned)((npairs + 255) / 256), num_frames, 1);   // launch_distance_pair
    for (unsigned by = 0; by < gridDim.y; ++by) for (unsigned bx = 0; bx < gridDim.x; ++bx) for (unsigned t = 0; t < 256; ++t) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = 0; threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
        mdg::k_distance_pair(fr, cells, ia, na, ib, nb, out, 0);
    }
    return 0;
}

// k_temporal (distance / angle / dihedral on atoms): a kernel that has passed on the GPU, run here as a check of the emulation itself.
extern "C" int emul_temporal(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const mdgpu_unitcell_t* cells,
                             int op, const int32_t atom[4], float* out) {
    mdg::TemporalArgs a{};
    a.frames.xyz = frames; a.frames.frame_stride = frame_stride; a.frames.axis_stride = axis_stride; a.frames.count = num_frames;
    a.cells = cells; a.op = op; for (int k = 0; k < 4; ++k) a.atom[k] = atom[k]; a.out = out; a.frame0 = 0; a.pos = nullptr; a.com_mask = 0;
    blockDim = dim3(64, 1, 1); gridDim = dim3((num_frames + 63) / 64, 1, 1);   // launch_temporal
    for (unsigned bx = 0; bx < gridDim.x; ++bx) for (unsigned t = 0; t < 64; ++t) {
        blockIdx.x = bx; blockIdx.y = 0; blockIdx.z = 0; threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
        mdg::k_temporal(a, (int)num_frames);
    }
    return 0;
}

// k_com_rows: com(x) rows; pos (may be NULL) holds the [num_frames][4][3] argument positions k_arg_com would have left
extern "C" int emul_com_rows(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, int atom, const float* pos, float* out) {
    mdg::TemporalArgs a{};
    a.frames.xyz = frames; a.frames.frame_stride = frame_stride; a.frames.axis_stride = axis_stride; a.frames.count = num_frames;
    a.atom[0] = atom; a.out = out; a.frame0 = 0; a.pos = pos; a.com_mask = pos ? 1u : 0u;
    blockDim = dim3(64, 1, 1); gridDim = dim3((num_frames + 63) / 64, 1, 1);   // launch_com_rows
    for (unsigned bx = 0; bx < gridDim.x; ++bx) for (unsigned t = 0; t < 64; ++t) {
        blockIdx.x = bx; blockIdx.y = 0; blockIdx.z = 0; threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
        mdg::k_com_rows(a, (int)num_frames);
    }
    return 0;
}

// k_min_distance: a block-cooperative kernel (warp shuffles + shared memory + __syncthreads) that has passed on the GPU, run through
// emul_launch as a check of the cooperative emulation.
extern "C" int emul_min_distance(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const mdgpu_unitcell_t* cells,
                                 const int32_t* ia, uint32_t na, const int32_t* ib, uint32_t nb, float* out) {
    mdg::BatchFrames fr{}; fr.xyz = frames; fr.frame_stride = frame_stride; fr.axis_stride = axis_stride; fr.count = num_frames;
    emul_launch(dim3(num_frames, 1, 1), dim3(256, 1, 1), [&]() { mdg::k_min_distance(fr, cells, ia, na, ib, nb, out, 0); });   // launch_min_distance
    return 0;
}
