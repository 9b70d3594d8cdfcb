// emul_props.cpp — TEST INFRASTRUCTURE: viamd_b200/csrc/props.cu compiled by g++ (cuda_emul.h); k_distance_pair (one thread per pair) runs on
// the CPU exactly as written, thread by thread.
#include "cuda_emul.h"
#include "props_nolaunch.cu"   // viamd_b200/csrc/props.cu with its <<<>>> launch statements blanked (build_emul.py)
#include <vector>

namespace mdg { void note_launch(const char*, cudaStream_t) {} }   // the launchers are compiled (their launch statements blanked) but never called

extern "C" int emul_distance_pair(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const mdgpu_unitcell_t* cells,
                                  const int32_t* ia, uint32_t na, const int32_t* ib, uint32_t nb, float* out) {
    mdg::BatchFrames fr{}; fr.xyz = frames; fr.frame_stride = frame_stride; fr.axis_stride = axis_stride; fr.count = num_frames;
    const unsigned long long npairs = (unsigned long long)na * nb;
    if (!npairs) return 0;
    blockDim = dim3(256, 1, 1); gridDim = dim3((unsigned)((npairs + 255) / 256), num_frames, 1);   // launch_distance_pair
    for (unsigned by = 0; by < gridDim.y; ++by) for (unsigned bx = 0; bx < gridDim.x; ++bx) for (unsigned t = 0; t < 256; ++t) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = 0; threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
        mdg::k_distance_pair(fr, cells, ia, na, ib, nb, nullptr, nullptr, out, 0);
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
    emul_launch(dim3(num_frames, 1, 1), dim3(256, 1, 1), [&]() { mdg::k_min_distance(fr, cells, ia, na, ib, nb, out, 0, mdg::DynSel{}, mdg::DynSel{}); });   // launch_min_distance
    return 0;
}

// distance / angle / dihedral whose arguments are selections: k_arg_com (8 lanes of a warp replay the AVX2 reference's accumulation, warp
// shuffles) per selection argument, then k_temporal — the launch sequence of plan.cu. args: idx[k] / count[k]; count 1 with direct[k] != 0
// means the atom's own position.
extern "C" int emul_temporal_args(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const mdgpu_unitcell_t* cells,
                                  const float* mass, int op, const int32_t* const idx[4], const uint32_t count[4], const int direct[4], float* out) {
    mdg::BatchFrames fr{}; fr.xyz = frames; fr.frame_stride = frame_stride; fr.axis_stride = axis_stride; fr.count = num_frames;
    std::vector<float> pos((size_t)num_frames * 12, 0.0f);
    mdg::TemporalArgs a{}; a.frames = fr; a.cells = cells; a.op = op; a.out = out; a.frame0 = 0; a.pos = pos.data(); a.com_mask = 0;
    const int nargs = op == MDGPU_OP_DISTANCE ? 2 : (op == MDGPU_OP_ANGLE ? 3 : 4);
    for (int k = 0; k < nargs; ++k) {
        a.atom[k] = idx[k][0];
        if (direct[k]) continue;
        a.com_mask |= 1u << k;
        emul_launch(dim3(num_frames), dim3(32), [&]() { mdg::k_arg_com(fr, cells, idx[k], count[k], mass, pos.data(), k, mdg::DynSel{}); });   // launch_arg_com
    }
    emul_launch(dim3((num_frames + 63) / 64), dim3(64), [&]() { mdg::k_temporal(a, (int)num_frames); });
    return 0;
}

// centres of mass of atom groups (rdf with an array of selections as reference): k_group_com
extern "C" int emul_group_com(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const int32_t* idx, const uint32_t* off,
                              uint32_t n_groups, const float* mass, float* out /* [num_frames][n_groups][3] */) {
    mdg::BatchFrames fr{}; fr.xyz = frames; fr.frame_stride = frame_stride; fr.axis_stride = axis_stride; fr.count = num_frames;
    emul_launch(dim3((n_groups + 127u) / 128u, num_frames), dim3(128), [&]() { mdg::k_group_com(fr, idx, off, n_groups, mass, out); });   // launch_group_com
    return 0;
}
