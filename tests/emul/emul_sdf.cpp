// emul_sdf.cpp — TEST INFRASTRUCTURE: viamd_b200/csrc/sdf.cu compiled by g++ (cuda_emul.h) so that k_rmsd — one warp per frame, lanes
// only extract, lane 0 does the ordered work — runs on the CPU exactly as written. Lanes 31..1 are run before lane 0, which is one legal
// schedule of the warp (the kernel's only cross-lane dependency is the __syncwarp between extraction and lane 0's serial part).
#include "cuda_emul.h"
#include "sdf_nolaunch.cu"   // viamd_b200/csrc/sdf.cu with its <<<>>> launch statements blanked (build_emul.py)

namespace mdg { void note_launch(const char*, cudaStream_t) {} }   // the launchers are compiled (their launch statements blanked) but never called
#include <vector>

extern "C" int emul_rmsd(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const mdgpu_unitcell_t* cells,
                         const float* init_xyz, size_t init_axis_stride, const float* mass, const int32_t* idx, uint32_t n,
                         const int32_t* pairs /* [n_pairs][2]: child, parent */, uint32_t n_pairs, float* out) {
    std::vector<float4> scratch((size_t)num_frames * 2 * (n ? n : 1));
    mdg::RmsdArgs a{};
    a.frames.xyz = frames; a.frames.frame_stride = frame_stride; a.frames.axis_stride = axis_stride; a.frames.count = num_frames;
    a.cells = cells; a.init_xyz = init_xyz; a.init_axis_stride = init_axis_stride; a.mass = mass; a.idx = idx; a.n = n;
    a.unwrap_pairs = (const int2*)pairs; a.n_unwrap = n_pairs; a.scratch_xyzw = scratch.data(); a.out = out; a.frame0 = 0;
    if (!n) return 0;   // launch_rmsd: nothing is launched for an empty selection
    blockDim = dim3(32, 1, 1); gridDim = dim3(num_frames, 1, 1);
    for (uint32_t f = 0; f < num_frames; ++f) {
        blockIdx.x = f; blockIdx.y = 0; blockIdx.z = 0;
        for (int lane = 31; lane >= 0; --lane) { threadIdx.x = (unsigned)lane; threadIdx.y = 0; threadIdx.z = 0; mdg::k_rmsd(a, (int)num_frames); }
    }
    return 0;
}

// k_plane: same structure as k_rmsd (lanes extract, lane 0 does the ordered part); out is [num_frames][4]
extern "C" int emul_plane(const float* frames, size_t frame_stride, size_t axis_stride, uint32_t num_frames, const mdgpu_unitcell_t* cells,
                          const int32_t* idx, uint32_t n, const int32_t* pairs, uint32_t n_pairs, float* out) {
    std::vector<float4> scratch((size_t)num_frames * (n ? n : 1));
    mdg::RmsdArgs a{};
    a.frames.xyz = frames; a.frames.frame_stride = frame_stride; a.frames.axis_stride = axis_stride; a.frames.count = num_frames;
    a.cells = cells; a.idx = idx; a.n = n; a.unwrap_pairs = (const int2*)pairs; a.n_unwrap = n_pairs; a.scratch_xyzw = scratch.data(); a.out = out; a.frame0 = 0;
    if (!n) return 0;
    blockDim = dim3(32, 1, 1); gridDim = dim3(num_frames, 1, 1);
    for (uint32_t f = 0; f < num_frames; ++f) {
        blockIdx.x = f; blockIdx.y = 0; blockIdx.z = 0;
        for (int lane = 31; lane >= 0; --lane) { threadIdx.x = (unsigned)lane; threadIdx.y = 0; threadIdx.z = 0; mdg::k_plane(a, (int)num_frames); }
    }
    return 0;
}
