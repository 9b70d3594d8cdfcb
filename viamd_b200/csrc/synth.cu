// synth.cu — synthetic trajectory frames generated directly in HBM (bench.py / tests); arithmetic is synth.h's.
#include "common.cuh"
#include "kernels.h"
#include "synth.h"

namespace mdg {

__global__ void k_synth_water(uint32_t seed, float L, uint32_t num_atoms, const float* __restrict__ base, size_t base_axis_stride,
                              uint32_t frame_beg, float* __restrict__ out, size_t frame_stride, size_t axis_stride) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_atoms) return;
    const uint32_t frame = frame_beg + blockIdx.y;
    const uint32_t mol = i / 3u;
    float* o = out + (size_t)blockIdx.y * frame_stride;
#pragma unroll
    for (uint32_t ax = 0; ax < 3; ++ax)
        o[ax * axis_stride + i] = mdsynth_frame_coord(base[ax * base_axis_stride + i], seed, frame, mol, ax, L);
}

void launch_synth_water(uint32_t seed, float L, uint32_t num_atoms, const float* d_base, size_t base_axis_stride, uint32_t frame_beg, uint32_t count,
                        float* d_out, size_t frame_stride, size_t axis_stride, cudaStream_t s) {
    if (!count || !num_atoms) return;
    dim3 grid((num_atoms + 255u) / 256u, count);
    k_synth_water<<<grid, 256, 0, s>>>(seed, L, num_atoms, d_base, base_axis_stride, frame_beg, d_out, frame_stride, axis_stride);
    note_launch("k_synth_water", s);
}

}  // namespace mdg
