// synth.cu — synthetic trajectory frames generated directly in HBM (bench.py / tests); arithmetic is synth.h's.
#include "common.cuh"
#include "kernels.h"
#include "synth.h"

namespace mdg {

// frame f, atom i: base + rigid displacement of the atom's molecule, wrapped per axis (mdsynth_frame_coord)
__global__ void k_synth_frames(uint32_t seed, float Lx, float Ly, float Lz, uint32_t num_atoms, const float* __restrict__ base, size_t base_axis_stride,
                               const uint32_t* __restrict__ mol_id /* null: 3-site water, mol = i/3 */, uint32_t frame_beg,
                               float* __restrict__ out, size_t frame_stride, size_t axis_stride) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_atoms) return;
    const uint32_t frame = frame_beg + blockIdx.y;
    const uint32_t mol = mol_id ? mol_id[i] : i / 3u;
    float* o = out + (size_t)blockIdx.y * frame_stride;
    o[i]                   = mdsynth_frame_coord(base[i], seed, frame, mol, 0, Lx);
    o[axis_stride + i]     = mdsynth_frame_coord(base[base_axis_stride + i], seed, frame, mol, 1, Ly);
    o[2 * axis_stride + i] = mdsynth_frame_coord(base[2 * base_axis_stride + i], seed, frame, mol, 2, Lz);
}

void launch_synth_frames(uint32_t seed, float Lx, float Ly, float Lz, uint32_t num_atoms, const float* d_base, size_t base_axis_stride,
                         const uint32_t* d_mol_id, uint32_t frame_beg, uint32_t count, float* d_out, size_t frame_stride, size_t axis_stride, cudaStream_t s) {
    if (!count || !num_atoms) return;
    dim3 grid((num_atoms + 255u) / 256u, count);
    k_synth_frames<<<grid, 256, 0, s>>>(seed, Lx, Ly, Lz, num_atoms, d_base, base_axis_stride, d_mol_id, frame_beg, d_out, frame_stride, axis_stride);
    note_launch("k_synth_frames", s);
}

}  // namespace mdg
