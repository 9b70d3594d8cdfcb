// common.cuh — shared declarations of libmdgpu (device structs, launch helpers).
//
// Numerical contract: every float operation that decides a bin or voxel index is carried out with
// the same operands, order and rounding as the reference's scalar/AVX code. The library is compiled with --fmad=false;
// fused multiply-adds appear only where the reference uses an explicit fmadd intrinsic and are written __fmaf_rn here.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <float.h>

#include "../../include/mdgpu.h"

#define MDG_HD __host__ __device__ __forceinline__
#define MDG_D  __device__ __forceinline__

namespace mdg {

// Geometry of one frame's cell grid: the state md_spatial_acc_init leaves in md_spatial_acc_t
// (core/md_spatial_acc.h:68-91, core/md_spatial_acc.c:155-438) plus the neighbour reach of the pair query (:1650-1659).
struct FrameGeom {
    float I[3][3];      // inverse basis, [col][row]
    float A[3][3];      // basis
    float origin[3];
    float G00, G11, G22, H01, H02, H12;
    float inv_cell_ext[3];
    float r2;           // calc_r2(cutoff) :541-544
    int   cdim[3];
    int   ncell[3];     // ceil(cutoff * inv_cell_ext * cdim)
    int   hlo[3];       // home grid (unclamped reference cell coordinates) lower bound
    int   hdim[3];      // home grid dims
    uint32_t flags;
    uint32_t num_cells;
    uint32_t num_home;
    int   valid;        // 0: the reference yields no pairs (degenerate cell / "cutoff too large for cell size")
    int   sym_ok;       // every periodic axis has cdim >= 2*ncell+1: each target cell is reached by exactly one neighbour offset
};

// One cell list (sorted fractional points + offsets) for a batch of frames.
struct CellList {
    float4*   sorted;     // [B][max_points] (sx, sy, sz, idx bits)
    float4*   scratch;    // [B][max_points] unsorted
    uint32_t* cell_of;    // [B][max_points]
    uint32_t* rank;       // [B][max_points]
    uint32_t* cell_cnt;   // [B][cap+1]  counts, then exclusive offsets; followed by [B] out-of-grid flags (oob)
    uint32_t* oob;        // [B] set when a reference point's unclamped cell coordinate lies outside [0,cdim) (home-grid lists only)
    uint32_t  max_points;
    uint32_t  cap;        // cell capacity per frame
};

struct BatchFrames {
    const float* xyz;       // frame i: x at xyz + i*frame_stride, y at +axis_stride, z at +2*axis_stride
    size_t frame_stride;
    size_t axis_stride;
    uint32_t count;
};

// A selection that changes per frame (within(...) evaluated on the device): frame f of the batch uses idx + f * stride, n[f] entries.
// n == nullptr: not dynamic, the caller's static list applies.
struct DynSel { const int32_t* idx; const uint32_t* n; uint32_t stride; };
MDG_HD const int32_t* sel_list(const int32_t* stat, const DynSel& d, int f) { return d.n ? d.idx + (size_t)f * d.stride : stat; }
MDG_HD uint32_t sel_count(uint32_t stat, const DynSel& d, int f) { return d.n ? d.n[f] : stat; }

// launch counter (mdgpu_launch_count)
void note_launch(const char* name, cudaStream_t s);

}  // namespace mdg
