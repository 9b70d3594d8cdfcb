// within.cu — dynamic selections, first consumer: count(within(radius, selection)) evaluated per frame.
//
// Replaces _within_expl_flt + within_float_cb (reference md_script_functions.inl:2478-2533) over the system-wide cell list of
// get_spatial_acc (:734-760: every atom of the system, cell extent ceil(radius / 6) * 6), and _count (:2868) on the result:
// the atoms of the system within `radius` of any atom of the selection, the selection's own atoms excluded (:2521-2525); the min:max form
// (_within_expl_frng :2609) queries at max and accepts a pair from d2 >= min * min on.
//
// The cell lists come from cells.cu exactly as for rdf(): targets = all atoms (clamped cells), references = the selection's atoms in
// the home grid. The pair enumeration is the reference's (core/md_spatial_acc.c:1649-1803 / :1498-1647): (2n+1)^3 neighbour offsets of
// the reference point's cell, wrapped once, the reference point shifted by the periodic image, wraps on non-periodic axes skipped;
// d2 = fma(G00, dx*dx, fma(G11, dy*dy, G22*dz*dz)) (+ cross terms, triclinic) compared with calc_r2(radius) (:541-544). A pair within
// the radius sets the target atom's flag; flags are idempotent, so no ordering between warps matters.
#include "common.cuh"
#include "kernels.h"
#include "cellmath.cuh"

namespace mdg {

constexpr int WITHIN_WARPS = 4;

template <bool TRI>
MDG_D float within_d2(float dx, float dy, float dz, const FrameGeom& g) {
    const float dx2 = __fmul_rn(dx, dx), dy2 = __fmul_rn(dy, dy), dz2 = __fmul_rn(dz, dz);
    const float acc = __fmaf_rn(g.G00, dx2, __fmaf_rn(g.G11, dy2, __fmul_rn(g.G22, dz2)));            // distance_squared_ort_256 :524-529
    if (!TRI) return acc;
    const float dxy = __fmul_rn(dx, dy), dxz = __fmul_rn(dx, dz), dyz = __fmul_rn(dy, dz);
    return __fadd_rn(acc, __fmaf_rn(g.H01, dxy, __fmaf_rn(g.H02, dxz, __fmul_rn(g.H12, dyz))));       // distance_squared_tri_256 :503-515
}

// One warp per home cell (grid-stride): for every neighbour cell the lanes take its target points 32 at a time and test them against
// the home cell's reference points until one is within the radius.
template <bool TRI>
__global__ void __launch_bounds__(WITHIN_WARPS * 32) k_within_mark(WithinArgs a) {
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const FrameGeom& g = a.geom[f];
    if (g.valid <= 0) return;
    const float4* __restrict__ trg = a.trg.sorted + (size_t)f * a.trg.max_points;
    const uint32_t* __restrict__ trg_off = a.trg.cell_cnt + (size_t)f * (a.trg.cap + 1);
    const float4* __restrict__ ref = a.ref.sorted + (size_t)f * a.ref.max_points;
    const uint32_t* __restrict__ ref_off = a.ref.cell_cnt + (size_t)f * (a.ref.cap + 1);
    uint8_t* __restrict__ flags = a.flags + (size_t)f * a.num_atoms;
    const int cd0 = g.cdim[0], cd1 = g.cdim[1], cd2 = g.cdim[2], n0 = g.ncell[0], n1 = g.ncell[1], n2 = g.ncell[2];
    const int w0 = 2 * n0 + 1, w1 = 2 * n1 + 1, w2 = 2 * n2 + 1, nn = w0 * w1 * w2;
    const uint32_t hd0 = (uint32_t)g.hdim[0], hd1 = (uint32_t)g.hdim[1];
    const float r2 = g.r2, min_r2 = a.min_r2;
    for (uint32_t h = blockIdx.x * WITHIN_WARPS + warp; h < g.num_home; h += gridDim.x * WITHIN_WARPS) {
        const uint32_t rb = ref_off[h], re = ref_off[h + 1];
        if (rb == re) continue;
        const int cvx = (int)(h % hd0) + g.hlo[0], cvy = (int)((h / hd0) % hd1) + g.hlo[1], cvz = (int)(h / (hd0 * hd1)) + g.hlo[2];   // unclamped cell of the reference points
        for (int n = 0; n < nn; ++n) {
            const int ox = n % w0 - n0, oy = (n / w0) % w1 - n1, oz = n / (w0 * w1) - n2;
            int nx = cvx + ox, ny = cvy + oy, nz = cvz + oz;
            const bool upx = nx > cd0 - 1, lox = nx < 0, upy = ny > cd1 - 1, loy = ny < 0, upz = nz > cd2 - 1, loz = nz < 0;
            if (!TRI) {   // wraps on non-periodic axes are skipped (:1733); triclinic cells are periodic in all axes (:1556-1557)
                if ((upx || lox) && !(g.flags & MDGPU_CELL_PBC_X)) continue;
                if ((upy || loy) && !(g.flags & MDGPU_CELL_PBC_Y)) continue;
                if ((upz || loz) && !(g.flags & MDGPU_CELL_PBC_Z)) continue;
            }
            nx += lox ? cd0 : 0; nx -= upx ? cd0 : 0;
            ny += loy ? cd1 : 0; ny -= upy ? cd1 : 0;
            nz += loz ? cd2 : 0; nz -= upz ? cd2 : 0;
            if (nx < 0 || nx >= cd0 || ny < 0 || ny >= cd1 || nz < 0 || nz >= cd2) continue;   // the reference wraps once only
            const uint32_t cj = ((uint32_t)nz * (uint32_t)cd1 + (uint32_t)ny) * (uint32_t)cd0 + (uint32_t)nx;
            const uint32_t start = trg_off[cj], len = trg_off[cj + 1] - start;
            const float shx = (float)((lox ? 1 : 0) - (upx ? 1 : 0)), shy = (float)((loy ? 1 : 0) - (upy ? 1 : 0)), shz = (float)((loz ? 1 : 0) - (upz ? 1 : 0));
            for (uint32_t j = lane; j < len; j += 32) {
                const float4 t = trg[start + j];
                const uint32_t tj = __float_as_uint(t.w);
                if (flags[tj]) continue;   // already marked (by this or another warp): nothing to add
                bool hit = false;
                for (uint32_t i = rb; i < re && !hit; ++i) {
                    const float4 rf = ref[i];
                    const float fx = __fadd_rn(rf.x, shx), fy = __fadd_rn(rf.y, shy), fz = __fadd_rn(rf.z, shz);   // f + image shift (:1755)
                    const float d2 = within_d2<TRI>(__fsub_rn(fx, t.x), __fsub_rn(fy, t.y), __fsub_rn(fz, t.z), g);
                    hit = d2 <= r2 && d2 >= min_r2;   // within_frng_cb (:2599-2607); min_r2 = 0 for the plain form
                }
                if (hit) flags[tj] = 1;
            }
        }
    }
}

// The selection's own atoms leave the result (md_bitfield_andnot_inplace :2524), then _count (:2868): one CTA per frame.
__global__ void __launch_bounds__(256) k_within_count(WithinArgs a) {
    const int f = blockIdx.x;
    uint8_t* __restrict__ flags = a.flags + (size_t)f * a.num_atoms;
    for (uint32_t k = threadIdx.x; k < a.n_sel; k += blockDim.x) flags[a.sel[k]] = 0;
    __syncthreads();
    uint32_t c = 0;
    for (uint32_t i = threadIdx.x; i < a.num_atoms; i += blockDim.x) c += (flags[i] != 0 && (!a.and_mask || a.and_mask[i] != 0)) ? 1u : 0u;
    __shared__ uint32_t total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    if (c) atomicAdd(&total, c);
    __syncthreads();
    if (threadIdx.x == 0) a.out[a.frame0 + f] = (float)total;
}

// ---------------------------------------------------------------------------------------------------------------
// The dynamic selection as the REFERENCE set of an rdf(): rdf(within(radius, selection), targets, cutoff). The marked atoms (selection
// removed) become a per-frame index list; its reference cell list is built like cells.cu builds a static one, but with the frame's own
// list and count (coordinate_extract on a single bitfield: the atoms' positions, ascending index; compute_rdf :5281-5290).
// ---------------------------------------------------------------------------------------------------------------
// flags -> ascending index list, one CTA per frame: 256 atoms per step, ballot prefix inside the warps, running offset across steps
__global__ void __launch_bounds__(256) k_within_compact(WithinArgs a, int32_t* __restrict__ dyn_idx, uint32_t* __restrict__ dyn_n) {
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* __restrict__ flags = a.flags + (size_t)f * a.num_atoms;
    for (uint32_t k = threadIdx.x; k < a.n_sel; k += blockDim.x) flags[a.sel[k]] = 0;   // md_bitfield_andnot_inplace (:2524)
    __shared__ uint32_t s_warp[8]; __shared__ uint32_t s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    int32_t* out = dyn_idx + (size_t)f * a.num_atoms;
    for (uint32_t i0 = 0; i0 < a.num_atoms; i0 += 256u) {
        const uint32_t i = i0 + threadIdx.x;
        const bool on = i < a.num_atoms && flags[i] != 0 && (!a.and_mask || a.and_mask[i] != 0);
        const uint32_t m = __ballot_sync(0xffffffffu, on);
        if (lane == 0) s_warp[warp] = (uint32_t)__popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { const uint32_t c = s_warp[w]; before += (w < warp) ? c : 0u; total += c; }
        const uint32_t base = s_base;
        if (on) out[base + before + (uint32_t)__popc(m & ((1u << lane) - 1u))] = (int32_t)i;
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) dyn_n[f] = s_base;
}

// zero the flags, mark: a few CTAs per frame, enough to fill the SMs across the batch
static void launch_mark(const WithinArgs& a, int B, bool tri, int sm_count, cudaStream_t s) {
    cudaMemsetAsync(a.flags, 0, (size_t)B * a.num_atoms, s);
    if (!a.n_sel) return;
    const dim3 grid((unsigned)max(1, (4 * sm_count) / max(B, 1) + 1), (unsigned)B);
    if (tri) k_within_mark<true><<<grid, WITHIN_WARPS * 32, 0, s>>>(a); else k_within_mark<false><<<grid, WITHIN_WARPS * 32, 0, s>>>(a);
    note_launch("k_within_mark", s);
}

// within() marks -> per-frame reference list (dyn_idx [B][num_atoms], dyn_n [B])
void launch_within_list(const WithinArgs& a, int B, bool tri, int sm_count, int32_t* d_dyn_idx, uint32_t* d_dyn_n, cudaStream_t s) {
    if (B <= 0) return;
    launch_mark(a, B, tri, sm_count, s);
    k_within_compact<<<B, 256, 0, s>>>(a, d_dyn_idx, d_dyn_n);
    note_launch("k_within_compact", s);
}

void launch_within_count(const WithinArgs& a, int B, bool tri, int sm_count, cudaStream_t s) {
    if (B <= 0) return;
    launch_mark(a, B, tri, sm_count, s);
    k_within_count<<<B, 256, 0, s>>>(a);
    note_launch("k_within_count", s);
}

}  // namespace mdg
