// props.cu — K5 (density_x/_y/_z streaming histogram) and K6 (distance / angle / dihedral, batched over frames).
//
// K5 replaces _internal_density (reference md_script_functions.inl:4825-4947): deperiodise about the centre of the
// INITIAL frame's cell, bin = clamp((int)(fc*1024)), bins[bin] += mass. The reference sums masses in float in atom order;
// here masses are accumulated exactly as 64-bit fixed point (unit 2^-24 Da, exact for every float mass >= 1 Da), which is
// order independent and therefore deterministic; the float result agrees with the reference to its own rounding drift.
// K6 replaces _distance / _angle / _dihedral for single-atom arguments (:3851-3890, :4099-4114, :4171-4196).
#include "common.cuh"
#include "kernels.h"

namespace mdg {

MDG_D float deperiodize1p(float x, float r, float ext) {   // vec4_deperiodize_ortho core/md_vec_math.h:1242-1253
    if (ext == 0.0f) return x;
    const float inv = __fdiv_rn(1.0f, ext);
    const float dx = __fmul_rn(__fsub_rn(x, r), inv);
    const float dxp = __fsub_rn(dx, rintf(dx));
    return __fadd_rn(r, __fmul_rn(dxp, ext));
}

// Per-CTA histogram in two 32-bit limbs: a 64-bit shared atomicAdd is a CAS loop (64 cyc/warp, more under contention), a 32-bit one is native.
// The low limb takes the low word of the 2^-24 fixed-point mass; the carry out of each individual add (old + lo wraps) goes to the
// high limb together with the high word, which for masses < 256 u happens for mass/256 of the atoms only.
__global__ void __launch_bounds__(256) k_density(DensityArgs a) {
    const int f = blockIdx.y;
    __shared__ uint32_t hist_lo[MDGPU_DIST_BINS], hist_hi[MDGPU_DIST_BINS];
    for (int b = threadIdx.x; b < MDGPU_DIST_BINS; b += blockDim.x) { hist_lo[b] = 0u; hist_hi[b] = 0u; }
    __syncthreads();
    const float* src = a.frames.xyz + (size_t)f * a.frames.frame_stride + (size_t)a.axis * a.frames.axis_stride;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const int at = a.idx[i];
        const float v = deperiodize1p(src[at], a.rc, a.re);
        const float fc = __fmul_rn(__fsub_rn(v, a.min_point), a.inv_ext);
        const int b = max(0, min(__float2int_rz(__fmul_rn(fc, (float)MDGPU_DIST_BINS)), MDGPU_DIST_BINS - 1));
        const unsigned long long m = __float2ull_rn(__fmul_rn(a.mass[at], 16777216.0f));
        const uint32_t lo = (uint32_t)m;
        const uint32_t old = atomicAdd(&hist_lo[b], lo);
        const uint32_t hi = (uint32_t)(m >> 32) + ((old + lo) < old ? 1u : 0u);
        if (hi) atomicAdd(&hist_hi[b], hi);
    }
    __syncthreads();
    unsigned long long* out = a.frame_bins + (size_t)f * MDGPU_DIST_BINS;
    for (int b = threadIdx.x; b < MDGPU_DIST_BINS; b += blockDim.x) {
        const unsigned long long v = ((unsigned long long)hist_hi[b] << 32) + hist_lo[b];
        if (v) atomicAdd(&out[b], v);
    }
}

__global__ void k_density_finalize(DensityArgs a) {
    const int f = blockIdx.x, t = threadIdx.x;
    const unsigned long long v = a.frame_bins[(size_t)f * MDGPU_DIST_BINS + t];
    const uint32_t gf = a.frame0 + f;
    if (v) atomicAdd(&a.acc[t], v);
    if (a.keep) a.keep[(size_t)gf * MDGPU_DIST_BINS + t] = v;
    unsigned long long mn = v, mx = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long m1 = __shfl_xor_sync(0xffffffffu, mn, o), m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = m1 < mn ? m1 : mn; mx = m2 > mx ? m2 : mx;
    }
    __shared__ unsigned long long s_mn[32], s_mx[32];
    if ((t & 31) == 0) { s_mn[t >> 5] = mn; s_mx[t >> 5] = mx; }
    __syncthreads();
    if (t < 32) {
        mn = s_mn[t]; mx = s_mx[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long m1 = __shfl_xor_sync(0xffffffffu, mn, o), m2 = __shfl_xor_sync(0xffffffffu, mx, o);
            mn = m1 < mn ? m1 : mn; mx = m2 > mx ? m2 : mx;
        }
        if (t == 0) { a.frame_min[gf] = mn; a.frame_max[gf] = mx; }
    }
}

void launch_density(const DensityArgs& a, int B, cudaStream_t s) {
    cudaMemsetAsync(a.frame_bins, 0, sizeof(unsigned long long) * (size_t)B * MDGPU_DIST_BINS, s);
    if (a.n) {
        const uint32_t blocks = min((a.n + 256u * 8u - 1u) / (256u * 8u), 64u);   // ~8 atoms per thread, <=64 CTAs per frame
        dim3 grid(blocks ? blocks : 1u, B);
        k_density<<<grid, 256, 0, s>>>(a);
        note_launch("k_density", s);
    }
    k_density_finalize<<<B, MDGPU_DIST_BINS, 0, s>>>(a);
    note_launch("k_density_finalize", s);
}

// ---------------------------------------------------------------------------------------------------------------
// Centres of mass of atom groups, as coordinate_extract() produces them for an array of bitfields
// (md_script_functions.inl:1496-1507 -> extract_com :857-874): NO periodic treatment, one sequential float pass in ascending atom
// order, sum += (x*w, y*w, z*w, 1*w), then xyz / w (w == 0 -> 1). One thread per (group, frame): the order of the float additions
// is the result.
__global__ void k_group_com(BatchFrames fr, const int32_t* __restrict__ idx, const uint32_t* __restrict__ off, uint32_t n_groups,
                            const float* __restrict__ mass, float* __restrict__ out /* [B][n_groups][3] */) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (g >= n_groups) return;
    const float* x = fr.xyz + (size_t)f * fr.frame_stride; const float* y = x + fr.axis_stride; const float* z = y + fr.axis_stride;
    float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
    for (uint32_t k = off[g]; k < off[g + 1]; ++k) {
        const int a = idx[k]; const float w = mass[a];
        sx = __fadd_rn(sx, __fmul_rn(x[a], w)); sy = __fadd_rn(sy, __fmul_rn(y[a], w)); sz = __fadd_rn(sz, __fmul_rn(z[a], w));
        sw = __fadd_rn(sw, __fmul_rn(1.0f, w));
    }
    if (sw == 0.0f) sw = 1.0f;
    float* o = out + ((size_t)f * n_groups + g) * 3;
    o[0] = __fdiv_rn(sx, sw); o[1] = __fdiv_rn(sy, sw); o[2] = __fdiv_rn(sz, sw);
}

void launch_group_com(const BatchFrames& fr, const int32_t* d_idx, const uint32_t* d_off, uint32_t n_groups, const float* d_mass, float* d_out, cudaStream_t s) {
    if (!n_groups || !fr.count) return;
    dim3 grid((n_groups + 127u) / 128u, fr.count);
    k_group_com<<<grid, 128, 0, s>>>(fr, d_idx, d_off, n_groups, d_mass, d_out);
    note_launch("k_group_com", s);
}

// ---------------------------------------------------------------------------------------------------------------
MDG_D void normalize3(float v[3]) {   // vec3_normalize core/md_vec_math.h:505-514 (threshold compared in double)
    const float len = __fsqrt_rn(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if ((double)len > 1.0e-5) { v[0] = v[0] / len; v[1] = v[1] / len; v[2] = v[2] / len; } else { v[0] = v[1] = v[2] = 0.0f; }
}

__global__ void k_temporal(TemporalArgs a, int B) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= B) return;
    const float* x = a.frames.xyz + (size_t)f * a.frames.frame_stride; const float* y = x + a.frames.axis_stride; const float* z = y + a.frames.axis_stride;
    const mdgpu_unitcell_t uc = a.cells[f];
    const float ext[3] = { (float)uc.x, (float)uc.y, (float)uc.z };
    float out = 0.0f;
    if (a.op == MDGPU_OP_DISTANCE) {
        const int ia = a.atom[0], ib = a.atom[1];
        const float pa[3] = { x[ia], y[ia], z[ia] }; float pb[3] = { x[ib], y[ib], z[ib] };
        if (uc.flags & MDGPU_CELL_ORTHO) for (int k = 0; k < 3; ++k) pb[k] = deperiodize1p(pb[k], pa[k], ext[k]);   // md_util_deperiodize_vec4 md_util.c:8971
        const float d[3] = { pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2] };
        out = __fsqrt_rn(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    } else if (a.op == MDGPU_OP_ANGLE) {
        const int ia = a.atom[0], ib = a.atom[1], ic = a.atom[2];
        float v0[3] = { x[ia] - x[ib], y[ia] - y[ib], z[ia] - z[ib] }, v1[3] = { x[ic] - x[ib], y[ic] - y[ib], z[ic] - z[ib] };
        normalize3(v0); normalize3(v1);
        out = acosf(v0[0] * v1[0] + v0[1] * v1[1] + v0[2] * v1[2]);
    } else if (a.op == MDGPU_OP_DIHEDRAL) {
        float dx[3][3];
        for (int k = 0; k < 3; ++k) { const int p = a.atom[k], q = a.atom[k + 1]; dx[k][0] = x[q] - x[p]; dx[k][1] = y[q] - y[p]; dx[k][2] = z[q] - z[p]; }
        if (uc.flags & MDGPU_CELL_ORTHO) {   // min_image_ortho md_util.c:8424-8436
            for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) {
                const float half = ext[i] * 0.5f;
                if (ext[i] > 0.0f) {
                    int guard = 0;
                    while (dx[k][i] > half && guard++ < 64) dx[k][i] -= ext[i];
                    while (dx[k][i] <= -half && guard++ < 128) dx[k][i] += ext[i];
                }
            }
        }
        const float* d1 = dx[0]; const float* d2 = dx[1]; const float* d3 = dx[2];   // vec3_dihedral_angle core/md_vec_math.h:558-567
        const float v1[3] = { d1[1] * d2[2] - d1[2] * d2[1], d1[2] * d2[0] - d1[0] * d2[2], d1[0] * d2[1] - d1[1] * d2[0] };
        const float v2[3] = { d2[1] * d3[2] - d2[2] * d3[1], d2[2] * d3[0] - d2[0] * d3[2], d2[0] * d3[1] - d2[1] * d3[0] };
        const float w[3] = { v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0] };
        const float wl = __fsqrt_rn(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        const float sdot = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
        float angle = atan2f(wl, sdot);
        const float dot = d1[0] * v2[0] + d1[1] * v2[1] + d1[2] * v2[2];
        if (dot < 0.0f) angle = -angle;
        out = angle;
    }
    a.out[a.frame0 + f] = out;
}

// fold of an integer accumulator into the float mean the property data exposes: (float)((double)count / (double)n), IEEE on the device
__global__ void k_mean_u32(const uint32_t* __restrict__ in, float* __restrict__ out, size_t count, unsigned long long n) {
    const double dn = (double)n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        out[i] = n ? (float)((double)in[i] / dn) : 0.0f;
}
void launch_mean_u32(const uint32_t* d_in, float* d_out, size_t count, unsigned long long n, cudaStream_t s) {
    k_mean_u32<<<148 * 4, 256, 0, s>>>(d_in, d_out, count, n);
    note_launch("k_mean_u32", s);
}

void launch_temporal(const TemporalArgs& a, int B, cudaStream_t s) {
    k_temporal<<<(B + 63) / 64, 64, 0, s>>>(a, B);
    note_launch("k_temporal", s);
}

}  // namespace mdg
