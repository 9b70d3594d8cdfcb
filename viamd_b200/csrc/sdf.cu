// sdf.cu — K3 (local reference-frame fit per structure) and K4 (AABB gather + transform + voxel scatter) for sdf().
//
// Replaces _sdf / sdf_cb (reference md_script_functions.inl:5643-5856), md_util_unwrap_vec4 (md_util.c:8738-8819,8938),
// com_vec4 (md_util.c:8048), mat3_covariance_matrix_vec4 / mat3_cross_covariance_matrix_vec4 / mat3_eigen /
// mat3_extract_rotation (core/md_vec_math.c:22-42,101-156,227-300), svd (ext/svd3/svd3.c) and
// md_spatial_acc_for_each_point_in_aabb (core/md_spatial_acc.c:1805-2007).
//
// All float expressions keep the reference's association; double accumulators stay double; the library is built with
// --fmad=false so nothing is contracted. The only fused operations are the reference's explicit fmadd intrinsics.
#include "common.cuh"
#include "kernels.h"
#include <string.h>
#include <stdlib.h>

namespace mdg {

constexpr int SDF_REC = 32;   // floats per structure record: M[16] com[3] pad lo[3] hi[3] cmin[3] cmax[3]

// ------------------------------------------------------------------------------------------------- 3x3 SVD (McAdams)
struct M3 { float e[3][3]; };   // e[col][row] as in the reference's mat3_t; the svd routine itself is row-major A[r][c]
struct M4 { float e[4][4]; };

MDG_D float inv_sqrtf_(float v) { return __fdiv_rn(1.0f, __fsqrt_rn(v)); }
MDG_D void cond_swap(bool c, float& X, float& Y) { const float Z = X; X = c ? Y : X; Y = c ? Z : Y; }
MDG_D void cond_neg_swap(bool c, float& X, float& Y) { const float Z = -X; X = c ? Y : X; Y = c ? Z : Y; }

MDG_D void approx_givens(float a11, float a12, float a22, float& ch, float& sh) {
    ch = 2.0f * (a11 - a22);
    sh = a12;
    // _gamma is a double literal in svd3.c: the comparison is evaluated in double
    const bool b = 5.828427124746190097 * (double)sh * (double)sh < (double)(ch * ch);
    const float w = inv_sqrtf_(ch * ch + sh * sh);
    ch = b ? w * ch : (float)0.923879532511286756;
    sh = b ? w * sh : (float)0.382683432365089771;
}

MDG_D void jacobi_conj(const int x, const int y, const int z, float S[3][3], float q[4]) {
    float ch, sh; approx_givens(S[0][0], S[1][0], S[1][1], ch, sh);
    const float scale = ch * ch + sh * sh;
    const float a = (ch * ch - sh * sh) / scale;
    const float b = (2.0f * sh * ch) / scale;
    const float s00 = S[0][0], s10 = S[1][0], s11 = S[1][1], s20 = S[2][0], s21 = S[2][1], s22 = S[2][2];
    const float n00 = a * (a * s00 + b * s10) + b * (a * s10 + b * s11);
    const float n10 = a * (-b * s00 + a * s10) + b * (-b * s10 + a * s11);
    const float n11 = -b * (-b * s00 + a * s10) + a * (-b * s10 + a * s11);
    const float n20 = a * s20 + b * s21;
    const float n21 = -b * s20 + a * s21;
    const float n22 = s22;
    const float tmp0 = q[0] * sh, tmp1 = q[1] * sh, tmp2 = q[2] * sh;
    const float tmp[3] = { tmp0, tmp1, tmp2 };
    sh *= q[3];
    q[0] *= ch; q[1] *= ch; q[2] *= ch; q[3] *= ch;
    q[z] += sh; q[3] -= tmp[z]; q[x] += tmp[y]; q[y] -= tmp[x];
    S[0][0] = n11; S[1][0] = n21; S[1][1] = n22; S[2][0] = n10; S[2][1] = n20; S[2][2] = n00;
}

MDG_D float dist2_(float a, float b, float c) { return a * a + b * b + c * c; }

MDG_D void qr_givens(float a1, float a2, float& ch, float& sh) {
    const float epsilon = (float)1e-6;
    const float rho = __fsqrt_rn(a1 * a1 + a2 * a2);
    sh = rho > epsilon ? a2 : 0.0f;
    ch = fabsf(a1) + fmaxf(rho, epsilon);
    const bool b = a1 < 0.0f;
    cond_swap(b, sh, ch);
    const float w = inv_sqrtf_(ch * ch + sh * sh);
    ch *= w; sh *= w;
}

__device__ __noinline__ void svd3(const float A[3][3], float U[3][3], float S[3][3], float V[3][3]) {
    float ATA[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) ATA[i][j] = A[0][i] * A[0][j] + A[1][i] * A[1][j] + A[2][i] * A[2][j];
    float q[4] = { 0.f, 0.f, 0.f, 1.f };
    for (int it = 0; it < 4; ++it) { jacobi_conj(0, 1, 2, ATA, q); jacobi_conj(1, 2, 0, ATA, q); jacobi_conj(2, 0, 1, ATA, q); }
    {
        const float x = q[0], y = q[1], z = q[2], w = q[3];
        const float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
        V[0][0] = 1.0f - 2.0f * (qyy + qzz); V[0][1] = 2.0f * (qxy - qwz);        V[0][2] = 2.0f * (qxz + qwy);
        V[1][0] = 2.0f * (qxy + qwz);        V[1][1] = 1.0f - 2.0f * (qxx + qzz); V[1][2] = 2.0f * (qyz - qwx);
        V[2][0] = 2.0f * (qxz - qwy);        V[2][1] = 2.0f * (qyz + qwx);        V[2][2] = 1.0f - 2.0f * (qxx + qyy);
    }
    float B[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) B[i][j] = A[i][0] * V[0][j] + A[i][1] * V[1][j] + A[i][2] * V[2][j];
    {
        float rho1 = dist2_(B[0][0], B[1][0], B[2][0]), rho2 = dist2_(B[0][1], B[1][1], B[2][1]), rho3 = dist2_(B[0][2], B[1][2], B[2][2]);
        bool c = rho1 < rho2;
#pragma unroll
        for (int r = 0; r < 3; ++r) { cond_neg_swap(c, B[r][0], B[r][1]); cond_neg_swap(c, V[r][0], V[r][1]); }
        cond_swap(c, rho1, rho2);
        c = rho1 < rho3;
#pragma unroll
        for (int r = 0; r < 3; ++r) { cond_neg_swap(c, B[r][0], B[r][2]); cond_neg_swap(c, V[r][0], V[r][2]); }
        cond_swap(c, rho1, rho3);
        c = rho2 < rho3;
#pragma unroll
        for (int r = 0; r < 3; ++r) { cond_neg_swap(c, B[r][1], B[r][2]); cond_neg_swap(c, V[r][1], V[r][2]); }
    }
    {
        float (*Q)[3] = U; float (*R)[3] = S;
        float ch1, sh1, ch2, sh2, ch3, sh3, a, b;
        qr_givens(B[0][0], B[1][0], ch1, sh1);
        a = 1.0f - 2.0f * sh1 * sh1; b = 2.0f * ch1 * sh1;
        R[0][0] = a * B[0][0] + b * B[1][0];  R[0][1] = a * B[0][1] + b * B[1][1];  R[0][2] = a * B[0][2] + b * B[1][2];
        R[1][0] = -b * B[0][0] + a * B[1][0]; R[1][1] = -b * B[0][1] + a * B[1][1]; R[1][2] = -b * B[0][2] + a * B[1][2];
        R[2][0] = B[2][0]; R[2][1] = B[2][1]; R[2][2] = B[2][2];
        qr_givens(R[0][0], R[2][0], ch2, sh2);
        a = 1.0f - 2.0f * sh2 * sh2; b = 2.0f * ch2 * sh2;
        B[0][0] = a * R[0][0] + b * R[2][0];  B[0][1] = a * R[0][1] + b * R[2][1];  B[0][2] = a * R[0][2] + b * R[2][2];
        B[1][0] = R[1][0]; B[1][1] = R[1][1]; B[1][2] = R[1][2];
        B[2][0] = -b * R[0][0] + a * R[2][0]; B[2][1] = -b * R[0][1] + a * R[2][1]; B[2][2] = -b * R[0][2] + a * R[2][2];
        qr_givens(B[1][1], B[2][1], ch3, sh3);
        a = 1.0f - 2.0f * sh3 * sh3; b = 2.0f * ch3 * sh3;
        R[0][0] = B[0][0]; R[0][1] = B[0][1]; R[0][2] = B[0][2];
        R[1][0] = a * B[1][0] + b * B[2][0];  R[1][1] = a * B[1][1] + b * B[2][1];  R[1][2] = a * B[1][2] + b * B[2][2];
        R[2][0] = -b * B[1][0] + a * B[2][0]; R[2][1] = -b * B[1][1] + a * B[2][1]; R[2][2] = -b * B[1][2] + a * B[2][2];
        const float sh12 = sh1 * sh1, sh22 = sh2 * sh2, sh32 = sh3 * sh3;
        Q[0][0] = (-1.0f + 2.0f * sh12) * (-1.0f + 2.0f * sh22);
        Q[0][1] = 4.0f * ch2 * ch3 * (-1.0f + 2.0f * sh12) * sh2 * sh3 + 2.0f * ch1 * sh1 * (-1.0f + 2.0f * sh32);
        Q[0][2] = 4.0f * ch1 * ch3 * sh1 * sh3 - 2.0f * ch2 * (-1.0f + 2.0f * sh12) * sh2 * (-1.0f + 2.0f * sh32);
        Q[1][0] = 2.0f * ch1 * sh1 * (1.0f - 2.0f * sh22);
        Q[1][1] = -8.0f * ch1 * ch2 * ch3 * sh1 * sh2 * sh3 + (-1.0f + 2.0f * sh12) * (-1.0f + 2.0f * sh32);
        Q[1][2] = -2.0f * ch3 * sh3 + 4.0f * sh1 * (ch3 * sh1 * sh3 + ch1 * ch2 * sh2 * (-1.0f + 2.0f * sh32));
        Q[2][0] = 2.0f * ch2 * sh2;
        Q[2][1] = 2.0f * ch3 * (1.0f - 2.0f * sh22) * sh3;
        Q[2][2] = (-1.0f + 2.0f * sh22) * (-1.0f + 2.0f * sh32);
    }
}

// ------------------------------------------------------------------------------------------------- small matrix helpers
MDG_D M3 m3_transpose(const M3& M) { M3 T; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T.e[i][j] = M.e[j][i]; return T; }
MDG_D M3 m3_mul(const M3& A, const M3& B) {   // core/md_vec_math.h:1631
    M3 C;
    for (int col = 0; col < 3; ++col) for (int row = 0; row < 3; ++row)
        C.e[col][row] = A.e[0][row] * B.e[col][0] + A.e[1][row] * B.e[col][1] + A.e[2][row] * B.e[col][2];
    return C;
}
MDG_D float m3_det(const M3& M) {             // :1687
    return M.e[0][0] * (M.e[1][1] * M.e[2][2] - M.e[2][1] * M.e[1][2])
         - M.e[1][0] * (M.e[0][1] * M.e[2][2] - M.e[2][1] * M.e[0][2])
         + M.e[2][0] * (M.e[0][1] * M.e[1][2] - M.e[1][1] * M.e[0][2]);
}
MDG_D M4 m4_from_m3(const M3& M) { M4 R; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) R.e[i][j] = (i < 3 && j < 3) ? M.e[i][j] : 0.0f; R.e[3][3] = 1.0f; return R; }
MDG_D M4 m4_mul(const M4& A, const M4& B) {   // linear_combine_4 (:1512)
    M4 C;
    for (int j = 0; j < 4; ++j) for (int r = 0; r < 4; ++r) {
        float v = B.e[j][0] * A.e[0][r];
        v = v + B.e[j][1] * A.e[1][r];
        v = v + B.e[j][2] * A.e[2][r];
        v = v + B.e[j][3] * A.e[3][r];
        C.e[j][r] = v;
    }
    return C;
}
struct Svd { M3 U, V; float s[3]; };
MDG_D Svd m3_svd(const M3& M) {               // core/md_vec_math.c:7-20
    M3 Mt = m3_transpose(M), U, S, V;
    svd3(Mt.e, U.e, S.e, V.e);
    Svd r; r.U = m3_transpose(U); r.V = m3_transpose(V); r.s[0] = S.e[0][0]; r.s[1] = S.e[1][1]; r.s[2] = S.e[2][2];
    return r;
}

// vec4_deperiodize_ortho (core/md_vec_math.h:1242-1253): round = nearest-even
MDG_D float deperiodize1(float x, float r, float ext) {
    if (ext == 0.0f) return x;
    const float inv = __fdiv_rn(1.0f, ext);
    const float dx = __fmul_rn(__fsub_rn(x, r), inv);
    const float dxp = __fsub_rn(dx, rintf(dx));
    return __fadd_rn(r, __fmul_rn(dxp, ext));
}

// minimum_image_triclinic md_util.c:1677-1718: the 27 images, squared length compared in double, first minimum in loop order wins.
// The sums are mixed precision as written there: float products (box * int) and float sums where both operands are float.
MDG_D void min_image_triclinic(float dx[3], const float box[3][3]) {
    double m0 = 0.0, m1 = 0.0, m2 = 0.0, dsq_min = (double)3.402823466e+38f;
    for (int ix = -1; ix < 2; ++ix) {
        const double rx = (double)__fadd_rn(dx[0], __fmul_rn(box[0][0], (float)ix));
        for (int iy = -1; iy < 2; ++iy) {
            const double ry0 = __dadd_rn(rx, (double)__fmul_rn(box[1][0], (float)iy));
            const double ry1 = (double)__fadd_rn(dx[1], __fmul_rn(box[1][1], (float)iy));
            for (int iz = -1; iz < 2; ++iz) {
                const double rz0 = __dadd_rn(ry0, (double)__fmul_rn(box[2][0], (float)iz)), rz1 = __dadd_rn(ry1, (double)__fmul_rn(box[2][1], (float)iz));
                const double rz2 = (double)__fadd_rn(dx[2], __fmul_rn(box[2][2], (float)iz));
                const double dsq = __dadd_rn(__dadd_rn(__dmul_rn(rz0, rz0), __dmul_rn(rz1, rz1)), __dmul_rn(rz2, rz2));
                if (dsq < dsq_min) { dsq_min = dsq; m0 = rz0; m1 = rz1; m2 = rz2; }
            }
        }
    }
    dx[0] = (float)m0; dx[1] = (float)m1; dx[2] = (float)m2;
}

// extract + unwrap + centre of mass of one structure into scratch (xyz, mass)
MDG_D void load_unwrap_com(float4* p, const float* x, const float* y, const float* z, const float* mass, const int32_t* sidx, uint32_t n,
                           const int2* pairs, uint32_t n_pairs, const mdgpu_unitcell_t& uc, float com[3]) {
    for (uint32_t k = 0; k < n; ++k) { const int a = sidx[k]; p[k] = make_float4(x[a], y[a], z[a], mass[a]); }
    if (uc.flags & MDGPU_CELL_ORTHO) {
        const float ext[3] = { (float)uc.x, (float)uc.y, (float)uc.z };
        for (uint32_t k = 0; k < n_pairs; ++k) {
            const int2 pr = pairs[k];
            const float4 ref = p[pr.y]; float4 v = p[pr.x];
            v.x = deperiodize1(v.x, ref.x, ext[0]); v.y = deperiodize1(v.y, ref.y, ext[1]); v.z = deperiodize1(v.z, ref.z, ext[2]);
            p[pr.x] = v;
        }
    } else if (uc.flags & MDGPU_CELL_TRICLINIC) {   // unwrap_atom_triclinic_vec4 -> deperiodize_triclinic md_util.c:1754-1766
        const float box[3][3] = { { (float)uc.x, 0.f, 0.f }, { (float)uc.xy, (float)uc.y, 0.f }, { (float)uc.xz, (float)uc.yz, (float)uc.z } };
        for (uint32_t k = 0; k < n_pairs; ++k) {
            const int2 pr = pairs[k];
            const float4 ref = p[pr.y]; float4 v = p[pr.x];
            float d[3] = { __fsub_rn(v.x, ref.x), __fsub_rn(v.y, ref.y), __fsub_rn(v.z, ref.z) };
            min_image_triclinic(d, box);
            v.x = __fadd_rn(ref.x, d[0]); v.y = __fadd_rn(ref.y, d[1]); v.z = __fadd_rn(ref.z, d[2]);
            p[pr.x] = v;
        }
    }
    float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;   // com_vec4 md_util.c:8048-8061
    for (uint32_t k = 0; k < n; ++k) { const float4 v = p[k]; ax = ax + v.x * v.w; ay = ay + v.y * v.w; az = az + v.z * v.w; aw = aw + v.w * 1.0f; }
    com[0] = ax / aw; com[1] = ay / aw; com[2] = az / aw;
}

// Reference structure (structure 0 of the INITIAL frame, unwrapped with the CURRENT frame's cell :5762-5782): PCA frame and V*A
__global__ void k_sdf_ref0(SdfArgs a, int B) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= B) return;
    const uint32_t n = a.struct_size;
    float4* p = a.scratch_xyzw + ((size_t)f * (a.n_struct + 1) + a.n_struct) * n;
    float com0[3];
    load_unwrap_com(p, a.init_xyz, a.init_xyz + a.init_axis_stride, a.init_xyz + 2 * a.init_axis_stride, a.mass, a.struct_idx, n,
                    a.unwrap_pairs, a.n_unwrap, a.cells[f], com0);
    double C[3][3] = { { 0 } }; double ws = 0.0;   // mat3_covariance_matrix_vec4
    for (uint32_t k = 0; k < n; ++k) {
        const float4 v = p[k];
        const float x = v.x - com0[0], y = v.y - com0[1], z = v.z - com0[2], w = v.w;
        C[0][0] += w * x * x; C[0][1] += w * x * y; C[0][2] += w * x * z;
        C[1][0] += w * y * x; C[1][1] += w * y * y; C[1][2] += w * y * z;
        C[2][0] += w * z * x; C[2][1] += w * z * y; C[2][2] += w * z * z;
        ws += w;
    }
    M3 cov; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov.e[i][j] = (float)(C[i][j] / ws);
    // mat3_eigen (core/md_vec_math.c:22-42)
    const Svd s = m3_svd(cov);
    const float mx = fmaxf(s.s[0], fmaxf(s.s[1], s.s[2]));
    const float ev[3] = { s.s[0] / mx, s.s[1] / mx, s.s[2] / mx };
    int l0 = 0, l1 = 1, l2 = 2, t;
    if (ev[l0] < ev[l1]) { t = l0; l0 = l1; l1 = t; }
    if (ev[l1] < ev[l2]) { t = l1; l1 = l2; l2 = t; }
    if (ev[l0] < ev[l1]) { t = l0; l0 = l1; l1 = t; }
    const int l[3] = { l0, l1, l2 };
    M3 eig; for (int k = 0; k < 3; ++k) for (int r = 0; r < 3; ++r) eig.e[k][r] = s.U.e[l[k]][r];
    const M4 A = m4_from_m3(m3_transpose(eig));
    // compute_volume_matrix (:5643-5655)
    const float voxel_ext = (2.0f * a.cutoff) / (float)MDGPU_VOL_DIM;
    const float sc = 1.0f / voxel_ext;
    M4 S; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) S.e[i][j] = 0.0f; S.e[0][0] = sc; S.e[1][1] = sc; S.e[2][2] = sc; S.e[3][3] = 1.0f;
    const float tt = (float)(MDGPU_VOL_DIM / 2);
    M4 T; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T.e[i][j] = (i == j) ? 1.0f : 0.0f; T.e[3][0] = tt; T.e[3][1] = tt; T.e[3][2] = tt;
    const M4 VA = m4_mul(m4_mul(T, S), A);
    float* o = a.ref0 + (size_t)f * 20;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o[i * 4 + j] = VA.e[i][j];
    o[16] = com0[0]; o[17] = com0[1]; o[18] = com0[2]; o[19] = 0.0f;
}

// K3: one thread per (structure, frame): M = V*A*R*T(-com) (:5788-5799)
__global__ void k_sdf_fit(SdfArgs a, int B) {
    const int f = blockIdx.y;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_struct) return;
    const uint32_t n = a.struct_size;
    const float* x = a.frames.xyz + (size_t)f * a.frames.frame_stride;
    float4* p1 = a.scratch_xyzw + ((size_t)f * (a.n_struct + 1) + s) * n;
    const float4* p0 = a.scratch_xyzw + ((size_t)f * (a.n_struct + 1) + a.n_struct) * n;
    const float* r0 = a.ref0 + (size_t)f * 20;
    const float com0[3] = { r0[16], r0[17], r0[18] };
    float com1[3];
    load_unwrap_com(p1, x, x + a.frames.axis_stride, x + 2 * a.frames.axis_stride, a.mass, a.struct_idx + (size_t)s * n, n,
                    a.unwrap_pairs, a.n_unwrap, a.cells[f], com1);
    double C[3][3] = { { 0 } }; double ws = 0.0;   // mat3_cross_covariance_matrix_vec4 (core/md_vec_math.c:256-289)
    for (uint32_t k = 0; k < n; ++k) {
        const float4 u = p0[k], v = p1[k];
        const float px = u.x - com0[0], py = u.y - com0[1], pz = u.z - com0[2], pw = u.w - 0.0f;
        const float qx = v.x - com1[0], qy = v.y - com1[1], qz = v.z - com1[2], qw = v.w - 0.0f;
        const float w = (pw + qw) * 0.5f;
        C[0][0] += w * px * qx; C[0][1] += w * px * qy; C[0][2] += w * px * qz;
        C[1][0] += w * py * qx; C[1][1] += w * py * qy; C[1][2] += w * py * qz;
        C[2][0] += w * pz * qx; C[2][1] += w * pz * qy; C[2][2] += w * pz * qz;
        ws += w;
    }
    M3 cc; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cc.e[i][j] = (float)(C[i][j] / ws);
    // mat3_extract_rotation (core/md_vec_math.c:292-300)
    const Svd sv = m3_svd(cc);
    const M3 Ut = m3_transpose(sv.U);
    const float d = m3_det(m3_mul(sv.V, Ut));
    M3 D; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) D.e[i][j] = 0.0f; D.e[0][0] = 1.0f; D.e[1][1] = 1.0f; D.e[2][2] = (float)((d > 0.0f) - (d < 0.0f));
    const M3 R = m3_mul(m3_mul(sv.V, D), Ut);
    M4 Tm; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Tm.e[i][j] = (i == j) ? 1.0f : 0.0f; Tm.e[3][0] = -com1[0]; Tm.e[3][1] = -com1[1]; Tm.e[3][2] = -com1[2];
    const M4 RT = m4_mul(m4_from_m3(R), Tm);
    M4 VA; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) VA.e[i][j] = r0[i * 4 + j];
    const M4 M = m4_mul(VA, RT);
    float* o = a.matrices + ((size_t)f * a.n_struct + s) * SDF_REC;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o[i * 4 + j] = M.e[i][j];
    o[16] = com1[0]; o[17] = com1[1]; o[18] = com1[2]; o[19] = 0.0f;

    // cell_range_from_aabb_center_radius + the fractional bounds of for_each_point_in_aabb_ortho (core/md_spatial_acc.c:1805-1923),
    // double precision with the float matrix entries widened, evaluated once per structure
    const FrameGeom& g = a.geom[f];
    const double cen[3] = { (double)com1[0], (double)com1[1], (double)com1[2] };
    const double rad = (double)a.cutoff;
    const int pbc[3] = { (g.flags & MDGPU_CELL_PBC_X) != 0, (g.flags & MDGPU_CELL_PBC_Y) != 0, (g.flags & MDGPU_CELL_PBC_Z) != 0 };
    double sc[3], ccen[3];
    {
        const double px = cen[0] - g.origin[0], py = cen[1] - g.origin[1], pz = cen[2] - g.origin[2];
        sc[0] = g.I[0][0] * px + g.I[1][0] * py + g.I[2][0] * pz;
        sc[1] = g.I[0][1] * px + g.I[1][1] * py + g.I[2][1] * pz;
        sc[2] = g.I[0][2] * px + g.I[1][2] * py + g.I[2][2] * pz;
    }
    for (int k = 0; k < 3; ++k) if (pbc[k]) sc[k] = sc[k] - floor(sc[k]);
    ccen[0] = g.A[0][0] * sc[0] + g.A[1][0] * sc[1] + g.A[2][0] * sc[2] + g.origin[0];
    ccen[1] = g.A[0][1] * sc[0] + g.A[1][1] * sc[1] + g.A[2][1] * sc[2] + g.origin[1];
    ccen[2] = g.A[0][2] * sc[0] + g.A[1][2] * sc[1] + g.A[2][2] * sc[2] + g.origin[2];
    double fmin_[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, fmax_[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    for (int corner = 0; corner < 8; ++corner) {
        const double pz = ccen[2] + ((corner & 4) ? rad : -rad), py = ccen[1] + ((corner & 2) ? rad : -rad), px = ccen[0] + ((corner & 1) ? rad : -rad);
        const double qx = px - g.origin[0], qy = py - g.origin[1], qz = pz - g.origin[2];
        double sv[3];
        sv[0] = g.I[0][0] * qx + g.I[1][0] * qy + g.I[2][0] * qz;
        sv[1] = g.I[0][1] * qx + g.I[1][1] * qy + g.I[2][1] * qz;
        sv[2] = g.I[0][2] * qx + g.I[1][2] * qy + g.I[2][2] * qz;
        for (int k = 0; k < 3; ++k) { fmin_[k] = fmin(fmin_[k], sv[k]); fmax_[k] = fmax(fmax_[k], sv[k]); }
    }
    int* oi = (int*)(o + 26);
    for (int k = 0; k < 3; ++k) {
        const int cdk = g.cdim[k];
        double frad = 0.5 * (fmax_[k] - fmin_[k]);
        int lo = (int)floor(fmin_[k] * (double)cdk), hi = (int)ceil(fmax_[k] * (double)cdk);
        if (hi <= lo) hi = lo + 1;
        if (!pbc[k]) { lo = max(0, min(lo, cdk)); hi = max(0, min(hi, cdk)); if (hi <= lo) hi = min(lo + 1, cdk); }
        frad = fmin(frad, 0.5);
        if (g.flags & MDGPU_CELL_TRICLINIC) { o[20 + k] = (float)(ccen[k] - rad); o[23 + k] = (float)(ccen[k] + rad); }   // cartesian bounds (:2041-2047)
        else { o[20 + k] = (float)(sc[k] - frad); o[23 + k] = (float)(sc[k] + frad); }                                   // fractional bounds (:1913-1923)
        oi[k] = lo; oi[3 + k] = hi;
    }
}

MDG_D int wrap_coord(int v, int N) { v += (v < 0) ? N : 0; v -= (v >= N) ? N : 0; return v; }
MDG_D int isign(int v) { return (v > 0) - (v < 0); }

// K4: one warp per (structure, frame): target points of the cells overlapping AABB(com, cutoff) -> voxel increments.
//  * lanes enumerate the cells of the range once (wrap, image code, offsets) into a small per-warp segment table;
//  * each HALF-warp then walks one cell at a time, 16 points per step (a cell of the bench workload holds ~45 targets: three steps at 94 %
//    lane use, where a full warp per cell would run two steps at 70 % and a flattened index space pays a boundary search per step);
//  * the box test passes ~60 % of the candidates (the box is 20 A wide, the 2-3 cells per axis it overlaps 22-33 A), so the transform
//    and the voxel increment run in place under the hit predicate — compacting hits first costs more than the idle lanes it would fill
//    (profiles/r2_01_fullset_ncu.txt: the ring version spent a third of its instructions on the compaction and another third on set-up).
constexpr int SDF_WARPS = 8;
constexpr int SDF_MAXSEG = 128;
constexpr int SDF_EXCL_CACHE = 64;

struct SdfXform { float M[4][3]; float A00, A11, A22, O0, O1, O2, A10, A20, A21; };

// fractional (image-shifted) point -> cartesian -> structure frame -> voxel (:5664-5697)
template <bool TRI>
MDG_D void sdf_splat(float vx, float vy, float vz, const SdfXform& X, uint32_t* __restrict__ vol) {
    // ortho: batch_fract_to_cart_ort_256, one fused multiply-add per axis (md_spatial_acc.c:583-592).
    // triclinic: REFERENCE QUIRK — for_each_point_in_aabb_triclinic buffers the fractional image-shifted coordinates (:2122-2130) and its
    // *_CART_TRI callback macros (:715-737) skip the conversion, so sdf_cb transforms fractional numbers. Reproduced for parity.
    const float px = TRI ? vx : __fmaf_rn(vx, X.A00, X.O0), py = TRI ? vy : __fmaf_rn(vy, X.A11, X.O1), pz = TRI ? vz : __fmaf_rn(vz, X.A22, X.O2);
    float c[3];   // mat4_mul_vec4(M, (x,y,z,1)) = ((x*M0 + y*M1) + z*M2) + 1*M3; 1*M3 is M3 exactly
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float v = __fmul_rn(px, X.M[0][r]);
        v = __fadd_rn(v, __fmul_rn(py, X.M[1][r]));
        v = __fadd_rn(v, __fmul_rn(pz, X.M[2][r]));
        v = __fadd_rn(v, X.M[3][r]);
        c[r] = v;
    }
    const uint32_t ix = (uint32_t)max(0, min(__float2int_rz(c[0]), MDGPU_VOL_DIM - 1));
    const uint32_t iy = (uint32_t)max(0, min(__float2int_rz(c[1]), MDGPU_VOL_DIM - 1));
    const uint32_t iz = (uint32_t)max(0, min(__float2int_rz(c[2]), MDGPU_VOL_DIM - 1));
    atomicAdd(&vol[(iz * MDGPU_VOL_DIM + iy) * MDGPU_VOL_DIM + ix], 1u);
}

template <bool TRI, int MINB>
__global__ void __launch_bounds__(SDF_WARPS * 32, MINB) k_sdf_scatter(SdfArgs a, int B) {
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, hl = lane & 15, half = lane >> 4;
    const uint32_t s = blockIdx.x * SDF_WARPS + warp;
    __shared__ uint2 s_seg[SDF_WARPS][SDF_MAXSEG];        // x: first point of the cell, y: point count | image code << 26
    __shared__ int32_t s_excl[SDF_WARPS][SDF_EXCL_CACHE];
    if (s >= a.n_struct) return;
    const FrameGeom& g = a.geom[f];
    if (g.valid == -1) return;
    const float* rec = a.matrices + ((size_t)f * a.n_struct + s) * SDF_REC;
    SdfXform X;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) X.M[i][j] = rec[i * 4 + j];
    X.A00 = g.A[0][0]; X.A11 = g.A[1][1]; X.A22 = g.A[2][2]; X.O0 = g.origin[0]; X.O1 = g.origin[1]; X.O2 = g.origin[2];
    X.A10 = g.A[1][0]; X.A20 = g.A[2][0]; X.A21 = g.A[2][1];
    const float lo3[3] = { rec[20], rec[21], rec[22] }, hi3[3] = { rec[23], rec[24], rec[25] };
    const int* ri = (const int*)(rec + 26);
    const int cmin[3] = { ri[0], ri[1], ri[2] }, cmax[3] = { ri[3], ri[4], ri[5] };
    const int pbc[3] = { (g.flags & MDGPU_CELL_PBC_X) != 0, (g.flags & MDGPU_CELL_PBC_Y) != 0, (g.flags & MDGPU_CELL_PBC_Z) != 0 };
    const int cd[3] = { g.cdim[0], g.cdim[1], g.cdim[2] };
    const float4* __restrict__ pts = a.trg.sorted + (size_t)f * a.trg.max_points;
    const uint32_t* __restrict__ off = a.trg.cell_cnt + (size_t)f * (a.trg.cap + 1);
    const int32_t* sidx = a.struct_idx + (size_t)s * a.struct_size;
    // exclusion mask = the structure's own atoms (:5674). Ascending index lists: a contiguous run (the usual case: a residue)
    // is tested with one compare; otherwise the list (cached in shared memory when it fits) is scanned.
    const uint32_t ex_lo = (uint32_t)sidx[0], ex_n = a.struct_size;
    const bool ex_contig = ((uint32_t)sidx[a.struct_size - 1] - ex_lo + 1u) == ex_n;
    if (!ex_contig) { for (uint32_t k = lane; k < min(ex_n, (uint32_t)SDF_EXCL_CACHE); k += 32) s_excl[warp][k] = sidx[k]; }
    const int ex = cmax[0] - cmin[0], ey = cmax[1] - cmin[1], ez = cmax[2] - cmin[2];
    const int ncells = ex * ey * ez;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t local = 0;                    // voxel increments of this lane
    for (int c0 = 0; c0 < ncells; c0 += SDF_MAXSEG) {   // (:1925-1943) cells of the range, SDF_MAXSEG at a time
        const int nc = min(SDF_MAXSEG, ncells - c0);
        int nseg = 0;
        __syncwarp();
        for (int n0 = 0; n0 < nc; n0 += 32) {
            const int n = n0 + lane;
            uint32_t len = 0, start = 0, code = 0x15;
            if (n < nc) {
                const int q = c0 + n;
                const int qx = q % ex, qyz = q / ex;
                const int icx = cmin[0] + qx, icy = cmin[1] + qyz % ey, icz = cmin[2] + qyz / ey;
                const int cx = pbc[0] ? wrap_coord(icx, cd[0]) : icx, cy = pbc[1] ? wrap_coord(icy, cd[1]) : icy, cz = pbc[2] ? wrap_coord(icz, cd[2]) : icz;
                if (!(cx < 0 || cx >= cd[0] || cy < 0 || cy >= cd[1] || cz < 0 || cz >= cd[2])) {
                    const uint32_t ci = ((uint32_t)cz * (uint32_t)cd[1] + (uint32_t)cy) * (uint32_t)cd[0] + (uint32_t)cx;
                    start = off[ci]; len = off[ci + 1] - start;
                    code = (uint32_t)(isign(icx - cx) + 1) | ((uint32_t)(isign(icy - cy) + 1) << 2) | ((uint32_t)(isign(icz - cz) + 1) << 4);
                }
            }
            const uint32_t have = __ballot_sync(0xffffffffu, len != 0u);   // keep non-empty cells only
            if (len) s_seg[warp][nseg + __popc(have & lt)] = make_uint2(start, len | (code << 26));
            nseg += __popc(have);
        }
        __syncwarp();
        for (int k0 = 0; k0 < nseg; k0 += 2) {   // one cell per half-warp
            const int k = k0 + half;
            const uint2 sg = (k < nseg) ? s_seg[warp][k] : make_uint2(0u, 0u);
            const uint32_t len = sg.y & 0x3ffffffu, code = sg.y >> 26;
            const uint32_t steps = max(__shfl_sync(0xffffffffu, len, 0), __shfl_sync(0xffffffffu, len, 16));
            const float shx = (float)((int)(code & 3u) - 1), shy = (float)((int)((code >> 2) & 3u) - 1), shz = (float)((int)((code >> 4) & 3u) - 1);
            auto visit = [&](const float4& t) {   // image shift, box test, exclusion, splat of one candidate
                float vx = t.x, vy = t.y, vz = t.z;
                if (code != 0x15u) {   // periodic image of the cell: + (-1|0|+1), rounded (:1962-1964); +0 is the identity
                    vx = __fadd_rn(vx, shx); vy = __fadd_rn(vy, shy); vz = __fadd_rn(vz, shz);
                }
                bool hit;
                if (TRI) {   // box test on the cartesian image (fract_to_cart_tri_256 md_spatial_acc.c:594-603), all axes periodic (:2009)
                    const float cx_ = __fmaf_rn(vx, X.A00, __fmaf_rn(vy, X.A10, __fmaf_rn(vz, X.A20, X.O0))), cy_ = __fmaf_rn(vy, X.A11, __fmaf_rn(vz, X.A21, X.O1)), cz_ = __fmaf_rn(vz, X.A22, X.O2);
                    hit = cx_ >= lo3[0] && cy_ >= lo3[1] && cz_ >= lo3[2] && cx_ <= hi3[0] && cy_ <= hi3[1] && cz_ <= hi3[2];
                } else hit = vx >= lo3[0] && vy >= lo3[1] && vz >= lo3[2] && vx <= hi3[0] && vy <= hi3[1] && vz <= hi3[2];
                if (hit) {
                    const uint32_t idx = __float_as_uint(t.w);
                    if (ex_contig) hit = (idx - ex_lo) >= ex_n;
                    else {
                        bool excluded = false;
                        const uint32_t nc_ = min(ex_n, (uint32_t)SDF_EXCL_CACHE);
                        for (uint32_t q = 0; q < nc_; ++q) excluded |= ((uint32_t)s_excl[warp][q] == idx);
                        for (uint32_t q = nc_; q < ex_n; ++q) excluded |= ((uint32_t)sidx[q] == idx);
                        hit = !excluded;
                    }
                    if (hit) { sdf_splat<TRI>(vx, vy, vz, X, a.vol); ++local; }
                }
            };
            for (uint32_t j0 = (uint32_t)hl; j0 < steps; j0 += 32u) {   // two candidates per lane and round: both loads in flight before the tests
                const uint32_t j1 = j0 + 16u;
                const bool k0_ = j0 < len, k1_ = j1 < len;
                float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
                if (k0_) t0 = pts[sg.x + j0];
                if (k1_) t1 = pts[sg.x + j1];
                if (k0_) visit(t0);
                if (k1_) visit(t1);
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0 && local) atomicAdd(&a.frame_total[a.frame0 + f], (unsigned long long)local);
}

// The round-1 form: flattened candidate index space (boundary search per step) and a per-warp ring that compacts hits before the splat.
// Kept selectable (MDGPU_SDF=ring) as the measured alternative: 0.750 vs 0.743 ms per 148 frames for the half-warp form above.
constexpr int SDF_RING = 64;
template <bool TRI>
__global__ void __launch_bounds__(SDF_WARPS * 32) k_sdf_scatter_ring(SdfArgs a, int B) {
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t s = blockIdx.x * SDF_WARPS + warp;
    __shared__ uint32_t s_pre[SDF_WARPS][SDF_MAXSEG + 1];
    __shared__ uint2 s_seg[SDF_WARPS][SDF_MAXSEG];        // x: start - pre (first point of the segment minus its flattened offset), y: image code
    __shared__ float s_ring[SDF_WARPS][3][SDF_RING];
    __shared__ int32_t s_excl[SDF_WARPS][SDF_EXCL_CACHE];
    if (s >= a.n_struct) return;
    const FrameGeom& g = a.geom[f];
    if (g.valid == -1) return;
    const float* rec = a.matrices + ((size_t)f * a.n_struct + s) * SDF_REC;
    SdfXform X;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) X.M[i][j] = rec[i * 4 + j];
    X.A00 = g.A[0][0]; X.A11 = g.A[1][1]; X.A22 = g.A[2][2]; X.O0 = g.origin[0]; X.O1 = g.origin[1]; X.O2 = g.origin[2];
    X.A10 = g.A[1][0]; X.A20 = g.A[2][0]; X.A21 = g.A[2][1];
    const float lo3[3] = { rec[20], rec[21], rec[22] }, hi3[3] = { rec[23], rec[24], rec[25] };
    const int* ri = (const int*)(rec + 26);
    const int cmin[3] = { ri[0], ri[1], ri[2] }, cmax[3] = { ri[3], ri[4], ri[5] };
    const int pbc[3] = { (g.flags & MDGPU_CELL_PBC_X) != 0, (g.flags & MDGPU_CELL_PBC_Y) != 0, (g.flags & MDGPU_CELL_PBC_Z) != 0 };
    const int cd[3] = { g.cdim[0], g.cdim[1], g.cdim[2] };
    const float4* __restrict__ pts = a.trg.sorted + (size_t)f * a.trg.max_points;
    const uint32_t* __restrict__ off = a.trg.cell_cnt + (size_t)f * (a.trg.cap + 1);
    const int32_t* sidx = a.struct_idx + (size_t)s * a.struct_size;
    // exclusion mask = the structure's own atoms (:5674). Ascending index lists: a contiguous run (the usual case: a residue)
    // is tested with one compare; otherwise the list (cached in shared memory when it fits) is scanned.
    const uint32_t ex_lo = (uint32_t)sidx[0], ex_n = a.struct_size;
    const bool ex_contig = ((uint32_t)sidx[a.struct_size - 1] - ex_lo + 1u) == ex_n;
    if (!ex_contig) { for (uint32_t k = lane; k < min(ex_n, (uint32_t)SDF_EXCL_CACHE); k += 32) s_excl[warp][k] = sidx[k]; }
    const int ex = cmax[0] - cmin[0], ey = cmax[1] - cmin[1], ez = cmax[2] - cmin[2];
    const int ncells = ex * ey * ez;
    const uint32_t lt = (1u << lane) - 1u;
    float* rx = s_ring[warp][0]; float* ry = s_ring[warp][1]; float* rz = s_ring[warp][2];
    uint32_t cnt = 0;                      // candidates waiting in the ring (warp-uniform, < 32 between steps)
    unsigned long long local = 0;
    for (int c0 = 0; c0 < ncells; c0 += SDF_MAXSEG) {   // (:1925-1943) cells of the range, SDF_MAXSEG at a time
        const int nc = min(SDF_MAXSEG, ncells - c0);
        uint32_t base = 0; int nseg = 0;
        __syncwarp();
        for (int n0 = 0; n0 < nc; n0 += 32) {
            const int n = n0 + lane;
            uint32_t len = 0, start = 0, code = 0x15;
            if (n < nc) {
                const int q = c0 + n;
                const int icx = cmin[0] + q % ex, icy = cmin[1] + (q / ex) % ey, icz = cmin[2] + q / (ex * ey);
                const int cx = pbc[0] ? wrap_coord(icx, cd[0]) : icx, cy = pbc[1] ? wrap_coord(icy, cd[1]) : icy, cz = pbc[2] ? wrap_coord(icz, cd[2]) : icz;
                if (!(cx < 0 || cx >= cd[0] || cy < 0 || cy >= cd[1] || cz < 0 || cz >= cd[2])) {
                    const uint32_t ci = ((uint32_t)cz * (uint32_t)cd[1] + (uint32_t)cy) * (uint32_t)cd[0] + (uint32_t)cx;
                    start = off[ci]; len = off[ci + 1] - start;
                    code = (uint32_t)(isign(icx - cx) + 1) | ((uint32_t)(isign(icy - cy) + 1) << 2) | ((uint32_t)(isign(icz - cz) + 1) << 4);
                }
            }
            const uint32_t have = __ballot_sync(0xffffffffu, len != 0u);   // keep non-empty cells only
            uint32_t incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (len) { const int slot = nseg + __popc(have & lt); const uint32_t pre = base + incl - len; s_pre[warp][slot] = pre; s_seg[warp][slot] = make_uint2(start - pre, code); }
            base += __shfl_sync(0xffffffffu, incl, 31);
            nseg += __popc(have);
        }
        const uint32_t total = base;
        if (lane == 0) s_pre[warp][nseg] = total;
        __syncwarp();
        int kbase = 0;                     // segment that contains candidate j0 (warp-uniform)
        for (uint32_t j0 = 0; j0 < total; j0 += 32) {
            // segment boundaries inside (j0, j0+32]: bit (b - j0 - 1). Segments are non-empty, so they are among the next 32 table entries.
            const int kb = kbase + 1 + lane;
            const uint32_t bnd = (kb <= nseg) ? s_pre[warp][kb] : 0xffffffffu;
            const uint32_t rel = bnd - j0 - 1u;
            const uint32_t bm = __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
            const uint32_t j = j0 + lane;
            bool hit = false; float vx = 0.f, vy = 0.f, vz = 0.f;
            if (j < total) {
                const uint2 sg = s_seg[warp][kbase + __popc(bm & lt)];
                const float4 t = pts[sg.x + j];
                vx = t.x; vy = t.y; vz = t.z;
                if (sg.y != 0x15u) {   // periodic image of the cell: + (-1|0|+1), rounded (:1962-1964); +0 is the identity
                    vx = __fadd_rn(vx, (float)((int)(sg.y & 3u) - 1)); vy = __fadd_rn(vy, (float)((int)((sg.y >> 2) & 3u) - 1)); vz = __fadd_rn(vz, (float)((int)((sg.y >> 4) & 3u) - 1));
                }
                if (TRI) {   // box test on the cartesian image (fract_to_cart_tri_256 md_spatial_acc.c:594-603), all axes periodic (:2009)
                    const float cx_ = __fmaf_rn(vx, X.A00, __fmaf_rn(vy, X.A10, __fmaf_rn(vz, X.A20, X.O0))), cy_ = __fmaf_rn(vy, X.A11, __fmaf_rn(vz, X.A21, X.O1)), cz_ = __fmaf_rn(vz, X.A22, X.O2);
                    hit = cx_ >= lo3[0] && cy_ >= lo3[1] && cz_ >= lo3[2] && cx_ <= hi3[0] && cy_ <= hi3[1] && cz_ <= hi3[2];
                } else hit = vx >= lo3[0] && vy >= lo3[1] && vz >= lo3[2] && vx <= hi3[0] && vy <= hi3[1] && vz <= hi3[2];
                if (hit) {
                    const uint32_t idx = __float_as_uint(t.w);
                    if (ex_contig) hit = (idx - ex_lo) >= ex_n;
                    else {
                        bool excluded = false;
                        const uint32_t nc_ = min(ex_n, (uint32_t)SDF_EXCL_CACHE);
                        for (uint32_t q = 0; q < nc_; ++q) excluded |= ((uint32_t)s_excl[warp][q] == idx);
                        for (uint32_t q = nc_; q < ex_n; ++q) excluded |= ((uint32_t)sidx[q] == idx);
                        hit = !excluded;
                    }
                }
            }
            kbase += __popc(bm);
            const uint32_t hm = __ballot_sync(0xffffffffu, hit);
            if (hit) { const uint32_t p = cnt + __popc(hm & lt); rx[p] = vx; ry[p] = vy; rz[p] = vz; }
            cnt += __popc(hm);
            if (cnt >= 32u) {
                __syncwarp();
                sdf_splat<TRI>(rx[lane], ry[lane], rz[lane], X, a.vol);
                const uint32_t rem = cnt - 32u;
                float mx = 0.f, my = 0.f, mz = 0.f;
                if (lane < rem) { mx = rx[32 + lane]; my = ry[32 + lane]; mz = rz[32 + lane]; }
                __syncwarp();
                if (lane < rem) { rx[lane] = mx; ry[lane] = my; rz[lane] = mz; }
                cnt = rem; local += 32;
                __syncwarp();
            }
        }
    }
    __syncwarp();
    if (lane < cnt) sdf_splat<TRI>(rx[lane], ry[lane], rz[lane], X, a.vol);
    local += cnt;
    if (lane == 0 && local) atomicAdd(&a.frame_total[a.frame0 + f], local);
}

// ------------------------------------------------------------------------------------------------- rmsd(selection)
// _rmsd (md_script_functions.inl:4287-4345): the selection's atoms of the INITIAL frame and of the current frame, both wrapped into the
// current cell (md_util_pbc_vec4 md_util.c:8603), made whole along the bonds (md_util_unwrap_vec4 :8938, same local-index-as-atom quirk as
// in _sdf), centred on their plain centres of mass, fitted with mat3_optimal_rotation_vec4 (core/md_vec_math.c:337) and compared:
// sqrt(sum w |u - R v|^2 / sum w) with double sums (md_util_rmsd_compute_vec4 md_util.c:9037-9068).
//
// One warp per frame. The lanes extract and wrap the atoms (independent per atom); lane 0 then runs the parts whose result depends on
// the order of operations — the bond walk, the float centre-of-mass sums, the double covariance and deviation sums — exactly in the
// reference's order. Selections of rmsd() are one molecule or its backbone (10^2..10^4 atoms), a serial pass over them costs microseconds.
MDG_D float4 pbc_wrap(float4 v, const mdgpu_unitcell_t& uc) {
    if (uc.flags & MDGPU_CELL_ORTHO) {            // pbc_ortho_vec4 :8506-8512: vec4_deperiodize_ortho about the box centre
        const float ex = (float)uc.x, ey = (float)uc.y, ez = (float)uc.z;
        v.x = deperiodize1(v.x, ex * 0.5f, ex); v.y = deperiodize1(v.y, ey * 0.5f, ey); v.z = deperiodize1(v.z, ez * 0.5f, ez);
    } else if (uc.flags & MDGPU_CELL_TRICLINIC) {  // pbc_triclinic_vec4 :8554-8574: A * fract(I * r) on the periodic axes, float matrices
        const double i11 = uc.x > 0.0 ? 1.0 / uc.x : 0.0, i22 = uc.y > 0.0 ? 1.0 / uc.y : 0.0, i33 = uc.z > 0.0 ? 1.0 / uc.z : 0.0;   // md_unitcell.inl:158-176
        const double i12 = (uc.x * uc.y) > 0.0 ? -uc.xy / (uc.x * uc.y) : 0.0;
        const double i13 = (uc.x * uc.y * uc.z) > 0.0 ? (uc.xy * uc.yz - uc.xz * uc.y) / (uc.x * uc.y * uc.z) : 0.0;
        const double i23 = (uc.y * uc.z) > 0.0 ? -uc.yz / (uc.y * uc.z) : 0.0;
        const float I00 = (float)i11, I10 = (float)i12, I11 = (float)i22, I20 = (float)i13, I21 = (float)i23, I22 = (float)i33;
        const float A00 = (float)uc.x, A10 = (float)uc.xy, A11 = (float)uc.y, A20 = (float)uc.xz, A21 = (float)uc.yz, A22 = (float)uc.z;
        // linear_combine_3 (core/md_vec_math.h:1521): (x * col0 + y * col1) + z * col2, the zero entries of the matrices included
        float f0 = (v.x * I00 + v.y * I10) + v.z * I20;
        float f1 = (v.x * 0.0f + v.y * I11) + v.z * I21;
        float f2 = (v.x * 0.0f + v.y * 0.0f) + v.z * I22;
        f0 = f0 - floorf(f0); f1 = f1 - floorf(f1); f2 = f2 - floorf(f2);
        const float r0 = (f0 * A00 + f1 * A10) + f2 * A20;
        const float r1 = (f0 * 0.0f + f1 * A11) + f2 * A21;
        const float r2 = (f0 * 0.0f + f1 * 0.0f) + f2 * A22;
        if (uc.flags & MDGPU_CELL_PBC_X) v.x = r0;
        if (uc.flags & MDGPU_CELL_PBC_Y) v.y = r1;
        if (uc.flags & MDGPU_CELL_PBC_Z) v.z = r2;
    }
    return v;
}

// bond walk over (child, parent) pairs in BFS order + com_vec4, on atoms already in scratch (the second half of load_unwrap_com)
MDG_D void unwrap_com(float4* p, uint32_t n, const int2* pairs, uint32_t n_pairs, const mdgpu_unitcell_t& uc, float com[3]) {
    if (uc.flags & MDGPU_CELL_ORTHO) {
        const float ext[3] = { (float)uc.x, (float)uc.y, (float)uc.z };
        for (uint32_t k = 0; k < n_pairs; ++k) {
            const int2 pr = pairs[k];
            const float4 ref = p[pr.y]; float4 v = p[pr.x];
            v.x = deperiodize1(v.x, ref.x, ext[0]); v.y = deperiodize1(v.y, ref.y, ext[1]); v.z = deperiodize1(v.z, ref.z, ext[2]);
            p[pr.x] = v;
        }
    } else if (uc.flags & MDGPU_CELL_TRICLINIC) {
        const float box[3][3] = { { (float)uc.x, 0.f, 0.f }, { (float)uc.xy, (float)uc.y, 0.f }, { (float)uc.xz, (float)uc.yz, (float)uc.z } };
        for (uint32_t k = 0; k < n_pairs; ++k) {
            const int2 pr = pairs[k];
            const float4 ref = p[pr.y]; float4 v = p[pr.x];
            float d[3] = { __fsub_rn(v.x, ref.x), __fsub_rn(v.y, ref.y), __fsub_rn(v.z, ref.z) };
            min_image_triclinic(d, box);
            v.x = __fadd_rn(ref.x, d[0]); v.y = __fadd_rn(ref.y, d[1]); v.z = __fadd_rn(ref.z, d[2]);
            p[pr.x] = v;
        }
    }
    float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;   // com_vec4 md_util.c:8048-8061
    for (uint32_t k = 0; k < n; ++k) { const float4 v = p[k]; ax = ax + v.x * v.w; ay = ay + v.y * v.w; az = az + v.z * v.w; aw = aw + v.w * 1.0f; }
    com[0] = ax / aw; com[1] = ay / aw; com[2] = az / aw;
}

__global__ void __launch_bounds__(32) k_rmsd(RmsdArgs a, int B) {
    const int f = blockIdx.x, lane = threadIdx.x;
    if (f >= B) return;
    const uint32_t n = a.n;
    const mdgpu_unitcell_t uc = a.cells[f];
    const float* x = a.frames.xyz + (size_t)f * a.frames.frame_stride;
    float4* p0 = a.scratch_xyzw + (size_t)f * 2 * n;   // initial frame
    float4* p1 = p0 + n;                               // current frame
    for (uint32_t k = lane; k < n; k += 32) {           // extract_xyzw_vec4 (:966) + md_util_pbc_vec4
        const int at = a.idx[k]; const float w = a.mass[at];
        p0[k] = pbc_wrap(make_float4(a.init_xyz[at], a.init_xyz[a.init_axis_stride + at], a.init_xyz[2 * a.init_axis_stride + at], w), uc);
        p1[k] = pbc_wrap(make_float4(x[at], x[a.frames.axis_stride + at], x[2 * a.frames.axis_stride + at], w), uc);
    }
    __syncwarp();
    if (lane != 0) return;
    float com0[3], com1[3];
    unwrap_com(p0, n, a.unwrap_pairs, a.n_unwrap, uc, com0);
    unwrap_com(p1, n, a.unwrap_pairs, a.n_unwrap, uc, com1);
    double C[3][3] = { { 0 } }; double ws = 0.0;   // mat3_cross_covariance_matrix_vec4 (core/md_vec_math.c:256-289)
    for (uint32_t k = 0; k < n; ++k) {
        const float4 u = p0[k], v = p1[k];
        const float px = u.x - com0[0], py = u.y - com0[1], pz = u.z - com0[2], pw = u.w - 0.0f;
        const float qx = v.x - com1[0], qy = v.y - com1[1], qz = v.z - com1[2], qw = v.w - 0.0f;
        const float w = (pw + qw) * 0.5f;
        C[0][0] += w * px * qx; C[0][1] += w * px * qy; C[0][2] += w * px * qz;
        C[1][0] += w * py * qx; C[1][1] += w * py * qy; C[1][2] += w * py * qz;
        C[2][0] += w * pz * qx; C[2][1] += w * pz * qy; C[2][2] += w * pz * qz;
        ws += w;
    }
    M3 cc; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cc.e[i][j] = (float)(C[i][j] / ws);
    const Svd sv = m3_svd(cc);                      // mat3_extract_rotation (core/md_vec_math.c:292-300)
    const M3 Ut = m3_transpose(sv.U);
    const float det = m3_det(m3_mul(sv.V, Ut));
    M3 D; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) D.e[i][j] = 0.0f; D.e[0][0] = 1.0f; D.e[1][1] = 1.0f; D.e[2][2] = (float)((det > 0.0f) - (det < 0.0f));
    const M3 R = m3_mul(m3_mul(sv.V, D), Ut);
    double d_sum = 0.0, w_sum = 0.0;
    for (uint32_t k = 0; k < n; ++k) {
        const float4 u4 = p0[k], v4 = p1[k];
        const float u[3] = { u4.x - com0[0], u4.y - com0[1], u4.z - com0[2] };
        const float v[3] = { v4.x - com1[0], v4.y - com1[1], v4.z - com1[2] };
        float d[3];   // mat3_mul_vec3 (core/md_vec_math.h:1623): (x * col0 + y * col1) + z * col2
        for (int r = 0; r < 3; ++r) d[r] = u[r] - ((R.e[0][r] * v[0] + R.e[1][r] * v[1]) + R.e[2][r] * v[2]);
        const float w = (u4.w + v4.w) * 0.5f;
        const float dd = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
        d_sum += (double)(w * dd); w_sum += (double)w;
    }
    a.out[a.frame0 + f] = (float)sqrt(d_sum / w_sum);
}

// plane(selection) (_plane md_script_functions.inl:4755-4822): positions with unit weights, bond walk, plain centre, covariance
// (mat3_covariance_matrix_vec4 core/md_vec_math.c:101-156), eigenvectors sorted by eigenvalue (mat3_eigen :22-42); the frame's row is
// (normalised third axis, normal . centre). One warp per frame, lane 0 does the ordered part, as in k_rmsd (init_xyz is not used).
__global__ void __launch_bounds__(32) k_plane(RmsdArgs a, int B) {
    const int f = blockIdx.x, lane = threadIdx.x;
    if (f >= B) return;
    const uint32_t n = a.n;
    const mdgpu_unitcell_t uc = a.cells[f];
    const float* x = a.frames.xyz + (size_t)f * a.frames.frame_stride;
    float4* p = a.scratch_xyzw + (size_t)f * n;
    if (a.pos) for (uint32_t k = lane; k < n; k += 32) { const float* q = a.pos + ((size_t)f * n + k) * 3; p[k] = make_float4(q[0], q[1], q[2], 1.0f); }   // one centre of mass per selection (coordinate_extract :1503)
    else for (uint32_t k = lane; k < n; k += 32) { const int at = a.idx[k]; p[k] = make_float4(x[at], x[a.frames.axis_stride + at], x[2 * a.frames.axis_stride + at], 1.0f); }
    __syncwarp();
    if (lane != 0) return;
    float com[3];
    unwrap_com(p, n, a.unwrap_pairs, a.n_unwrap, uc, com);
    double C[3][3] = { { 0 } }; double ws = 0.0;
    for (uint32_t k = 0; k < n; ++k) {
        const float4 v = p[k];
        const float px = v.x - com[0], py = v.y - com[1], pz = v.z - com[2], w = v.w;
        C[0][0] += w * px * px; C[0][1] += w * px * py; C[0][2] += w * px * pz;
        C[1][0] += w * py * px; C[1][1] += w * py * py; C[1][2] += w * py * pz;
        C[2][0] += w * pz * px; C[2][1] += w * pz * py; C[2][2] += w * pz * pz;
        ws += w;
    }
    M3 cov; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov.e[i][j] = (float)(C[i][j] / ws);
    const Svd s = m3_svd(cov);
    const float mx = fmaxf(s.s[0], fmaxf(s.s[1], s.s[2]));
    const float ev[3] = { s.s[0] / mx, s.s[1] / mx, s.s[2] / mx };
    int l0 = 0, l1 = 1, l2 = 2, t;
    if (ev[l0] < ev[l1]) { t = l0; l0 = l1; l1 = t; }
    if (ev[l1] < ev[l2]) { t = l1; l1 = l2; l2 = t; }
    if (ev[l0] < ev[l1]) { t = l0; l0 = l1; l1 = t; }
    float nx = s.U.e[l2][0], ny = s.U.e[l2][1], nz = s.U.e[l2][2];
    const float len = __fsqrt_rn((nx * nx + ny * ny) + nz * nz);   // vec3_normalize core/md_vec_math.h:505-514 (the threshold is a double literal)
    if ((double)len > 1.0e-5) { nx = nx / len; ny = ny / len; nz = nz / len; } else { nx = ny = nz = 0.0f; }
    float* o = a.out + (size_t)(a.frame0 + f) * 4;
    o[0] = nx; o[1] = ny; o[2] = nz; o[3] = (nx * com[0] + ny * com[1]) + nz * com[2];
}

// ------------------------------------------------------------------------------------------------- shape weights of structures
// The loop body of VIAMD's shape-space component (src/components/shapespace/shapespace.cpp:418-431) and of _shape_weights
// (md_script_functions.inl:6033-6040): xyzw (mass or 1) -> md_util_com_compute_vec4 with the cell (com_pbc_vec4 md_util.c:8063-8162) ->
// md_util_deperiodize_vec4 about that centre (:8971-9005) -> mat3_covariance_matrix_vec4 -> md_util_shape_weights (:9070-9076).
// One lane of the 4-lane md_mm_sincos_ps (core/md_simd.h:1093-1176): the Cephes sequence of props.cu's ref_sincosf with the third
// Cody-Waite constant as that variant spells it (:1137).
MDG_D void sincos_cephes4(float x, float& out_s, float& out_c) {
    uint32_t sign_bit_sin = __float_as_uint(x) & 0x80000000u;
    x = fabsf(x);
    float y = __fmul_rn(x, 1.27323954473516f);
    int imm2 = __float2int_rz(y);
    imm2 = (imm2 + 1) & ~1;
    y = (float)imm2;
    const uint32_t swap_sign_bit_sin = ((uint32_t)(imm2 & 4)) << 29;
    const bool poly_mask = (imm2 & 2) == 0;
    const uint32_t sign_bit_cos = ((uint32_t)(~(imm2 - 2) & 4)) << 29;
    sign_bit_sin ^= swap_sign_bit_sin;
    x = __fmaf_rn(y, -0.78515625f, x);
    x = __fmaf_rn(y, -2.4187564849853515625e-4f, x);
    x = __fmaf_rn(y, -3.77489497744594108e-8f, x);
    const float x2 = __fmul_rn(x, x), x3 = __fmul_rn(x2, x), x4 = __fmul_rn(x2, x2);
    y = __fmaf_rn(x2, __fmaf_rn(x2, 2.443315711809948E-005f, -1.388731625493765E-003f), 4.166664568298827E-002f);
    y = __fmaf_rn(x2, -0.5f, __fmul_rn(y, x4));
    y = __fadd_rn(y, 1.0f);
    float y2 = __fmaf_rn(x2, __fmaf_rn(x2, -1.9515295891E-4f, 8.3321608736E-3f), -1.6666654611E-1f);
    y2 = __fmaf_rn(y2, x3, x);
    const float ysin2 = poly_mask ? y2 : 0.0f, ysin1 = poly_mask ? 0.0f : y;
    y2 = __fsub_rn(y2, ysin2); y = __fsub_rn(y, ysin1);
    out_s = __uint_as_float(__float_as_uint(__fadd_rn(ysin1, ysin2)) ^ sign_bit_sin);
    out_c = __uint_as_float(__float_as_uint(__fadd_rn(y, y2)) ^ sign_bit_cos);
}

// md_util_com_compute_vec4 (md_util.c:8188-8201) of n points xyzw: com_pbc_vec4 (:8063-8162) in a cell — serial float sums of w*sin, w*cos
// per axis in index order, 4-lane sincos, double atan2; the triclinic branch as written (in_idx == NULL: theta through the 1/2pi-scaled
// inverse, and the result through it again) — com_vec4 (:8048) without one. One thread.
MDG_D void com_compute_vec4(const float4* p, uint32_t n, const mdgpu_unitcell_t& uc, float com[3]) {
    const double TWO_PI_D = 2.0 * 3.1415926535897932, PI_D = 3.1415926535897932;
    if (uc.flags & MDGPU_CELL_ORTHO) {
        const float ext[3] = { (float)uc.x, (float)uc.y, (float)uc.z };
        const float tp = (float)TWO_PI_D;
        const float scl[4] = { tp / ext[0], tp / ext[1], tp / ext[2], tp / tp };
        float as[4] = { 0.f, 0.f, 0.f, 0.f }, ac[4] = { 0.f, 0.f, 0.f, 0.f }, ax[4] = { 0.f, 0.f, 0.f, 0.f };
        for (uint32_t k = 0; k < n; ++k) {
            const float4 v = p[k]; const float e[4] = { v.x, v.y, v.z, v.w }, www1[4] = { v.w, v.w, v.w, 1.0f };
            for (int c = 0; c < 4; ++c) {
                float sn, cs; sincos_cephes4(e[c] * scl[c], sn, cs);
                as[c] = as[c] + sn * www1[c]; ac[c] = ac[c] + cs * www1[c]; ax[c] = ax[c] + e[c] * www1[c];
            }
        }
        const float w = ax[3];
        for (int c = 0; c < 3; ++c) {
            const double yy = (double)(as[c] / w), xx = (double)(ac[c] / w), r2 = xx * xx + yy * yy;
            double theta = PI_D; if (r2 > 1.0e-15) theta += atan2(-yy, -xx);
            com[c] = (float)((theta / TWO_PI_D) * (double)ext[c]);
        }
    } else if (uc.flags & MDGPU_CELL_TRICLINIC) {
        const double i11 = uc.x > 0.0 ? 1.0 / uc.x : 0.0, i22 = uc.y > 0.0 ? 1.0 / uc.y : 0.0, i33 = uc.z > 0.0 ? 1.0 / uc.z : 0.0;   // md_unitcell.inl:158-176
        const double i12 = (uc.x * uc.y) > 0.0 ? -uc.xy / (uc.x * uc.y) : 0.0;
        const double i13 = (uc.x * uc.y * uc.z) > 0.0 ? (uc.xy * uc.yz - uc.xz * uc.y) / (uc.x * uc.y * uc.z) : 0.0;
        const double i23 = (uc.y * uc.z) > 0.0 ? -uc.yz / (uc.y * uc.z) : 0.0;
        const float Ai[3][3] = { { (float)i11, 0.f, 0.f }, { (float)i12, (float)i22, 0.f }, { (float)i13, (float)i23, (float)i33 } };   // [col][row]
        const float inv_tp = 1.0f / (float)TWO_PI_D;
        float I[3][3];   // mat3_mul(mat3_scale(1/2pi), Ai) (core/md_vec_math.h:1631): the three products of MULT(col,row), two of them with a zero factor
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) {
            const float s0 = (r == 0) ? inv_tp : 0.0f, s1 = (r == 1) ? inv_tp : 0.0f, s2 = (r == 2) ? inv_tp : 0.0f;
            I[c][r] = (s0 * Ai[c][0] + s1 * Ai[c][1]) + s2 * Ai[c][2];
        }
        float as[4] = { 0.f, 0.f, 0.f, 0.f }, ac[4] = { 0.f, 0.f, 0.f, 0.f }, ax[4] = { 0.f, 0.f, 0.f, 0.f };
        for (uint32_t k = 0; k < n; ++k) {
            const float4 v = p[k]; const float e[4] = { v.x, v.y, v.z, v.w }, www1[4] = { v.w, v.w, v.w, 1.0f };
            float th[4];   // mat4x3_mul_vec4(I, xyzw), the in_idx == NULL branch (:8138)
            for (int c = 0; c < 3; ++c) th[c] = (v.x * I[0][c] + v.y * I[1][c]) + v.z * I[2][c];
            th[3] = (v.x * 0.0f + v.y * 0.0f) + v.z * 0.0f;
            for (int c = 0; c < 4; ++c) {
                float sn, cs; sincos_cephes4(th[c], sn, cs);
                as[c] = as[c] + sn * www1[c]; ac[c] = ac[c] + cs * www1[c]; ax[c] = ax[c] + e[c] * www1[c];
            }
        }
        for (int c = 0; c < 3; ++c) {
            const double yy = (double)(as[c] / ax[3]), xx = (double)(ac[c] / ax[3]), r2 = xx * xx + yy * yy;
            double theta = PI_D; if (r2 > 1.0e-8) theta += atan2(-yy, -xx);
            com[c] = (float)(theta * (double)I[c][0] + theta * (double)I[c][1] + theta * (double)I[c][2]);   // :8158, as written
        }
    } else {   // no cell: com_vec4, nothing to deperiodize
        float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
        for (uint32_t k = 0; k < n; ++k) { const float4 v = p[k]; ax = ax + v.x * v.w; ay = ay + v.y * v.w; az = az + v.z * v.w; aw = aw + v.w * 1.0f; }
        com[0] = ax / aw; com[1] = ay / aw; com[2] = az / aw;
    }
}

// One warp per (structure, frame): the lanes extract, lane 0 does the ordered work (float sums over the atoms in index order, double
// covariance). Row (frame0 + f) of the property holds n_struct x (linear, planar, isotropic).
__global__ void __launch_bounds__(32) k_shape_weights(ShapeArgs a, int B) {
    const int f = blockIdx.y, lane = threadIdx.x;
    const uint32_t sidx = blockIdx.x;
    if (f >= B || sidx >= a.n_struct) return;
    const uint32_t beg = a.soff[sidx], n = a.soff[sidx + 1] - beg;
    float* o = a.out + ((size_t)(a.frame0 + f) * a.n_struct + sidx) * 3;
    if (n == 0) { if (lane == 0) { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; } return; }   // count == 0: the entry keeps its zero (:6029)
    const mdgpu_unitcell_t uc = a.cells[f];
    const float* x = a.frames.xyz + (size_t)f * a.frames.frame_stride;
    float4* p = a.scratch_xyzw + (size_t)f * a.n_atoms_total + beg;
    for (uint32_t k = lane; k < n; k += 32) { const int at = a.idx[beg + k]; p[k] = make_float4(x[at], x[a.frames.axis_stride + at], x[2 * a.frames.axis_stride + at], a.use_mass ? a.mass[at] : 1.0f); }
    __syncwarp();
    if (lane != 0) return;
    float com[3];
    com_compute_vec4(p, n, uc, com);
    if (uc.flags & MDGPU_CELL_ORTHO) {
        const float ext[3] = { (float)uc.x, (float)uc.y, (float)uc.z };
        for (uint32_t k = 0; k < n; ++k) { float4 v = p[k]; v.x = deperiodize1(v.x, com[0], ext[0]); v.y = deperiodize1(v.y, com[1], ext[1]); v.z = deperiodize1(v.z, com[2], ext[2]); p[k] = v; }
    } else if (uc.flags & MDGPU_CELL_TRICLINIC) {
        const float box[3][3] = { { (float)uc.x, 0.f, 0.f }, { (float)uc.xy, (float)uc.y, 0.f }, { (float)uc.xz, (float)uc.yz, (float)uc.z } };
        for (uint32_t k = 1; k < n; ++k) {   // deperiodize_triclinic from atom 1 on (:8993)
            float4 v = p[k];
            float d[3] = { __fsub_rn(v.x, com[0]), __fsub_rn(v.y, com[1]), __fsub_rn(v.z, com[2]) };
            min_image_triclinic(d, box);
            v.x = __fadd_rn(com[0], d[0]); v.y = __fadd_rn(com[1], d[1]); v.z = __fadd_rn(com[2], d[2]); p[k] = v;
        }
    }
    double C[3][3] = { { 0 } }; double ws = 0.0;   // mat3_covariance_matrix_vec4 (core/md_vec_math.c:101-156)
    for (uint32_t k = 0; k < n; ++k) {
        const float4 v = p[k];
        const float px = v.x - com[0], py = v.y - com[1], pz = v.z - com[2], w = v.w;
        C[0][0] += w * px * px; C[0][1] += w * px * py; C[0][2] += w * px * pz;
        C[1][0] += w * py * px; C[1][1] += w * py * py; C[1][2] += w * py * pz;
        C[2][0] += w * pz * px; C[2][1] += w * pz * py; C[2][2] += w * pz * pz;
        ws += w;
    }
    M3 cov; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov.e[i][j] = (float)(C[i][j] / ws);
    const Svd sv = m3_svd(cov);   // mat3_eigen (:22-42): values normalised by the largest, sorted descending
    const float mx = fmaxf(sv.s[0], fmaxf(sv.s[1], sv.s[2]));
    const float ev[3] = { sv.s[0] / mx, sv.s[1] / mx, sv.s[2] / mx };
    int l0 = 0, l1 = 1, l2 = 2, t;
    if (ev[l0] < ev[l1]) { t = l0; l0 = l1; l1 = t; }
    if (ev[l1] < ev[l2]) { t = l1; l1 = l2; l2 = t; }
    if (ev[l0] < ev[l1]) { t = l0; l0 = l1; l1 = t; }
    const float e0 = ev[l0], e1 = ev[l1], e2 = ev[l2];
    const float sc = 1.0f / ((e0 + e1) + e2);   // md_util_shape_weights md_util.c:9070-9076
    o[0] = (e0 - e1) * sc; o[1] = 2.0f * (e1 - e2) * sc; o[2] = 3.0f * e2 * sc;
}

// An ARRAY of selections as one position argument of distance / angle / dihedral / com: the centres k_arg_com_parts left in `parts`
// (weight 1 each) -> md_util_com_compute_vec4 with the frame's cell (coordinate_extract_com md_script_functions.inl:1841). One thread per frame.
__global__ void k_arg_combine(const float4* __restrict__ parts, uint32_t n_parts, const mdgpu_unitcell_t* __restrict__ cells, float* __restrict__ out /* [B][4][3] */, int arg, int B) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= B) return;
    float com[3];
    com_compute_vec4(parts + (size_t)f * n_parts, n_parts, cells[f], com);
    float* o = out + ((size_t)f * 4 + arg) * 3;
    o[0] = com[0]; o[1] = com[1]; o[2] = com[2];
}

void launch_arg_combine(const float4* d_parts, uint32_t n_parts, const mdgpu_unitcell_t* d_cells, float* d_out, int arg, int B, cudaStream_t s) {
    if (B <= 0 || !n_parts) return;
    k_arg_combine<<<(B + 63) / 64, 64, 0, s>>>(d_parts, n_parts, d_cells, d_out, arg, B);
    note_launch("k_arg_combine", s);
}

void launch_shape_weights(const ShapeArgs& a, int B, cudaStream_t s) {
    if (!a.n_struct || B <= 0) return;
    k_shape_weights<<<dim3(a.n_struct, (unsigned)B), 32, 0, s>>>(a, B);
    note_launch("k_shape_weights", s);
}

void launch_plane(const RmsdArgs& a, int B, cudaStream_t s) {
    if (!a.n || B <= 0) return;
    k_plane<<<B, 32, 0, s>>>(a, B);
    note_launch("k_plane", s);
}

void launch_rmsd(const RmsdArgs& a, int B, cudaStream_t s) {
    if (!a.n || B <= 0) return;   // empty selection: the property stays 0 (:4311)
    k_rmsd<<<B, 32, 0, s>>>(a, B);
    note_launch("k_rmsd", s);
}

void launch_sdf(const SdfArgs& a, int B, bool tri, cudaStream_t s) {
    k_sdf_ref0<<<(B + 31) / 32, 32, 0, s>>>(a, B);
    note_launch("k_sdf_ref0", s);
    dim3 g1((a.n_struct + 63) / 64, B);
    k_sdf_fit<<<g1, 64, 0, s>>>(a, B);
    note_launch("k_sdf_fit", s);
    dim3 g2((a.n_struct + SDF_WARPS - 1) / SDF_WARPS, B);
    static const bool ring = []() { const char* e = getenv("MDGPU_SDF"); return e && strcmp(e, "ring") == 0; }();
    if (ring) { if (tri) k_sdf_scatter_ring<true><<<g2, SDF_WARPS * 32, 0, s>>>(a, B); else k_sdf_scatter_ring<false><<<g2, SDF_WARPS * 32, 0, s>>>(a, B); }
    else {
        static const int occ = []() { const char* e = getenv("MDGPU_SDF_OCC"); return e ? atoi(e) : 4; }();   // resident CTAs / SM the register allocation aims for
        if (occ >= 5)      { if (tri) k_sdf_scatter<true, 5><<<g2, SDF_WARPS * 32, 0, s>>>(a, B); else k_sdf_scatter<false, 5><<<g2, SDF_WARPS * 32, 0, s>>>(a, B); }
        else if (occ == 4) { if (tri) k_sdf_scatter<true, 4><<<g2, SDF_WARPS * 32, 0, s>>>(a, B); else k_sdf_scatter<false, 4><<<g2, SDF_WARPS * 32, 0, s>>>(a, B); }
        else               { if (tri) k_sdf_scatter<true, 3><<<g2, SDF_WARPS * 32, 0, s>>>(a, B); else k_sdf_scatter<false, 3><<<g2, SDF_WARPS * 32, 0, s>>>(a, B); }
    }
    note_launch("k_sdf_scatter", s);
}

}  // namespace mdg
