/* md_oracle.c — plain-C CPU restatement of the reference hot path. TEST INFRASTRUCTURE ONLY (see md_oracle.h).
 *
 * Written from the reference's behaviour, scalar and in program order, so that every float operation that
 * decides a bin / voxel index is performed with the same operands, in the same order and with the same
 * rounding as the reference's strict (non -ffast-math) build:
 *   - a*b+c written in the reference as separate mul and add stays separate (compile with -ffp-contract=off);
 *   - explicit fmadd intrinsics in the reference are fmaf() here;
 *   - double accumulators / double setup stay double.
 * Paths cited below are relative to /root/reference/ext/mdlib/src.
 */
#include "md_oracle.h"

#include <float.h>
#include <math.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>

#define MAXV(a, b) ((a) > (b) ? (a) : (b))
#define MINV(a, b) ((a) < (b) ? (a) : (b))
#define CLAMPV(v, lo, hi) MINV(MAXV(v, lo), hi)

/* ------------------------------------------------------------------------------------------------
 * Unit cell matrices: md_unitcell_A_extract_double / md_unitcell_I_extract_double (md_unitcell.inl:129-175)
 * A[col][row]; columns are the basis vectors.
 */
static void cell_A(double A[3][3], const mdo_unitcell_t* c) {
    A[0][0] = c->x;  A[0][1] = 0;     A[0][2] = 0;
    A[1][0] = c->xy; A[1][1] = c->y;  A[1][2] = 0;
    A[2][0] = c->xz; A[2][1] = c->yz; A[2][2] = c->z;
}
static void cell_I(double I[3][3], const mdo_unitcell_t* c) {
    if (!c->flags) { memset(I, 0, sizeof(double) * 9); return; }
    const double i11 = c->x > 0.0 ? 1.0 / c->x : 0.0;
    const double i22 = c->y > 0.0 ? 1.0 / c->y : 0.0;
    const double i33 = c->z > 0.0 ? 1.0 / c->z : 0.0;
    const double i12 = (c->x * c->y) > 0.0 ? -c->xy / (c->x * c->y) : 0.0;
    const double i13 = (c->x * c->y * c->z) > 0.0 ? (c->xy * c->yz - c->xz * c->y) / (c->x * c->y * c->z) : 0.0;
    const double i23 = (c->y * c->z) > 0.0 ? -c->yz / (c->y * c->z) : 0.0;
    I[0][0] = i11; I[0][1] = 0.0; I[0][2] = 0.0;
    I[1][0] = i12; I[1][1] = i22; I[1][2] = 0.0;
    I[2][0] = i13; I[2][1] = i23; I[2][2] = i33;
}

/* ------------------------------------------------------------------------------------------------
 * Spatial acceleration structure: md_spatial_acc_init (core/md_spatial_acc.c:155-438)
 */
typedef struct acc_t {
    size_t num_elems, num_cells;
    float *ex, *ey, *ez; uint32_t* eidx; uint32_t* cell_off;
    uint32_t cell_dim[3]; float inv_cell_ext[3];
    float G00, G11, G22, H01, H02, H12;
    float A[3][3], I[3][3], origin[3];
    uint32_t flags;
} acc_t;

static void acc_free(acc_t* a) { free(a->ex); free(a->ey); free(a->ez); free(a->eidx); free(a->cell_off); memset(a, 0, sizeof(*a)); }

/* coordinate stream: SoA x/y/z + optional index, or AoS xyz (md_coord_stream_t, core/md_spatial_acc.h:15-37) */
typedef struct stream_t { const float *x, *y, *z; const int32_t* idx; const float* aos; size_t count; } stream_t;
static inline void stream_load(const stream_t* s, size_t i, float r[3]) {
    if (s->aos) { r[0] = s->aos[3 * i]; r[1] = s->aos[3 * i + 1]; r[2] = s->aos[3 * i + 2]; return; }
    const size_t src = s->idx ? (size_t)s->idx[i] : i;
    r[0] = s->x[src]; r[1] = s->y[src]; r[2] = s->z[src];
}

/* vec4_linear_combine_3(r - origin, I) (core/md_vec_math.h:1323): ((I0*a.x) + (I1*a.y)) + (I2*a.z), component-wise */
static inline void cart_to_fract_f(float s[3], const float r[3], const float origin[3], const float I[3][3]) {
    const float ax = r[0] - origin[0], ay = r[1] - origin[1], az = r[2] - origin[2];
    for (int k = 0; k < 3; ++k) {
        float v = I[0][k] * ax;
        v = v + I[1][k] * ay;
        v = v + I[2][k] * az;
        s[k] = v;
    }
}

static void acc_init(acc_t* acc, const stream_t* st, double in_cell_ext, const mdo_unitcell_t* cell, bool use_supplied_idx) {
    memset(acc, 0, sizeof(*acc));
    if (st->count == 0) return;                                   /* :174 */
    if (in_cell_ext <= 0.0) in_cell_ext = 6.0;                    /* :179 */
    const double CELL_EXT = MAXV(in_cell_ext, 3.0);               /* :187 */
    double A[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } }, I[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
    uint32_t flags = 0;
    if (cell) { cell_A(A, cell); cell_I(I, cell); flags = cell->flags; }
    float origin[3] = { 0, 0, 0 };

    if ((flags & MDO_CELL_PBC_ALL) != MDO_CELL_PBC_ALL) {         /* :201-243 : AABB fit on non-periodic axes */
        float mn[3] = { 0, 0, 0 }, mx[3] = { 0, 0, 0 };           /* NB: the box starts at the origin (vec4 {0}) */
        for (size_t i = 0; i < st->count; ++i) {
            float r[3]; stream_load(st, i, r);
            for (int k = 0; k < 3; ++k) { mn[k] = MINV(mn[k], r[k]); mx[k] = MAXV(mx[k], r[k]); }
        }
        for (int k = 0; k < 3; ++k) {
            float ext = mx[k] - mn[k];
            ext = ceilf(ext / (float)CELL_EXT) * (float)CELL_EXT;
            const float cen = (mn[k] + mx[k]) * 0.5f;
            const float lo = cen - ext * 0.5f;
            if ((flags & (MDO_CELL_PBC_X << k)) == 0) {
                origin[k] = lo;
                if (ext > 0.0f) { A[k][k] = ext; I[k][k] = 1.0 / ext; }
            }
        }
    }
    const double a[3] = { A[0][0], A[0][1], A[0][2] }, b[3] = { A[1][0], A[1][1], A[1][2] }, c[3] = { A[2][0], A[2][1], A[2][2] };
    const double G00 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    const double G11 = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
    const double G22 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    const double G01 = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    const double G02 = a[0] * c[0] + a[1] * c[1] + a[2] * c[2];
    const double G12 = b[0] * c[0] + b[1] * c[1] + b[2] * c[2];
    double H01 = 0, H02 = 0, H12 = 0;
    const double na = sqrt(G00), nb = sqrt(G11), nc = sqrt(G22);
    float inv_cell_ext[3] = { (float)(na > 0.0 ? 1.0 / na : 0.0), (float)(nb > 0.0 ? 1.0 / nb : 0.0), (float)(nc > 0.0 ? 1.0 / nc : 0.0) };
    if (flags & MDO_CELL_TRICLINIC) {                             /* :270-290 */
        H01 = 2.0 * G01; H02 = 2.0 * G02; H12 = 2.0 * G12;
        const double det = G00 * (G11 * G22 - G12 * G12) - G01 * (G01 * G22 - G12 * G02) + G02 * (G01 * G12 - G11 * G02);
        if (det < DBL_EPSILON) return;
        inv_cell_ext[0] = (float)sqrt((G11 * G22 - G12 * G12) / det);
        inv_cell_ext[1] = (float)sqrt((G00 * G22 - G02 * G02) / det);
        inv_cell_ext[2] = (float)sqrt((G00 * G11 - G01 * G01) / det);
    }
    uint32_t cd[3] = { (uint32_t)(na / CELL_EXT), (uint32_t)(nb / CELL_EXT), (uint32_t)(nc / CELL_EXT) };   /* :294-298 */
    for (int k = 0; k < 3; ++k) cd[k] = CLAMPV(cd[k], 1u, 1024u);
    const uint32_t c0 = cd[0], c01 = cd[0] * cd[1];
    const size_t num_cells = (size_t)cd[0] * cd[1] * cd[2];

    const size_t n = st->count;
    uint32_t* local_idx = malloc(n * 4); uint32_t* cell_idx = malloc(n * 4);
    float* sx = malloc(n * 4); float* sy = malloc(n * 4); float* sz = malloc(n * 4); uint32_t* sidx = malloc(n * 4);
    acc->ex = calloc(n + 16, 4); acc->ey = calloc(n + 16, 4); acc->ez = calloc(n + 16, 4); acc->eidx = calloc(n + 16, 4);
    acc->cell_off = calloc(num_cells + 1, 4);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { acc->A[i][j] = (float)A[i][j]; acc->I[i][j] = (float)I[i][j]; }
    for (int k = 0; k < 3; ++k) acc->origin[k] = origin[k];

    for (size_t i = 0; i < n; ++i) {                              /* :341-371 */
        float r[3], s[3]; stream_load(st, i, r);
        cart_to_fract_f(s, r, origin, acc->I);
        uint32_t cc[3];
        for (int k = 0; k < 3; ++k) {
            if (flags & (MDO_CELL_PBC_X << k)) s[k] = s[k] - floorf(s[k]);     /* vec4_fract, blended by pbc mask */
            int ic = (int)floorf(s[k] * (float)cd[k]);                          /* cvtps of an integral float */
            ic = CLAMPV(ic, 0, (int)cd[k] - 1);
            cc[k] = (uint32_t)ic;
        }
        const size_t ci = (size_t)cc[2] * c01 + (size_t)cc[1] * c0 + cc[0];
        local_idx[i] = acc->cell_off[ci]++;
        cell_idx[i] = (uint32_t)ci;
        sx[i] = s[0]; sy[i] = s[1]; sz[i] = s[2];
        sidx[i] = (use_supplied_idx && st->idx) ? (uint32_t)st->idx[i] : (uint32_t)i;
    }
    uint32_t sum = 0;                                             /* :373-379 */
    for (size_t ci = 0; ci <= num_cells; ++ci) { uint32_t len = acc->cell_off[ci]; acc->cell_off[ci] = sum; sum += len; }
    for (size_t i = 0; i < n; ++i) {                              /* :383-390 : input order preserved inside a cell */
        const uint32_t dst = acc->cell_off[cell_idx[i]] + local_idx[i];
        acc->ex[dst] = sx[i]; acc->ey[dst] = sy[i]; acc->ez[dst] = sz[i]; acc->eidx[dst] = sidx[i];
    }
    free(local_idx); free(cell_idx); free(sx); free(sy); free(sz); free(sidx);
    acc->num_elems = n; acc->num_cells = num_cells;
    memcpy(acc->cell_dim, cd, sizeof(cd)); memcpy(acc->inv_cell_ext, inv_cell_ext, sizeof(inv_cell_ext));
    acc->G00 = (float)G00; acc->G11 = (float)G11; acc->G22 = (float)G22;
    acc->H01 = (float)H01; acc->H02 = (float)H02; acc->H12 = (float)H12;
    acc->flags = flags;
}

/* calc_r2 core/md_spatial_acc.c:541-544 */
static float calc_r2(double cutoff) { float r2 = (float)(cutoff * cutoff); return nextafterf(r2, r2 + 1.0f); }

typedef void (*pair_cb_t)(uint32_t i, uint32_t j, float d2, void* user);

/* for_each_external_pair_within_cutoff_ortho / _triclinic (core/md_spatial_acc.c:1649-1803, 1498-1647) */
static void acc_ext_pairs(const acc_t* acc, const stream_t* ext, double cutoff, bool use_supplied_idx, pair_cb_t cb, void* user) {
    if (acc->num_elems == 0) return;
    int ncell[3];
    for (int k = 0; k < 3; ++k) ncell[k] = (int)ceil(cutoff * (double)acc->inv_cell_ext[k] * acc->cell_dim[k]);
    for (int k = 0; k < 3; ++k) if (2 * ncell[k] + 1 > 5) return;  /* "cutoff too large for cell size": logs an error, yields no pairs */
    const float r2 = calc_r2(cutoff);
    const bool tri = (acc->flags & MDO_CELL_TRICLINIC) != 0;
    const int cd[3] = { (int)acc->cell_dim[0], (int)acc->cell_dim[1], (int)acc->cell_dim[2] };
    const uint32_t c0 = acc->cell_dim[0], c01 = acc->cell_dim[0] * acc->cell_dim[1];

    for (size_t ei = 0; ei < ext->count; ++ei) {
        float r[3], f[3]; stream_load(ext, ei, r);
        cart_to_fract_f(f, r, acc->origin, acc->I);
        int cv[3];
        for (int k = 0; k < 3; ++k) {
            if (!tri && (acc->flags & (MDO_CELL_PBC_X << k))) f[k] = f[k] - floorf(f[k]);   /* ortho only (:1718); the triclinic path does not wrap (:1565) */
            cv[k] = (int)floorf(f[k] * (float)cd[k]);                                       /* NOT clamped */
        }
        const uint32_t idx_i = (use_supplied_idx && ext->idx) ? (uint32_t)ext->idx[ei] : (uint32_t)ei;
        for (int dz = -ncell[2]; dz <= ncell[2]; ++dz)
        for (int dy = -ncell[1]; dy <= ncell[1]; ++dy)
        for (int dx = -ncell[0]; dx <= ncell[0]; ++dx) {
            const int off[3] = { dx, dy, dz };
            int nv[3]; float fs[3]; bool skip = false;
            for (int k = 0; k < 3; ++k) {
                nv[k] = cv[k] + off[k];
                const bool up = nv[k] > cd[k] - 1, lo = nv[k] < 0;
                if ((up || lo) && !tri && !(acc->flags & (MDO_CELL_PBC_X << k))) skip = true;   /* :1733 */
                if (lo) nv[k] += cd[k];
                if (up) nv[k] -= cd[k];
                fs[k] = f[k] + (float)((lo ? 1 : 0) - (up ? 1 : 0));                              /* :1750-1755 */
            }
            if (skip) continue;
            /* the reference wraps once only; an index still out of range would be an out-of-bounds read there */
            if (nv[0] < 0 || nv[0] >= cd[0] || nv[1] < 0 || nv[1] >= cd[1] || nv[2] < 0 || nv[2] >= cd[2]) continue;
            const size_t cj = (size_t)nv[2] * c01 + (size_t)nv[1] * c0 + (size_t)nv[0];
            const uint32_t o = acc->cell_off[cj], len = acc->cell_off[cj + 1] - o;
            for (uint32_t j = 0; j < len; ++j) {
                const float ddx = fs[0] - acc->ex[o + j], ddy = fs[1] - acc->ey[o + j], ddz = fs[2] - acc->ez[o + j];
                const float dx2 = ddx * ddx, dy2 = ddy * ddy, dz2 = ddz * ddz;
                float d2 = fmaf(acc->G00, dx2, fmaf(acc->G11, dy2, acc->G22 * dz2));              /* :517-528 */
                if (tri) {
                    const float dxy = ddx * ddy, dxz = ddx * ddz, dyz = ddy * ddz;
                    const float cross = fmaf(acc->H01, dxy, fmaf(acc->H02, dxz, acc->H12 * dyz));  /* :503-515 */
                    d2 = d2 + cross;
                }
                if (d2 <= r2) cb(idx_i, acc->eidx[o + j], d2, user);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * RDF: compute_rdf / rdf_cb / rdf_cb_excl_mask / rdf_increment_bin (md_script_functions.inl:5221-5338)
 */
typedef struct rdf_payload_t {
    float min_cutoff, inv_cutoff_range; float* bins; int32_t num_bins; uint64_t total;
    const uint32_t* excl_off; const int32_t* excl_idx;
} rdf_payload_t;

static void rdf_pair(uint32_t i, uint32_t j, float d2, void* user) {
    rdf_payload_t* p = user;
    const float min_r2 = p->min_cutoff * p->min_cutoff;
    if (d2 < min_r2) return;
    if (p->excl_off) {  /* md_bitfield_test_bit(&exclusion_masks[i], j) */
        for (uint32_t k = p->excl_off[i]; k < p->excl_off[i + 1]; ++k) if ((uint32_t)p->excl_idx[k] == j) return;
    }
    const float d = sqrtf(d2);
    int32_t b = (int32_t)(((d - p->min_cutoff) * p->inv_cutoff_range) * p->num_bins);
    b = CLAMPV(b, 0, p->num_bins - 1);
    p->bins[b] += 1.0f;
    p->total += 1;
}

static double sphere_volume(double r) { return (4.0 / 3.0) * 3.1415926535897932 * (r * r * r); }

uint64_t mdo_rdf_frame(const float* x, const float* y, const float* z,
                       const int32_t* ref_idx, const float* ref_pos_aos, size_t n_ref,
                       const int32_t* trg_idx, size_t n_trg,
                       const mdo_unitcell_t* cell, float min_cutoff, float max_cutoff,
                       const uint32_t* excl_off, const int32_t* excl_idx,
                       float* bins, float* weights) {
    const int num_bins = MDO_DIST_BINS;
    const float inv_cutoff_range = 1.0f / (max_cutoff - min_cutoff);   /* before the clamp (:5264) */
    min_cutoff = MAXV(min_cutoff, 1e-3f);                               /* :5269 */
    memset(bins, 0, sizeof(float) * num_bins);
    stream_t ref = { x, y, z, ref_idx, ref_pos_aos, n_ref };
    stream_t trg = { x, y, z, trg_idx, NULL, n_trg };
    acc_t acc; acc_init(&acc, &trg, max_cutoff, cell, true);
    rdf_payload_t p = { min_cutoff, inv_cutoff_range, bins, num_bins, 0, excl_off, excl_idx };
    /* with exclusion masks i must be the structure index: COM references carry no idx, so i = ei (:1721) */
    acc_ext_pairs(&acc, &ref, max_cutoff, ref_pos_aos == NULL, rdf_pair, &p);
    acc_free(&acc);
    const double total_vol = sphere_volume(max_cutoff) - sphere_volume(min_cutoff);
    const double ref_rho = (double)p.total / total_vol;
    const double dr = (max_cutoff - min_cutoff) / (float)num_bins;     /* float expression widened (:5330) */
    double prev = 0;
    for (int64_t i = 0; i < num_bins; ++i) {
        const double sv = sphere_volume(min_cutoff + (i + 0.5) * dr);
        const double bv = sv - prev; prev = sv;
        weights[i] = (float)(ref_rho * bv);
    }
    return p.total;
}

static void count_pair(uint32_t i, uint32_t j, float d2, void* user) { (void)i; (void)j; (void)d2; *(uint64_t*)user += 1; }
uint64_t mdo_count_pairs(const float* x, const float* y, const float* z, const int32_t* ref_idx, size_t n_ref,
                         const int32_t* trg_idx, size_t n_trg, const mdo_unitcell_t* cell, double cell_ext, double cutoff) {
    stream_t ref = { x, y, z, ref_idx, NULL, n_ref }, trg = { x, y, z, trg_idx, NULL, n_trg };
    acc_t acc; acc_init(&acc, &trg, cell_ext, cell, true);
    uint64_t n = 0; acc_ext_pairs(&acc, &ref, cutoff, true, count_pair, &n);
    acc_free(&acc); return n;
}

/* ------------------------------------------------------------------------------------------------
 * 3x3 SVD: ext/svd3/svd3.c (McAdams et al. TR1690; CPU version by E. Jang), restated line for line in float.
 */
static inline float inv_sqrt(float v) { return 1.0f / sqrtf(v); }
static inline void cond_swap(bool c, float* X, float* Y) { float Z = *X; *X = c ? *Y : *X; *Y = c ? Z : *Y; }
static inline void cond_neg_swap(bool c, float* X, float* Y) { float Z = -*X; *X = c ? *Y : *X; *Y = c ? Z : *Y; }

static void approx_givens(float a11, float a12, float a22, float* ch, float* sh) {
    *ch = 2 * (a11 - a22);
    *sh = a12;
    /* gamma is a double literal in the reference: the comparison is carried out in double */
    bool b = 5.828427124746190097 * *sh * *sh < *ch * *ch;
    float w = inv_sqrt(*ch * *ch + *sh * *sh);
    *ch = b ? w * *ch : (float)0.923879532511286756;
    *sh = b ? w * *sh : (float)0.382683432365089771;
}

static void jacobi_conj(const int x, const int y, const int z, float S[3][3], float q[4]) {
    float ch, sh; approx_givens(S[0][0], S[1][0], S[1][1], &ch, &sh);
    float scale = ch * ch + sh * sh;
    float a = (ch * ch - sh * sh) / scale;
    float b = (2 * sh * ch) / scale;
    float _S[3][3];
    _S[0][0] = S[0][0]; _S[1][0] = S[1][0]; _S[1][1] = S[1][1]; _S[2][0] = S[2][0]; _S[2][1] = S[2][1]; _S[2][2] = S[2][2];
    S[0][0] = a * (a * _S[0][0] + b * _S[1][0]) + b * (a * _S[1][0] + b * _S[1][1]);
    S[1][0] = a * (-b * _S[0][0] + a * _S[1][0]) + b * (-b * _S[1][0] + a * _S[1][1]);
    S[1][1] = -b * (-b * _S[0][0] + a * _S[1][0]) + a * (-b * _S[1][0] + a * _S[1][1]);
    S[2][0] = a * _S[2][0] + b * _S[2][1];
    S[2][1] = -b * _S[2][0] + a * _S[2][1];
    S[2][2] = _S[2][2];
    float tmp[3] = { q[0] * sh, q[1] * sh, q[2] * sh };
    sh *= q[3];
    q[0] *= ch; q[1] *= ch; q[2] *= ch; q[3] *= ch;
    q[z] += sh; q[3] -= tmp[z]; q[x] += tmp[y]; q[y] -= tmp[x];
    _S[0][0] = S[1][1]; _S[1][0] = S[2][1]; _S[1][1] = S[2][2]; _S[2][0] = S[1][0]; _S[2][1] = S[2][0]; _S[2][2] = S[0][0];
    S[0][0] = _S[0][0]; S[1][0] = _S[1][0]; S[1][1] = _S[1][1]; S[2][0] = _S[2][0]; S[2][1] = _S[2][1]; S[2][2] = _S[2][2];
}

static inline float dist2(float a, float b, float c) { return a * a + b * b + c * c; }

static void qr_givens(float a1, float a2, float* ch, float* sh) {
    float epsilon = (float)1e-6;
    float rho = sqrtf(a1 * a1 + a2 * a2);
    *sh = rho > epsilon ? a2 : 0;
    *ch = fabsf(a1) + fmaxf(rho, epsilon);
    bool b = a1 < 0;
    cond_swap(b, sh, ch);
    float w = inv_sqrt(*ch * *ch + *sh * *sh);
    *ch *= w; *sh *= w;
}

void mdo_svd3(const float A[3][3], float U[3][3], float S[3][3], float V[3][3]) {
    float ATA[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) ATA[i][j] = A[0][i] * A[0][j] + A[1][i] * A[1][j] + A[2][i] * A[2][j];  /* multAtB */
    float q[4] = { 0, 0, 0, 1 };
    for (int i = 0; i < 4; ++i) { jacobi_conj(0, 1, 2, ATA, q); jacobi_conj(1, 2, 0, ATA, q); jacobi_conj(2, 0, 1, ATA, q); }
    {   /* quatToMat3 */
        float x = q[0], y = q[1], z = q[2], w = q[3];
        float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
        V[0][0] = 1 - 2 * (qyy + qzz); V[0][1] = 2 * (qxy - qwz);     V[0][2] = 2 * (qxz + qwy);
        V[1][0] = 2 * (qxy + qwz);     V[1][1] = 1 - 2 * (qxx + qzz); V[1][2] = 2 * (qyz - qwx);
        V[2][0] = 2 * (qxz - qwy);     V[2][1] = 2 * (qyz + qwx);     V[2][2] = 1 - 2 * (qxx + qyy);
    }
    float B[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[i][j] = A[i][0] * V[0][j] + A[i][1] * V[1][j] + A[i][2] * V[2][j];   /* multAB */
    {   /* sortSingularValues */
        float rho1 = dist2(B[0][0], B[1][0], B[2][0]), rho2 = dist2(B[0][1], B[1][1], B[2][1]), rho3 = dist2(B[0][2], B[1][2], B[2][2]);
        bool c = rho1 < rho2;
        for (int r = 0; r < 3; ++r) { cond_neg_swap(c, &B[r][0], &B[r][1]); cond_neg_swap(c, &V[r][0], &V[r][1]); }
        cond_swap(c, &rho1, &rho2);
        c = rho1 < rho3;
        for (int r = 0; r < 3; ++r) { cond_neg_swap(c, &B[r][0], &B[r][2]); cond_neg_swap(c, &V[r][0], &V[r][2]); }
        cond_swap(c, &rho1, &rho3);
        c = rho2 < rho3;
        for (int r = 0; r < 3; ++r) { cond_neg_swap(c, &B[r][1], &B[r][2]); cond_neg_swap(c, &V[r][1], &V[r][2]); }
    }
    {   /* QRDecomposition(B, Q=U, R=S) */
        float (*Q)[3] = U, (*R)[3] = S;
        float ch1, sh1, ch2, sh2, ch3, sh3, a, b;
        qr_givens(B[0][0], B[1][0], &ch1, &sh1);
        a = 1 - 2 * sh1 * sh1; b = 2 * ch1 * sh1;
        R[0][0] = a * B[0][0] + b * B[1][0];  R[0][1] = a * B[0][1] + b * B[1][1];  R[0][2] = a * B[0][2] + b * B[1][2];
        R[1][0] = -b * B[0][0] + a * B[1][0]; R[1][1] = -b * B[0][1] + a * B[1][1]; R[1][2] = -b * B[0][2] + a * B[1][2];
        R[2][0] = B[2][0]; R[2][1] = B[2][1]; R[2][2] = B[2][2];
        qr_givens(R[0][0], R[2][0], &ch2, &sh2);
        a = 1 - 2 * sh2 * sh2; b = 2 * ch2 * sh2;
        B[0][0] = a * R[0][0] + b * R[2][0];  B[0][1] = a * R[0][1] + b * R[2][1];  B[0][2] = a * R[0][2] + b * R[2][2];
        B[1][0] = R[1][0]; B[1][1] = R[1][1]; B[1][2] = R[1][2];
        B[2][0] = -b * R[0][0] + a * R[2][0]; B[2][1] = -b * R[0][1] + a * R[2][1]; B[2][2] = -b * R[0][2] + a * R[2][2];
        qr_givens(B[1][1], B[2][1], &ch3, &sh3);
        a = 1 - 2 * sh3 * sh3; b = 2 * ch3 * sh3;
        R[0][0] = B[0][0]; R[0][1] = B[0][1]; R[0][2] = B[0][2];
        R[1][0] = a * B[1][0] + b * B[2][0];  R[1][1] = a * B[1][1] + b * B[2][1];  R[1][2] = a * B[1][2] + b * B[2][2];
        R[2][0] = -b * B[1][0] + a * B[2][0]; R[2][1] = -b * B[1][1] + a * B[2][1]; R[2][2] = -b * B[1][2] + a * B[2][2];
        float sh12 = sh1 * sh1, sh22 = sh2 * sh2, sh32 = sh3 * sh3;
        Q[0][0] = (-1 + 2 * sh12) * (-1 + 2 * sh22);
        Q[0][1] = 4 * ch2 * ch3 * (-1 + 2 * sh12) * sh2 * sh3 + 2 * ch1 * sh1 * (-1 + 2 * sh32);
        Q[0][2] = 4 * ch1 * ch3 * sh1 * sh3 - 2 * ch2 * (-1 + 2 * sh12) * sh2 * (-1 + 2 * sh32);
        Q[1][0] = 2 * ch1 * sh1 * (1 - 2 * sh22);
        Q[1][1] = -8 * ch1 * ch2 * ch3 * sh1 * sh2 * sh3 + (-1 + 2 * sh12) * (-1 + 2 * sh32);
        Q[1][2] = -2 * ch3 * sh3 + 4 * sh1 * (ch3 * sh1 * sh3 + ch1 * ch2 * sh2 * (-1 + 2 * sh32));
        Q[2][0] = 2 * ch2 * sh2;
        Q[2][1] = 2 * ch3 * (1 - 2 * sh22) * sh3;
        Q[2][2] = (-1 + 2 * sh22) * (-1 + 2 * sh32);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Small matrix helpers with the reference's storage convention elem[col][row] (core/md_vec_math.h:107-121)
 */
typedef struct { float e[3][3]; } m3;
typedef struct { float e[4][4]; } m4;

static m3 m3_transpose(m3 M) { m3 T; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T.e[i][j] = M.e[j][i]; return T; }
static m3 m3_mul(m3 A, m3 B) {   /* core/md_vec_math.h:1631 */
    m3 C;
    for (int col = 0; col < 3; ++col) for (int row = 0; row < 3; ++row)
        C.e[col][row] = A.e[0][row] * B.e[col][0] + A.e[1][row] * B.e[col][1] + A.e[2][row] * B.e[col][2];
    return C;
}
static float m3_det(m3 M) {      /* :1687 */
    return M.e[0][0] * (M.e[1][1] * M.e[2][2] - M.e[2][1] * M.e[1][2])
         - M.e[1][0] * (M.e[0][1] * M.e[2][2] - M.e[2][1] * M.e[0][2])
         + M.e[2][0] * (M.e[0][1] * M.e[1][2] - M.e[1][1] * M.e[0][2]);
}
static m4 m4_from_m3(m3 M) { m4 R; memset(&R, 0, sizeof(R)); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.e[i][j] = M.e[i][j]; R.e[3][3] = 1; return R; }
static m4 m4_mul(m4 A, m4 B) {   /* linear_combine_4 (:1512): C.col[j] = ((B[j][0]*A0 + B[j][1]*A1) + B[j][2]*A2) + B[j][3]*A3 */
    m4 C;
    for (int j = 0; j < 4; ++j) for (int r = 0; r < 4; ++r) {
        float v = B.e[j][0] * A.e[0][r];
        v = v + B.e[j][1] * A.e[1][r];
        v = v + B.e[j][2] * A.e[2][r];
        v = v + B.e[j][3] * A.e[3][r];
        C.e[j][r] = v;
    }
    return C;
}
static void m4_mul_v(float out[4], const m4* M, const float v[4]) {
    for (int r = 0; r < 4; ++r) {
        float t = v[0] * M->e[0][r];
        t = t + v[1] * M->e[1][r];
        t = t + v[2] * M->e[2][r];
        t = t + v[3] * M->e[3][r];
        out[r] = t;
    }
}

typedef struct { m3 U, V; float s[3]; } svd_t;
static svd_t m3_svd(m3 M) {      /* core/md_vec_math.c:7-20 */
    m3 Mt = m3_transpose(M), U, S, V;
    mdo_svd3((const float(*)[3])Mt.e, U.e, S.e, V.e);
    svd_t r; r.U = m3_transpose(U); r.V = m3_transpose(V); r.s[0] = S.e[0][0]; r.s[1] = S.e[1][1]; r.s[2] = S.e[2][2];
    return r;
}
static m3 m3_eigen_vectors(m3 M) {   /* mat3_eigen core/md_vec_math.c:22-42 */
    svd_t s = m3_svd(M);
    const float mx = MAXV(s.s[0], MAXV(s.s[1], s.s[2]));
    const float ev[3] = { s.s[0] / mx, s.s[1] / mx, s.s[2] / mx };
    int l[3] = { 0, 1, 2 }, t;
    if (ev[l[0]] < ev[l[1]]) { t = l[0]; l[0] = l[1]; l[1] = t; }
    if (ev[l[1]] < ev[l[2]]) { t = l[1]; l[1] = l[2]; l[2] = t; }
    if (ev[l[0]] < ev[l[1]]) { t = l[0]; l[0] = l[1]; l[1] = t; }
    m3 R; for (int k = 0; k < 3; ++k) for (int r = 0; r < 3; ++r) R.e[k][r] = s.U.e[l[k]][r];
    return R;
}
static m3 m3_extract_rotation(m3 M) {   /* core/md_vec_math.c:292-300 */
    svd_t s = m3_svd(M);
    m3 Ut = m3_transpose(s.U);
    float d = m3_det(m3_mul(s.V, Ut));
    m3 D; memset(&D, 0, sizeof(D)); D.e[0][0] = 1; D.e[1][1] = 1; D.e[2][2] = (float)((d > 0.0f) - (d < 0.0f));
    return m3_mul(m3_mul(s.V, D), Ut);
}

/* vec4_deperiodize_ortho core/md_vec_math.h:1242-1253 (round = nearest-even, SSE4.1 roundps) */
static inline float deperiodize1(float x, float r, float ext) {
    if (ext == 0.0f) return x;
    const float inv = 1.0f / ext;
    float dx = (x - r) * inv;
    const float dxp = dx - rintf(dx);
    return r + dxp * ext;
}

/* ------------------------------------------------------------------------------------------------
 * SDF: _sdf (md_script_functions.inl:5699-5856)
 */
typedef float v4[4];

/* minimum_image_triclinic md_util.c:1677-1718: the 27 images, double comparison; note the mixed precision of the sums as written
 * there (float products, some float sums) */
static void min_image_triclinic(float dx[3], const float box[3][3]) {
    double dx_min[3] = { 0.0, 0.0, 0.0 }, dsq_min = FLT_MAX;
    for (int ix = -1; ix < 2; ++ix) {
        const double rx = (float)(dx[0] + box[0][0] * ix);
        for (int iy = -1; iy < 2; ++iy) {
            const double ry0 = rx + (float)(box[1][0] * iy);
            const double ry1 = (float)(dx[1] + box[1][1] * iy);
            for (int iz = -1; iz < 2; ++iz) {
                const double rz0 = ry0 + (float)(box[2][0] * iz), rz1 = ry1 + (float)(box[2][1] * iz), rz2 = (float)(dx[2] + box[2][2] * iz);
                const double dsq = rz0 * rz0 + rz1 * rz1 + rz2 * rz2;
                if (dsq < dsq_min) { dsq_min = dsq; dx_min[0] = rz0; dx_min[1] = rz1; dx_min[2] = rz2; }
            }
        }
    }
    dx[0] = (float)dx_min[0]; dx[1] = (float)dx_min[1]; dx[2] = (float)dx_min[2];
}

/* unwrap_topology_vec4 with indices == NULL (md_util.c:8738-8819): NB the BFS runs over the bonds of GLOBAL atoms
 * 0..count-1 (seed = local index used as a global atom index) — replicated as is. Ortho: unwrap_atom_ortho_vec4; triclinic:
 * deperiodize_triclinic :1754-1766 with the float basis (md_unitcell_A_extract_float). */
static void unwrap_vec4(v4* xyzw, size_t count, const uint32_t* conn_off, const int32_t* conn_idx, size_t conn_off_count, const mdo_unitcell_t* cell) {
    if (count == 0 || !conn_off || conn_off_count == 0) return;
    const int ortho = (cell->flags & MDO_CELL_ORTHO) != 0, tri = (cell->flags & MDO_CELL_TRICLINIC) != 0;
    if (!ortho && !tri) return;
    const float ext[3] = { (float)cell->x, (float)cell->y, (float)cell->z };
    const float box[3][3] = { { (float)cell->x, 0, 0 }, { (float)cell->xy, (float)cell->y, 0 }, { (float)cell->xz, (float)cell->yz, (float)cell->z } };
    const size_t atom_count = conn_off_count - 1;
    unsigned char* visited = calloc(atom_count + 1, 1);
    int* queue = malloc(sizeof(int) * (count + 1));
    for (size_t i = 0; i < count; ++i) {
        const int seed = (int)i;
        if ((size_t)seed >= atom_count || visited[seed]) continue;
        visited[seed] = 1;
        size_t qh = 0, qt = 0; queue[qt++] = seed;
        while (qh < qt) {
            const int cur = queue[qh++];
            for (uint32_t k = conn_off[cur]; k < conn_off[cur + 1]; ++k) {
                const int next = conn_idx[k];
                if ((size_t)next >= count) continue;
                if (visited[next]) continue;
                if (ortho) { for (int a = 0; a < 3; ++a) xyzw[next][a] = deperiodize1(xyzw[next][a], xyzw[cur][a], ext[a]); }
                else {
                    float d[3] = { xyzw[next][0] - xyzw[cur][0], xyzw[next][1] - xyzw[cur][1], xyzw[next][2] - xyzw[cur][2] };
                    min_image_triclinic(d, box);
                    for (int a = 0; a < 3; ++a) xyzw[next][a] = xyzw[cur][a] + d[a];
                }
                visited[next] = 1; queue[qt++] = next;
            }
        }
    }
    free(visited); free(queue);
}

/* com_vec4 md_util.c:8048-8061 (the unit_cell argument is 0 at both call sites in _sdf) */
static void com_v4(float com[3], const v4* p, size_t n) {
    float acc[4] = { 0, 0, 0, 0 };
    for (size_t i = 0; i < n; ++i) {
        const float w = p[i][3];
        acc[0] = acc[0] + p[i][0] * w; acc[1] = acc[1] + p[i][1] * w; acc[2] = acc[2] + p[i][2] * w; acc[3] = acc[3] + p[i][3] * 1.0f;
    }
    com[0] = acc[0] / acc[3]; com[1] = acc[1] / acc[3]; com[2] = acc[2] / acc[3];
}

/* mat3_covariance_matrix_vec4 core/md_vec_math.c:101-156 */
static m3 covariance_v4(const v4* p, size_t n, const float com[3]) {
    double A[3][3] = { { 0 } }; double ws = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const float x = p[i][0] - com[0], y = p[i][1] - com[1], z = p[i][2] - com[2], w = p[i][3];
        A[0][0] += w * x * x; A[0][1] += w * x * y; A[0][2] += w * x * z;
        A[1][0] += w * y * x; A[1][1] += w * y * y; A[1][2] += w * y * z;
        A[2][0] += w * z * x; A[2][1] += w * z * y; A[2][2] += w * z * z;
        ws += w;
    }
    m3 R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] /= ws; R.e[i][j] = (float)A[i][j]; }
    return R;
}
/* mat3_cross_covariance_matrix_vec4 core/md_vec_math.c:227-290 (in_idx == NULL) */
static m3 cross_covariance_v4(const v4* p0, const v4* p1, size_t n, const float com0[3], const float com1[3]) {
    double A[3][3] = { { 0 } }; double ws = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const float px = p0[i][0] - com0[0], py = p0[i][1] - com0[1], pz = p0[i][2] - com0[2], pw = p0[i][3] - 0.0f;
        const float qx = p1[i][0] - com1[0], qy = p1[i][1] - com1[1], qz = p1[i][2] - com1[2], qw = p1[i][3] - 0.0f;
        const float w = (pw + qw) * 0.5f;
        A[0][0] += w * px * qx; A[0][1] += w * px * qy; A[0][2] += w * px * qz;
        A[1][0] += w * py * qx; A[1][1] += w * py * qy; A[1][2] += w * py * qz;
        A[2][0] += w * pz * qx; A[2][1] += w * pz * qy; A[2][2] += w * pz * qz;
        ws += w;
    }
    m3 R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] /= ws; R.e[i][j] = (float)A[i][j]; }
    return R;
}

static inline int wrap_coord(int v, int N) { v += (v < 0) ? N : 0; v -= (v >= N) ? N : 0; return v; }
static inline int isign(int a) { return (a > 0) - (a < 0); }

typedef void (*point_cb_t)(uint32_t idx, float px, float py, float pz, void* user);

/* for_each_point_in_aabb_ortho + cell_range_from_aabb_center_radius (core/md_spatial_acc.c:1805-2007) */
static void acc_points_in_aabb_ortho(const acc_t* acc, const double cen[3], const double rad[3], point_cb_t cb, void* user) {
    if (acc->num_elems == 0) return;
    const int pbc[3] = { (acc->flags & MDO_CELL_PBC_X) != 0, (acc->flags & MDO_CELL_PBC_Y) != 0, (acc->flags & MDO_CELL_PBC_Z) != 0 };
    const int cd[3] = { (int)acc->cell_dim[0], (int)acc->cell_dim[1], (int)acc->cell_dim[2] };
    double sc[3], cc[3];
    {   /* cart_to_fract (double, float matrix entries widened) :556-565 */
        const double px = cen[0] - acc->origin[0], py = cen[1] - acc->origin[1], pz = cen[2] - acc->origin[2];
        sc[0] = acc->I[0][0] * px + acc->I[1][0] * py + acc->I[2][0] * pz;
        sc[1] = acc->I[0][1] * px + acc->I[1][1] * py + acc->I[2][1] * pz;
        sc[2] = acc->I[0][2] * px + acc->I[1][2] * py + acc->I[2][2] * pz;
    }
    for (int a = 0; a < 3; ++a) if (pbc[a]) sc[a] = sc[a] - floor(sc[a]);
    cc[0] = acc->A[0][0] * sc[0] + acc->A[1][0] * sc[1] + acc->A[2][0] * sc[2] + acc->origin[0];
    cc[1] = acc->A[0][1] * sc[0] + acc->A[1][1] * sc[1] + acc->A[2][1] * sc[2] + acc->origin[1];
    cc[2] = acc->A[0][2] * sc[0] + acc->A[1][2] * sc[1] + acc->A[2][2] * sc[2] + acc->origin[2];
    double fmin[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, fmax[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    for (int iz = 0; iz < 2; ++iz) { const double pz = cc[2] + (iz ? +rad[2] : -rad[2]);
    for (int iy = 0; iy < 2; ++iy) { const double py = cc[1] + (iy ? +rad[1] : -rad[1]);
    for (int ix = 0; ix < 2; ++ix) { const double px = cc[0] + (ix ? +rad[0] : -rad[0]);
        const double qx = px - acc->origin[0], qy = py - acc->origin[1], qz = pz - acc->origin[2];
        double s[3];
        s[0] = acc->I[0][0] * qx + acc->I[1][0] * qy + acc->I[2][0] * qz;
        s[1] = acc->I[0][1] * qx + acc->I[1][1] * qy + acc->I[2][1] * qz;
        s[2] = acc->I[0][2] * qx + acc->I[1][2] * qy + acc->I[2][2] * qz;
        for (int a = 0; a < 3; ++a) { fmin[a] = MINV(fmin[a], s[a]); fmax[a] = MAXV(fmax[a], s[a]); }
    } } }
    double frad[3]; int cmin[3], cmax[3];
    for (int a = 0; a < 3; ++a) {
        frad[a] = 0.5 * (fmax[a] - fmin[a]);
        int lo = (int)floor(fmin[a] * (double)cd[a]), hi = (int)ceil(fmax[a] * (double)cd[a]);
        if (hi <= lo) hi = lo + 1;
        if (!pbc[a]) { lo = CLAMPV(lo, 0, cd[a]); hi = CLAMPV(hi, 0, cd[a]); if (hi <= lo) hi = MINV(lo + 1, cd[a]); }
        cmin[a] = lo; cmax[a] = hi;
        frad[a] = MINV(frad[a], 0.5);                                      /* :1913-1915 */
    }
    const float lo3[3] = { (float)(sc[0] - frad[0]), (float)(sc[1] - frad[1]), (float)(sc[2] - frad[2]) };
    const float hi3[3] = { (float)(sc[0] + frad[0]), (float)(sc[1] + frad[1]), (float)(sc[2] + frad[2]) };
    const uint32_t c0 = acc->cell_dim[0], c01 = acc->cell_dim[0] * acc->cell_dim[1];
    for (int icz = cmin[2]; icz < cmax[2]; ++icz) { const int cz = pbc[2] ? wrap_coord(icz, cd[2]) : icz; const float shz = (float)isign(icz - cz);
    for (int icy = cmin[1]; icy < cmax[1]; ++icy) { const int cy = pbc[1] ? wrap_coord(icy, cd[1]) : icy; const float shy = (float)isign(icy - cy);
    for (int icx = cmin[0]; icx < cmax[0]; ++icx) { const int cx = pbc[0] ? wrap_coord(icx, cd[0]) : icx; const float shx = (float)isign(icx - cx);
        if (cx < 0 || cx >= cd[0] || cy < 0 || cy >= cd[1] || cz < 0 || cz >= cd[2]) continue;   /* single wrap only in the reference */
        const size_t ci = (size_t)cz * c01 + (size_t)cy * c0 + (size_t)cx;
        const uint32_t o = acc->cell_off[ci], len = acc->cell_off[ci + 1] - o;
        for (uint32_t j = 0; j < len; ++j) {
            const float vx = acc->ex[o + j] + shx, vy = acc->ey[o + j] + shy, vz = acc->ez[o + j] + shz;
            if (vx >= lo3[0] && vy >= lo3[1] && vz >= lo3[2] && vx <= hi3[0] && vy <= hi3[1] && vz <= hi3[2]) {
                /* fract_to_cart_ort_256: single fused multiply-add per axis (:583-592) */
                cb(acc->eidx[o + j], fmaf(vx, acc->A[0][0], acc->origin[0]), fmaf(vy, acc->A[1][1], acc->origin[1]), fmaf(vz, acc->A[2][2], acc->origin[2]), user);
            }
        }
    } } }
}

/* for_each_point_in_aabb_triclinic (core/md_spatial_acc.c:2009-2150): same cell range, the box test in CARTESIAN coordinates of the
 * image-shifted point (fract_to_cart_tri_256 :594-603), all axes periodic */
static void acc_points_in_aabb_triclinic(const acc_t* acc, const double cen[3], const double rad[3], point_cb_t cb, void* user) {
    if (acc->num_elems == 0) return;
    const int cd[3] = { (int)acc->cell_dim[0], (int)acc->cell_dim[1], (int)acc->cell_dim[2] };
    double sc[3], cc[3];
    {
        const double px = cen[0] - acc->origin[0], py = cen[1] - acc->origin[1], pz = cen[2] - acc->origin[2];
        sc[0] = acc->I[0][0] * px + acc->I[1][0] * py + acc->I[2][0] * pz;
        sc[1] = acc->I[0][1] * px + acc->I[1][1] * py + acc->I[2][1] * pz;
        sc[2] = acc->I[0][2] * px + acc->I[1][2] * py + acc->I[2][2] * pz;
    }
    const int pbc[3] = { (acc->flags & MDO_CELL_PBC_X) != 0, (acc->flags & MDO_CELL_PBC_Y) != 0, (acc->flags & MDO_CELL_PBC_Z) != 0 };
    for (int a = 0; a < 3; ++a) if (pbc[a]) sc[a] = sc[a] - floor(sc[a]);
    cc[0] = acc->A[0][0] * sc[0] + acc->A[1][0] * sc[1] + acc->A[2][0] * sc[2] + acc->origin[0];
    cc[1] = acc->A[0][1] * sc[0] + acc->A[1][1] * sc[1] + acc->A[2][1] * sc[2] + acc->origin[1];
    cc[2] = acc->A[0][2] * sc[0] + acc->A[1][2] * sc[1] + acc->A[2][2] * sc[2] + acc->origin[2];
    double fmin[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, fmax[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    for (int iz = 0; iz < 2; ++iz) { const double pz = cc[2] + (iz ? +rad[2] : -rad[2]);
    for (int iy = 0; iy < 2; ++iy) { const double py = cc[1] + (iy ? +rad[1] : -rad[1]);
    for (int ix = 0; ix < 2; ++ix) { const double px = cc[0] + (ix ? +rad[0] : -rad[0]);
        const double qx = px - acc->origin[0], qy = py - acc->origin[1], qz = pz - acc->origin[2];
        double s[3];
        s[0] = acc->I[0][0] * qx + acc->I[1][0] * qy + acc->I[2][0] * qz;
        s[1] = acc->I[0][1] * qx + acc->I[1][1] * qy + acc->I[2][1] * qz;
        s[2] = acc->I[0][2] * qx + acc->I[1][2] * qy + acc->I[2][2] * qz;
        for (int a = 0; a < 3; ++a) { fmin[a] = MINV(fmin[a], s[a]); fmax[a] = MAXV(fmax[a], s[a]); }
    } } }
    int cmin[3], cmax[3];
    for (int a = 0; a < 3; ++a) {
        int lo = (int)floor(fmin[a] * (double)cd[a]), hi = (int)ceil(fmax[a] * (double)cd[a]);
        if (hi <= lo) hi = lo + 1;
        if (!pbc[a]) { lo = CLAMPV(lo, 0, cd[a]); hi = CLAMPV(hi, 0, cd[a]); if (hi <= lo) hi = MINV(lo + 1, cd[a]); }
        cmin[a] = lo; cmax[a] = hi;
    }
    const float lo3[3] = { (float)(cc[0] - rad[0]), (float)(cc[1] - rad[1]), (float)(cc[2] - rad[2]) };   /* :2041-2047 */
    const float hi3[3] = { (float)(cc[0] + rad[0]), (float)(cc[1] + rad[1]), (float)(cc[2] + rad[2]) };
    const float A00 = acc->A[0][0], A10 = acc->A[1][0], A11 = acc->A[1][1], A20 = acc->A[2][0], A21 = acc->A[2][1], A22 = acc->A[2][2];
    const float O0 = acc->origin[0], O1 = acc->origin[1], O2 = acc->origin[2];
    const uint32_t c0 = acc->cell_dim[0], c01 = acc->cell_dim[0] * acc->cell_dim[1];
    for (int icz = cmin[2]; icz < cmax[2]; ++icz) { const int cz = wrap_coord(icz, cd[2]); const float shz = (float)isign(icz - cz);
    for (int icy = cmin[1]; icy < cmax[1]; ++icy) { const int cy = wrap_coord(icy, cd[1]); const float shy = (float)isign(icy - cy);
    for (int icx = cmin[0]; icx < cmax[0]; ++icx) { const int cx = wrap_coord(icx, cd[0]); const float shx = (float)isign(icx - cx);
        if (cx < 0 || cx >= cd[0] || cy < 0 || cy >= cd[1] || cz < 0 || cz >= cd[2]) continue;   /* single wrap only; the reference would index out of range */
        const size_t ci = (size_t)cz * c01 + (size_t)cy * c0 + (size_t)cx;
        const uint32_t o = acc->cell_off[ci], len = acc->cell_off[ci + 1] - o;
        for (uint32_t j = 0; j < len; ++j) {
            const float vx = acc->ex[o + j] + shx, vy = acc->ey[o + j] + shy, vz = acc->ez[o + j] + shz;
            const float px = fmaf(vx, A00, fmaf(vy, A10, fmaf(vz, A20, O0))), py = fmaf(vy, A11, fmaf(vz, A21, O1)), pz = fmaf(vz, A22, O2);
            if (px >= lo3[0] && py >= lo3[1] && pz >= lo3[2] && px <= hi3[0] && py <= hi3[1] && pz <= hi3[2])
                cb(acc->eidx[o + j], vx, vy, vz, user);   /* REFERENCE QUIRK: the buffer holds the FRACTIONAL image-shifted coordinates (:2122-2130) and
                                                            * POSSIBLY_INVOKE_CALLBACK_POINT_CART_TRI / FLUSH_TAIL_POINT_CART_TRI (:715-737) hand it to the callback without
                                                            * the fract->cart conversion the ortho path does — the callback sees fractional numbers. Replicated. */
        }
    } } }
}

typedef struct sdf_payload_t { m4 M; float* vol; const int32_t* excl; size_t n_excl; uint64_t count; } sdf_payload_t;
static void sdf_point(uint32_t idx, float px, float py, float pz, void* user) {   /* sdf_cb :5664-5697 */
    sdf_payload_t* p = user;
    for (size_t k = 0; k < p->n_excl; ++k) if ((uint32_t)p->excl[k] == idx) return;
    const float v[4] = { px, py, pz, 1.0f }; float c[4];
    m4_mul_v(c, &p->M, v);
    const uint32_t ix = (uint32_t)CLAMPV((int32_t)c[0], 0, MDO_VOL_DIM - 1);
    const uint32_t iy = (uint32_t)CLAMPV((int32_t)c[1], 0, MDO_VOL_DIM - 1);
    const uint32_t iz = (uint32_t)CLAMPV((int32_t)c[2], 0, MDO_VOL_DIM - 1);
    p->vol[(size_t)iz * (MDO_VOL_DIM * MDO_VOL_DIM) + iy * MDO_VOL_DIM + ix] += 1.0f;
    p->count += 1;
}

uint64_t mdo_sdf_frame(const float* x, const float* y, const float* z,
                       const float* init_x, const float* init_y, const float* init_z, const float* mass,
                       const int32_t* struct_idx, size_t n_struct, size_t struct_size,
                       const int32_t* trg_idx, size_t n_trg,
                       const uint32_t* conn_off, const int32_t* conn_idx, size_t conn_off_count,
                       const mdo_unitcell_t* cell, float cutoff, float* vol, float* out_matrices) {
    if (n_struct == 0 || struct_size == 0 || n_trg == 0) return 0;
    v4* r0 = malloc(sizeof(v4) * struct_size); v4* r1 = malloc(sizeof(v4) * struct_size);
    float com0[3], com1[3];
    for (size_t k = 0; k < struct_size; ++k) {   /* extract_xyzw_vec4 from the INITIAL frame, structure 0 (:5762) */
        const int32_t a = struct_idx[k];
        r0[k][0] = init_x[a]; r0[k][1] = init_y[a]; r0[k][2] = init_z[a]; r0[k][3] = mass ? mass[a] : 1.0f;
    }
    unwrap_vec4(r0, struct_size, conn_off, conn_idx, conn_off_count, cell);   /* with the CURRENT frame's cell (:5768) */
    com_v4(com0, r0, struct_size);
    stream_t trg = { x, y, z, trg_idx, NULL, n_trg };
    acc_t acc; acc_init(&acc, &trg, (double)cutoff, cell, true);
    m3 eig = m3_eigen_vectors(covariance_v4(r0, struct_size, com0));
    m4 A = m4_from_m3(m3_transpose(eig));
    m4 V;   /* compute_volume_matrix :5643-5655 */
    {
        const float voxel_ext = (2 * cutoff) / MDO_VOL_DIM;
        const float s = 1.0f / voxel_ext;
        m4 S; memset(&S, 0, sizeof(S)); S.e[0][0] = s; S.e[1][1] = s; S.e[2][2] = s; S.e[3][3] = 1;
        const float t = (MDO_VOL_DIM / 2);
        m4 T; memset(&T, 0, sizeof(T)); T.e[0][0] = T.e[1][1] = T.e[2][2] = T.e[3][3] = 1; T.e[3][0] = t; T.e[3][1] = t; T.e[3][2] = t;
        V = m4_mul(T, S);
    }
    m4 VA = m4_mul(V, A);
    uint64_t total = 0;
    for (size_t i = 0; i < n_struct; ++i) {
        const int32_t* sidx = struct_idx + i * struct_size;
        for (size_t k = 0; k < struct_size; ++k) { const int32_t a = sidx[k]; r1[k][0] = x[a]; r1[k][1] = y[a]; r1[k][2] = z[a]; r1[k][3] = mass ? mass[a] : 1.0f; }
        unwrap_vec4(r1, struct_size, conn_off, conn_idx, conn_off_count, cell);
        com_v4(com1, r1, struct_size);
        m3 R = m3_extract_rotation(cross_covariance_v4(r0, r1, struct_size, com0, com1));
        m4 Tm; memset(&Tm, 0, sizeof(Tm)); Tm.e[0][0] = Tm.e[1][1] = Tm.e[2][2] = Tm.e[3][3] = 1; Tm.e[3][0] = -com1[0]; Tm.e[3][1] = -com1[1]; Tm.e[3][2] = -com1[2];
        m4 RT = m4_mul(m4_from_m3(R), Tm);
        sdf_payload_t p; p.M = m4_mul(VA, RT); p.vol = vol; p.excl = sidx; p.n_excl = struct_size; p.count = 0;
        if (out_matrices) memcpy(out_matrices + 16 * i, p.M.e, sizeof(float) * 16);
        const double cen[3] = { com1[0], com1[1], com1[2] }, rad[3] = { cutoff, cutoff, cutoff };
        if (vol) {
            if (acc.flags & MDO_CELL_TRICLINIC) acc_points_in_aabb_triclinic(&acc, cen, rad, sdf_point, &p);
            else acc_points_in_aabb_ortho(&acc, cen, rad, sdf_point, &p);
        }
        total += p.count;
    }
    acc_free(&acc); free(r0); free(r1);
    return total;
}

/* ------------------------------------------------------------------------------------------------
 * Cell-grid geometry of md_spatial_acc_init (core/md_spatial_acc.c:155-438) + the neighbour reach of the pair query (:1656-1662), exported so
 * that the product's host-side geometry code (mdgpu_debug_frame_geom, the same code its device kernels run) can be checked without a GPU.
 * out_i[7] = cell_dim[3], ncell[3], num_cells; out_f[10] = G00, G11, G22, H01, H02, H12, r2, origin[3].
 */
void mdo_debug_geom(const float* x, const float* y, const float* z, size_t n, const mdo_unitcell_t* cell, double cell_ext, double cutoff,
                    int32_t* out_i, float* out_f) {
    stream_t st = { x, y, z, NULL, NULL, n };
    acc_t acc; acc_init(&acc, &st, cell_ext, cell, false);
    for (int k = 0; k < 3; ++k) { out_i[k] = (int32_t)acc.cell_dim[k]; out_i[3 + k] = (int32_t)ceil(cutoff * (double)acc.inv_cell_ext[k] * acc.cell_dim[k]); }
    out_i[6] = (int32_t)acc.num_cells;
    out_f[0] = acc.G00; out_f[1] = acc.G11; out_f[2] = acc.G22; out_f[3] = acc.H01; out_f[4] = acc.H02; out_f[5] = acc.H12; out_f[6] = calc_r2(cutoff);
    for (int k = 0; k < 3; ++k) out_f[7 + k] = acc.origin[k];
    acc_free(&acc);
}

/* ------------------------------------------------------------------------------------------------
 * within(radius, selection): _within_expl_flt md_script_functions.inl:2485-2533 — every atom of the system within `radius` of any atom of
 * the selection, the selection's own atoms removed (:2521-2525). The system-wide cell list comes from get_spatial_acc (:734-753): cell
 * extent ceil(radius / 6) * 6; the positions are an AoS stream (coordinate_extract of one bitfield). out_mask: one byte per atom.
 * Returns the number of atoms set. Groundwork for dynamic selections: the GPU path does not lower them yet.
 */
typedef struct { uint8_t* mask; float min_r2; } within_user_t;
static void within_pair(uint32_t i, uint32_t j, float d2, void* user) {   /* within_float_cb :2478, within_frng_cb :2599 (d2 >= min_r2) */
    (void)i; const within_user_t* u = user; if (d2 >= u->min_r2) u->mask[j] = 1;
}
size_t mdo_within_range(const float* x, const float* y, const float* z, size_t num_atoms, const int32_t* sel, size_t n_sel, float rmin, float radius,
                        const mdo_unitcell_t* cell, uint8_t* out_mask);
size_t mdo_within(const float* x, const float* y, const float* z, size_t num_atoms, const int32_t* sel, size_t n_sel, float radius,
                  const mdo_unitcell_t* cell, uint8_t* out_mask) {
    return mdo_within_range(x, y, z, num_atoms, sel, n_sel, 0.0f, radius, cell, out_mask);   /* every d2 is >= 0 */
}
/* within(min:max, selection): _within_expl_frng :2609-2661 — the grid and the query use max, a pair counts when d2 >= min * min (float) */
size_t mdo_within_range(const float* x, const float* y, const float* z, size_t num_atoms, const int32_t* sel, size_t n_sel, float rmin, float radius,
                        const mdo_unitcell_t* cell, uint8_t* out_mask) {
    memset(out_mask, 0, num_atoms);
    if (n_sel == 0 || num_atoms == 0) return 0;
    const double cell_ext = ceil((double)radius / 6.0) * 6.0;
    stream_t all = { x, y, z, NULL, NULL, num_atoms };
    acc_t acc; acc_init(&acc, &all, cell_ext, cell, false);
    float* pos = malloc(sizeof(float) * 3 * n_sel);
    for (size_t k = 0; k < n_sel; ++k) { pos[3 * k] = x[sel[k]]; pos[3 * k + 1] = y[sel[k]]; pos[3 * k + 2] = z[sel[k]]; }
    stream_t ext = { x, y, z, NULL, pos, n_sel };
    within_user_t user = { out_mask, rmin * rmin };
    acc_ext_pairs(&acc, &ext, (double)radius, false, within_pair, &user);
    for (size_t k = 0; k < n_sel; ++k) out_mask[sel[k]] = 0;
    size_t n = 0; for (size_t a = 0; a < num_atoms; ++a) n += out_mask[a];
    acc_free(&acc); free(pos);
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * rmsd(selection): _rmsd md_script_functions.inl:4287-4345. Both the INITIAL frame's and the current frame's atoms of the (flattened)
 * selection are wrapped into the cell about its centre (md_util_pbc_vec4 -> pbc_ortho_vec4 md_util.c:8506-8512), unwrapped along the bonds
 * (the same global-index quirk as in _sdf), centred on their plain centres of mass, optimally rotated (Kabsch through svd3) and compared:
 * sqrt(sum w |u - R v|^2 / sum w) in double (md_util_rmsd_compute_vec4 :9038-9068). The wrap is md_util_pbc_vec4 (md_util.c:8603): about the box
 * centre for orthorhombic cells, A * fract(I * r) for triclinic ones.
 */
double mdo_rmsd_frame(const float* x, const float* y, const float* z, const float* init_x, const float* init_y, const float* init_z,
                      const float* mass, const int32_t* idx, size_t n, const uint32_t* conn_off, const int32_t* conn_idx, size_t conn_off_count,
                      const mdo_unitcell_t* cell) {
    if (n == 0) return 0.0;
    v4* p[2] = { malloc(sizeof(v4) * n), malloc(sizeof(v4) * n) };
    for (size_t k = 0; k < n; ++k) {
        const int32_t a = idx[k]; const float w = mass ? mass[a] : 1.0f;
        p[0][k][0] = init_x[a]; p[0][k][1] = init_y[a]; p[0][k][2] = init_z[a]; p[0][k][3] = w;
        p[1][k][0] = x[a];      p[1][k][1] = y[a];      p[1][k][2] = z[a];      p[1][k][3] = w;
    }
    float com[2][3];
    for (int s = 0; s < 2; ++s) {
        if (cell->flags & MDO_CELL_ORTHO) {   /* pbc_ortho_vec4: deperiodize about ext * 0.5 */
            const float ext[3] = { (float)cell->x, (float)cell->y, (float)cell->z };
            for (size_t k = 0; k < n; ++k) for (int a = 0; a < 3; ++a) p[s][k][a] = deperiodize1(p[s][k][a], ext[a] * 0.5f, ext[a]);
        } else if (cell->flags & MDO_CELL_TRICLINIC) {   /* pbc_triclinic_vec4 md_util.c:8554-8574: c = A * fract(I * c) on the periodic axes */
            double Ad[3][3], Id[3][3]; cell_A(Ad, cell); cell_I(Id, cell);
            float A[3][3], I[3][3];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] = (float)Ad[i][j]; I[i][j] = (float)Id[i][j]; }
            const int pbc[3] = { (cell->flags & MDO_CELL_PBC_X) != 0, (cell->flags & MDO_CELL_PBC_Y) != 0, (cell->flags & MDO_CELL_PBC_Z) != 0 };
            for (size_t k = 0; k < n; ++k) {
                const float c[3] = { p[s][k][0], p[s][k][1], p[s][k][2] };
                float f[3], r[3];   /* linear_combine_3 core/md_vec_math.h:1521: (x*col0 + y*col1) + z*col2 */
                for (int a = 0; a < 3; ++a) { f[a] = (c[0] * I[0][a] + c[1] * I[1][a]) + c[2] * I[2][a]; f[a] = f[a] - floorf(f[a]); }
                for (int a = 0; a < 3; ++a) r[a] = (f[0] * A[0][a] + f[1] * A[1][a]) + f[2] * A[2][a];
                for (int a = 0; a < 3; ++a) if (pbc[a]) p[s][k][a] = r[a];
            }
        }
        unwrap_vec4(p[s], n, conn_off, conn_idx, conn_off_count, cell);
        com_v4(com[s], p[s], n);
    }
    const m3 R = m3_extract_rotation(cross_covariance_v4(p[0], p[1], n, com[0], com[1]));   /* mat3_optimal_rotation_vec4 core/md_vec_math.c:337 */
    double d_sum = 0.0, w_sum = 0.0;
    for (size_t k = 0; k < n; ++k) {
        const float u[3] = { p[0][k][0] - com[0][0], p[0][k][1] - com[0][1], p[0][k][2] - com[0][2] };
        const float v[3] = { p[1][k][0] - com[1][0], p[1][k][1] - com[1][1], p[1][k][2] - com[1][2] };
        float vp[3];   /* mat3_mul_vec3 core/md_vec_math.h:1623: col-major, (x*c0 + y*c1) + z*c2 */
        for (int r = 0; r < 3; ++r) vp[r] = R.e[0][r] * v[0] + R.e[1][r] * v[1] + R.e[2][r] * v[2];
        const float d[3] = { u[0] - vp[0], u[1] - vp[1], u[2] - vp[2] };
        const float w = (p[0][k][3] + p[1][k][3]) * 0.5f;
        const float dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        d_sum += w * dd; w_sum += w;
    }
    free(p[0]); free(p[1]);
    return sqrt(d_sum / w_sum);
}

/* ------------------------------------------------------------------------------------------------
 * plane(selection): _plane md_script_functions.inl:4755-4822. The atoms' positions with unit weights, made whole along the bonds (same
 * local-index quirk), plain centre, covariance, eigenvectors sorted by eigenvalue (mat3_eigen); out = (normalised third axis, normal . com).
 */
static void normalize3(float v[3]);   /* vec3_normalize, defined with the temporals below */
void mdo_plane_frame(const float* x, const float* y, const float* z, const int32_t* idx, size_t n, const uint32_t* conn_off, const int32_t* conn_idx,
                     size_t conn_off_count, const mdo_unitcell_t* cell, float out[4]) {
    v4* p = malloc(sizeof(v4) * (n ? n : 1));
    for (size_t k = 0; k < n; ++k) { const int32_t a = idx[k]; p[k][0] = x[a]; p[k][1] = y[a]; p[k][2] = z[a]; p[k][3] = 1.0f; }
    unwrap_vec4(p, n, conn_off, conn_idx, conn_off_count, cell);
    float com[3]; com_v4(com, p, n);
    const m3 E = m3_eigen_vectors(covariance_v4(p, n, com));
    float nrm[3] = { E.e[2][0], E.e[2][1], E.e[2][2] };
    normalize3(nrm);
    out[0] = nrm[0]; out[1] = nrm[1]; out[2] = nrm[2]; out[3] = nrm[0] * com[0] + nrm[1] * com[1] + nrm[2] * com[2];
    free(p);
}

/* ------------------------------------------------------------------------------------------------
 * density_x/_y/_z: _internal_density (md_script_functions.inl:4825-4947), axis in 0..2
 */
void mdo_density_frame(const float* x, const float* y, const float* z, const float* mass,
                       const int32_t* idx, size_t n, const mdo_unitcell_t* init_cell, int axis, float* bins, float* weights) {
    for (int i = 0; i < MDO_DIST_BINS; ++i) { weights[i] = 1.0f; bins[i] = 0.0f; }
    float Af[3][3]; { double A[3][3]; cell_A(A, init_cell); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Af[i][j] = (float)A[i][j]; }
    /* rc = A * (0.5,0.5,0.5) (mat3_mul_vec3 :1623), re = diag(A) */
    float rc[3], re[3];
    for (int r = 0; r < 3; ++r) rc[r] = Af[0][r] * 0.5f + Af[1][r] * 0.5f + Af[2][r] * 0.5f;
    for (int r = 0; r < 3; ++r) re[r] = Af[r][r];
    float inv_ext[3], min_point[3];
    for (int r = 0; r < 3; ++r) { inv_ext[r] = re[r] > 0.0f ? 1.0f / re[r] : 0.0f; min_point[r] = rc[r] - re[r] * 0.5f; }
    const float* src[3] = { x, y, z };
    for (size_t i = 0; i < n; ++i) {
        const int32_t a = idx[i];
        const float v = deperiodize1(src[axis][a], rc[axis], re[axis]);
        const float fc = (v - min_point[axis]) * inv_ext[axis];
        const int b = CLAMPV((int)(fc * MDO_DIST_BINS), 0, MDO_DIST_BINS - 1);
        bins[b] += mass[a];
    }
    const double slice_vol = (re[0] * re[1] * re[2]) / MDO_DIST_BINS;   /* float product widened (:4930) */
    const double factor = 1660.5390666 / slice_vol;
    for (int i = 0; i < MDO_DIST_BINS; ++i) bins[i] = (float)(bins[i] * factor);
}

/* ------------------------------------------------------------------------------------------------
 * Argument positions of distance / angle / dihedral: coordinate_extract_com (md_script_functions.inl:1717-1850).
 * A single integer index is the atom's position (:1755); a selection (bitfield) or several indices go through
 * md_util_com_compute (md_util.c:8163): no cell -> com() :7139, otherwise the trigonometric periodic centre of mass
 * com_pbc :8019 -> _com_pbc_iw :7850. Both are restated for the AVX2 build (8 float lanes + double remainder), which is what
 * oracle/_ref/ref_harness_strict is compiled as.
 */
static float f_from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t bits_from_f(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* one lane of md_mm256_sincos_ps (core/md_simd.h:1177-1258; the Cody-Waite variant at :1003 is compiled out by #if 0 :920) */
static void ref_sincosf_dp3(float x, float dp3, float* out_s, float* out_c);
static void ref_sincosf(float x, float* out_s, float* out_c) { ref_sincosf_dp3(x, -3.77489470793079817668E-8f, out_s, out_c); }
/* the 4-lane md_mm_sincos_ps (core/md_simd.h:1093-1176) is the same sequence with a differently rounded third Cody-Waite constant (:1137) */
static void ref_sincosf4(float x, float* out_s, float* out_c) { ref_sincosf_dp3(x, -3.77489497744594108e-8f, out_s, out_c); }
static void ref_sincosf_dp3(float x, float dp3, float* out_s, float* out_c) {
    uint32_t sign_bit_sin = bits_from_f(x) & 0x80000000u;
    x = fabsf(x);
    float y = x * 1.27323954473516f;
    int32_t imm2 = (int32_t)y;                       /* cvttps */
    imm2 = (imm2 + 1) & ~1;
    y = (float)imm2;
    int32_t imm3 = imm2;
    const uint32_t swap_sign_bit_sin = ((uint32_t)(imm2 & 4)) << 29;
    const int poly_mask = ((imm2 & 2) == 0);
    imm3 = imm3 - 2;
    const uint32_t sign_bit_cos = ((uint32_t)(~imm3 & 4)) << 29;
    sign_bit_sin ^= swap_sign_bit_sin;
    x = fmaf(y, -0.78515625f, x);
    x = fmaf(y, -2.4187564849853515625E-4f, x);
    x = fmaf(y, dp3, x);
    const float x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
    y = fmaf(x2, fmaf(x2, 2.443315711809948E-5f, -1.388731625493765E-3f), 4.166664568298827E-2f);
    y = fmaf(x2, -0.5f, y * x4);
    y = y + 1.0f;
    float y2 = fmaf(x2, fmaf(x2, -1.9515295891E-4f, 8.3321608736E-3f), -1.6666654611E-1f);
    y2 = fmaf(y2, x3, x);
    const float ysin2 = poly_mask ? y2 : 0.0f, ysin1 = poly_mask ? 0.0f : y;
    y2 = y2 - ysin2; y = y - ysin1;
    const float xmm1 = ysin1 + ysin2, xmm2 = y + y2;
    *out_s = f_from_bits(bits_from_f(xmm1) ^ sign_bit_sin);
    *out_c = f_from_bits(bits_from_f(xmm2) ^ sign_bit_cos);
}

/* md_mm256_reduce_add_ps (core/md_simd.h:691) over md_mm_reduce_add_ps (:678): ((l0+l4)+(l1+l5)) + ((l2+l6)+(l3+l7)) */
static float reduce8(const float v[8]) {
    const float a0 = v[0] + v[4], a1 = v[1] + v[5], a2 = v[2] + v[6], a3 = v[3] + v[7];
    return (a0 + a1) + (a2 + a3);
}

void mdo_com(const float* x, const float* y, const float* z, const float* mass, const int32_t* idx, size_t count,
             const mdo_unitcell_t* cell, float out[3]) {
    out[0] = out[1] = out[2] = 0.0f;
    if (count == 0) return;
    const size_t simd_count = count & ~(size_t)7;
    size_t i = 0;
    if (!cell || cell->flags == 0) {   /* com() md_util.c:7139-7380, indices + weights branch */
        float vx[8] = {0}, vy[8] = {0}, vz[8] = {0}, vw[8] = {0};
        for (; i < simd_count; i += 8) for (int l = 0; l < 8; ++l) {
            const int32_t a = idx[i + l]; const float w = mass[a];
            volatile float px = x[a] * w, py = y[a] * w, pz = z[a] * w;
            vx[l] = vx[l] + px; vy[l] = vy[l] + py; vz[l] = vz[l] + pz; vw[l] = vw[l] + w;
        }
        double ax = reduce8(vx), ay = reduce8(vy), az = reduce8(vz), aw = reduce8(vw);
        for (; i < count; ++i) {
            const int32_t a = idx[i]; const float w = mass[a];
            volatile float px = x[a] * w, py = y[a] * w, pz = z[a] * w;
            ax += px; ay += py; az += pz; aw += w;
        }
        out[0] = (float)(ax / aw); out[1] = (float)(ay / aw); out[2] = (float)(az / aw);
        return;
    }
    /* com_pbc :8019-8046: M = scale(2pi) * Ai, I = A * scale(1/2pi), float matrices (mat3_mul core/md_vec_math.h:1631) */
    float A[3][3] = { { (float)cell->x, 0, 0 }, { (float)cell->xy, (float)cell->y, 0 }, { (float)cell->xz, (float)cell->yz, (float)cell->z } };
    float Ai[3][3];
    {   /* md_unitcell_I_extract_double md_unitcell.inl:158-176 */
        const double cx = cell->x, cy = cell->y, cz = cell->z;
        const double i11 = cx > 0.0 ? 1.0 / cx : 0.0, i22 = cy > 0.0 ? 1.0 / cy : 0.0, i33 = cz > 0.0 ? 1.0 / cz : 0.0;
        const double i12 = (cx * cy) > 0.0 ? -cell->xy / (cx * cy) : 0.0;
        const double i13 = (cx * cy * cz) > 0.0 ? (cell->xy * cell->yz - cell->xz * cy) / (cx * cy * cz) : 0.0;
        const double i23 = (cy * cz) > 0.0 ? -cell->yz / (cy * cz) : 0.0;
        const double I[3][3] = { { i11, 0, 0 }, { i12, i22, 0 }, { i13, i23, i33 } };
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ai[r][c] = (float)I[r][c];
    }
    const float tp = (float)6.283185307179586, itp = (float)(1.0 / 6.283185307179586);
    float M[3][3], I[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { M[r][c] = tp * Ai[r][c]; I[r][c] = A[r][c] * itp; }
    float vs[3][8] = {{0}}, vc[3][8] = {{0}}, vw[8] = {0};
    const float* src[3] = { x, y, z };
    for (; i < simd_count; i += 8) for (int l = 0; l < 8; ++l) {
        const int32_t a = idx[i + l]; const float w = mass[a];
        for (int k = 0; k < 3; ++k) {
            const float p = src[k][a];
            const float t = fmaf(p, M[k][0], fmaf(p, M[k][1], p * M[k][2]));   /* :7917-7919 */
            float sn, cs; ref_sincosf(t, &sn, &cs);
            vs[k][l] = fmaf(sn, w, vs[k][l]); vc[k][l] = fmaf(cs, w, vc[k][l]);
        }
        vw[l] = vw[l] + w;
    }
    double acc_s[3], acc_c[3], acc_w = reduce8(vw);
    for (int k = 0; k < 3; ++k) { acc_s[k] = reduce8(vs[k]); acc_c[k] = reduce8(vc[k]); }
    for (; i < count; ++i) {   /* scalar remainder in double (:7988-8003) */
        const int32_t a = idx[i]; const double w = mass[a];
        for (int k = 0; k < 3; ++k) {
            const double p = src[k][a];
            const double t = p * M[k][0] + p * M[k][1] + p * M[k][2];
            acc_c[k] += w * cos(t); acc_s[k] += w * sin(t);
        }
        acc_w += w;
    }
    const double inv_w = 1.0 / acc_w;
    for (int k = 0; k < 3; ++k) {
        double theta = 3.14159265358979323846;
        const double px = acc_c[k] * inv_w, py = acc_s[k] * inv_w;
        if (px * px + py * py > 1.0e-8) theta += atan2(-py, -px);   /* TRIG_ATAN2_R2_THRESHOLD :5971 */
        out[k] = (float)(theta * I[k][0] + theta * I[k][1] + theta * I[k][2]);
    }
}

void mdo_arg_position(const float* x, const float* y, const float* z, const float* mass, const int32_t* idx, size_t count, int direct,
                      const mdo_unitcell_t* cell, float out[3]) {
    if (direct && count == 1) { out[0] = x[idx[0]]; out[1] = y[idx[0]]; out[2] = z[idx[0]]; return; }
    mdo_com(x, y, z, mass, idx, count, cell, out);
}

/* _distance md_script_functions.inl:3851-3890 on two positions */
float mdo_distance_pos(const float pa[3], const float pb_in[3], const mdo_unitcell_t* cell) {
    float pb[3] = { pb_in[0], pb_in[1], pb_in[2] };
    if (cell->flags & MDO_CELL_ORTHO) {   /* md_util_deperiodize_vec4 md_util.c:8971-8990 */
        const float ext[3] = { (float)cell->x, (float)cell->y, (float)cell->z };
        for (int k = 0; k < 3; ++k) pb[k] = deperiodize1(pb[k], pa[k], ext[k]);
    }
    const float d[3] = { pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2] };
    return sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}
float mdo_distance(const float* x, const float* y, const float* z, int32_t a, int32_t b, const mdo_unitcell_t* cell) {
    const float pa[3] = { x[a], y[a], z[a] }, pb[3] = { x[b], y[b], z[b] };
    return mdo_distance_pos(pa, pb, cell);
}

static void normalize3(float v[3]) {
    const float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (len > 1.0e-5) { v[0] = v[0] / len; v[1] = v[1] / len; v[2] = v[2] / len; } else { v[0] = v[1] = v[2] = 0; }
}

/* distance_min(a, b) / distance_max(a, b): _distance_min / _distance_max md_script_functions.inl:3892-3968 over the atoms of two selections.
 * Both call md_util_min_distance (md_util.c:8242-8297; _distance_max calling the MIN function is the reference's behaviour, :3944):
 * brute force over all pairs, vec4_periodic_distance (core/md_vec_math.h:1268-1273) in ortho cells, minimum_image_triclinic in triclinic. */
static float pair_distance(const float a[3], const float b[3], const mdo_unitcell_t* cell) {
    const float ext[3] = { (float)cell->x, (float)cell->y, (float)cell->z };
    const float box[3][3] = { { (float)cell->x, 0, 0 }, { (float)cell->xy, (float)cell->y, 0 }, { (float)cell->xz, (float)cell->yz, (float)cell->z } };
    float d[3] = { a[0] - b[0], a[1] - b[1], a[2] - b[2] };
    if (cell->flags == 0) return sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);                  /* vec3_distance */
    if (cell->flags & MDO_CELL_ORTHO) {
        for (int k = 0; k < 3; ++k) if (ext[k] != 0.0f) d[k] = d[k] - rintf(d[k] / ext[k]) * ext[k];
        return sqrtf((d[0] * d[0] + d[1] * d[1]) + (d[2] * d[2] + 0.0f));                          /* vec4_dot: md_mm_reduce_add_ps order */
    }
    min_image_triclinic(d, box); return sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);            /* vec3_length */
}

float mdo_min_distance(const float* x, const float* y, const float* z, const int32_t* ia, size_t na, const int32_t* ib, size_t nb, const mdo_unitcell_t* cell) {
    float min_dist = FLT_MAX;
    for (size_t i = 0; i < na; ++i) for (size_t j = 0; j < nb; ++j) {
        const float a[3] = { x[ia[i]], y[ia[i]], z[ia[i]] }, b[3] = { x[ib[j]], y[ib[j]], z[ib[j]] };
        const float dist = pair_distance(a, b, cell);
        if (dist < min_dist) min_dist = dist;
    }
    return min_dist;
}

/* distance_pair(a, b): _distance_pair md_script_functions.inl:3972-4064 -> md_util_distance_array (md_util.c:8210-8240): the na x nb
 * matrix out[i * nb + j] of the same three per-pair forms as above, a and b being the atoms of two selections (coordinate_extract). */
void mdo_distance_pair(const float* x, const float* y, const float* z, const int32_t* ia, size_t na, const int32_t* ib, size_t nb, const mdo_unitcell_t* cell, float* out) {
    for (size_t i = 0; i < na; ++i) for (size_t j = 0; j < nb; ++j) {
        const float a[3] = { x[ia[i]], y[ia[i]], z[ia[i]] }, b[3] = { x[ib[j]], y[ib[j]], z[ib[j]] };
        out[i * nb + j] = pair_distance(a, b, cell);
    }
}

/* The same matrix between given positions [na][3] x [nb][3] (arguments that were arrays of selections: one extract_com centre each) */
void mdo_distance_pair_pos(const float* pa, size_t na, const float* pb, size_t nb, const mdo_unitcell_t* cell, float* out) {
    for (size_t i = 0; i < na; ++i) for (size_t j = 0; j < nb; ++j) out[i * nb + j] = pair_distance(pa + 3 * i, pb + 3 * j, cell);
}

/* Per-frame aggregates of a multi-valued temporal: compute_min_max_mean_variance (md_script.c:5646-5677), two passes in float.
 * out[4] = min, max, mean, population variance. */
void mdo_aggregate(const float* data, size_t count, float out[4]) {
    const float N = (float)count;
    float mn = FLT_MAX, mx = -FLT_MAX, s1 = 0, s2 = 0;
    for (size_t i = 0; i < count; ++i) { s1 += data[i]; mn = MINV(mn, data[i]); mx = MAXV(mx, data[i]); }
    s1 = s1 / N;
    for (size_t i = 0; i < count; ++i) s2 += (data[i] - s1) * (data[i] - s1);
    s2 = s2 / N;
    out[0] = mn; out[1] = mx; out[2] = s1; out[3] = s2;
}

/* ------------------------------------------------------------------------------------------------
 * Shape weights of a structure: the loop body of VIAMD's shape-space component (src/components/shapespace/shapespace.cpp:418-431) and of
 * _shape_weights (md_script_functions.inl:6033-6040): xyzw (mass or 1) -> md_util_com_compute_vec4 with the cell (com_pbc_vec4
 * md_util.c:8063-8162: serial float accumulation of w*sin, w*cos per axis, 4-lane sincos, double atan2) -> md_util_deperiodize_vec4 about it
 * (:8971-9005; the triclinic branch starts at atom 1 and the triclinic centre goes through the 1/2pi-scaled inverse twice — both as written)
 * -> mat3_covariance_matrix_vec4 about that centre -> md_util_shape_weights (:9070-9076): eigenvalues e0 >= e1 >= e2 (normalised by the
 * largest) -> ((e0 - e1), 2 (e1 - e2), 3 e2) / (e0 + e1 + e2).
 */
static float m3_eigen_values(m3 M, float ev_sorted[3]) {
    svd_t s = m3_svd(M);
    const float mx = MAXV(s.s[0], MAXV(s.s[1], s.s[2]));
    const float ev[3] = { s.s[0] / mx, s.s[1] / mx, s.s[2] / mx };
    int l[3] = { 0, 1, 2 }, t;
    if (ev[l[0]] < ev[l[1]]) { t = l[0]; l[0] = l[1]; l[1] = t; }
    if (ev[l[1]] < ev[l[2]]) { t = l[1]; l[1] = l[2]; l[2] = t; }
    if (ev[l[0]] < ev[l[1]]) { t = l[0]; l[0] = l[1]; l[1] = t; }
    ev_sorted[0] = ev[l[0]]; ev_sorted[1] = ev[l[1]]; ev_sorted[2] = ev[l[2]];
    return mx;
}
/* md_util_com_compute_vec4 (md_util.c:8188-8201) of n points xyzw: com_pbc_vec4 (:8063-8162: serial float accumulation of w*sin, w*cos per
 * axis, 4-lane sincos, double atan2; the triclinic centre goes through the 1/2pi-scaled inverse twice, as written) in a cell, com_vec4 (:8048) without */
static void com_compute_v4(float com[3], const v4* p, size_t n, const mdo_unitcell_t* cell) {
    const double TWO_PI_D = 2.0 * 3.1415926535897932, PI_D = 3.1415926535897932;
    if (cell->flags & MDO_CELL_ORTHO) {
        const float ext[3] = { (float)cell->x, (float)cell->y, (float)cell->z };
        const float tp = (float)TWO_PI_D;
        const float scl[4] = { tp / ext[0], tp / ext[1], tp / ext[2], tp / tp };
        float as[4] = { 0 }, ac[4] = { 0 }, ax[4] = { 0 };
        for (size_t k = 0; k < n; ++k) {
            const float www1[4] = { p[k][3], p[k][3], p[k][3], 1.0f };
            for (int a = 0; a < 4; ++a) {
                float sn, cs; ref_sincosf4(p[k][a] * scl[a], &sn, &cs);
                as[a] = as[a] + sn * www1[a]; ac[a] = ac[a] + cs * www1[a]; ax[a] = ax[a] + p[k][a] * www1[a];
            }
        }
        const float w = ax[3];
        for (int a = 0; a < 3; ++a) {
            const double yy = as[a] / w, xx = ac[a] / w, r2 = xx * xx + yy * yy;
            double theta = PI_D; if (r2 > 1.0e-15) theta += atan2(-yy, -xx);
            com[a] = (float)((theta / TWO_PI_D) * ext[a]);
        }
    } else if (cell->flags & MDO_CELL_TRICLINIC) {
        double Id[3][3]; cell_I(Id, cell);
        float I[3][3];   /* I = mat3_mul(mat3_scale(1/2pi), Ai): C[col][row] = S[row][row] * Ai[col][row] (core/md_vec_math.h:1631) */
        const float inv_tp = 1.0f / (float)TWO_PI_D;
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) {   /* exact form of MULT(col,row): A.e[0][row]*B.e[col][0] + A.e[1][row]*B.e[col][1] + A.e[2][row]*B.e[col][2] with A = scale */
            const float S[3][3] = { { inv_tp, 0, 0 }, { 0, inv_tp, 0 }, { 0, 0, inv_tp } };
            const float Ai[3] = { (float)Id[c][0], (float)Id[c][1], (float)Id[c][2] };
            I[c][r] = S[0][r] * Ai[0] + S[1][r] * Ai[1] + S[2][r] * Ai[2];
        }
        float as[4] = { 0 }, ac[4] = { 0 }, ax[4] = { 0 };
        for (size_t k = 0; k < n; ++k) {
            const float www1[4] = { p[k][3], p[k][3], p[k][3], 1.0f };
            float th[4];   /* mat4x3_mul_vec4(I, xyzw) (in_idx == NULL branch :8138): linear_combine_3, w lane = 0 */
            for (int a = 0; a < 3; ++a) th[a] = (p[k][0] * I[0][a] + p[k][1] * I[1][a]) + p[k][2] * I[2][a];
            th[3] = (p[k][0] * 0.0f + p[k][1] * 0.0f) + p[k][2] * 0.0f;
            for (int a = 0; a < 4; ++a) {
                float sn, cs; ref_sincosf4(th[a], &sn, &cs);
                as[a] = as[a] + sn * www1[a]; ac[a] = ac[a] + cs * www1[a]; ax[a] = ax[a] + p[k][a] * www1[a];
            }
        }
        for (int a = 0; a < 3; ++a) {
            const double yy = as[a] / ax[3], xx = ac[a] / ax[3], r2 = xx * xx + yy * yy;
            double theta = PI_D; if (r2 > 1.0e-8) theta += atan2(-yy, -xx);
            com[a] = (float)(theta * I[a][0] + theta * I[a][1] + theta * I[a][2]);   /* :8158, I.elem[i][0..2] */
        }
    } else {
        com_v4(com, p, n);   /* no cell: com_vec4 */
    }
}

/* position of an argument that is an ARRAY of selections (coordinate_extract_com md_script_functions.inl:1826-1842): md_util_com_compute of each
 * selection (mdo_com), then md_util_com_compute_vec4 over those centres with weight 1. idx holds the selections back to back, off their CSR offsets. */
void mdo_arg_position_parts(const float* x, const float* y, const float* z, const float* mass, const int32_t* idx, const uint32_t* off, size_t n_parts,
                            const mdo_unitcell_t* cell, float out[3]) {
    v4* p = malloc(sizeof(v4) * (n_parts ? n_parts : 1));
    for (size_t k = 0; k < n_parts; ++k) {
        float c[3] = { 0.0f, 0.0f, 0.0f };
        if (off[k + 1] > off[k]) mdo_com(x, y, z, mass, idx + off[k], off[k + 1] - off[k], cell, c);   /* count == 0 -> (0, 0, 0) (md_util.c:8168) */
        p[k][0] = c[0]; p[k][1] = c[1]; p[k][2] = c[2]; p[k][3] = 1.0f;
    }
    out[0] = out[1] = out[2] = 0.0f;
    if (n_parts) com_compute_v4(out, p, n_parts, cell);
    free(p);
}

void mdo_shape_weights(const float* x, const float* y, const float* z, const float* mass, const int32_t* idx, size_t n, const mdo_unitcell_t* cell, float out[3]) {
    out[0] = out[1] = out[2] = 0.0f;
    if (n == 0) return;
    v4* p = malloc(sizeof(v4) * n);
    for (size_t k = 0; k < n; ++k) { const int32_t a = idx[k]; p[k][0] = x[a]; p[k][1] = y[a]; p[k][2] = z[a]; p[k][3] = mass ? mass[a] : 1.0f; }
    float com[3];
    com_compute_v4(com, p, n, cell);
    if (cell->flags & MDO_CELL_ORTHO) {
        const float ext[3] = { (float)cell->x, (float)cell->y, (float)cell->z };
        for (size_t k = 0; k < n; ++k) for (int a = 0; a < 3; ++a) p[k][a] = deperiodize1(p[k][a], com[a], ext[a]);
    } else if (cell->flags & MDO_CELL_TRICLINIC) {
        double Ad[3][3]; cell_A(Ad, cell);
        float A[3][3];
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) A[c][r] = (float)Ad[c][r];
        const float box[3][3] = { { A[0][0], 0, 0 }, { A[1][0], A[1][1], 0 }, { A[2][0], A[2][1], A[2][2] } };
        for (size_t k = 1; k < n; ++k) {   /* deperiodize_triclinic from atom 1 on (:8993) */
            float d[3] = { p[k][0] - com[0], p[k][1] - com[1], p[k][2] - com[2] };
            min_image_triclinic(d, box);
            p[k][0] = com[0] + d[0]; p[k][1] = com[1] + d[1]; p[k][2] = com[2] + d[2];
        }
    }
    float ev[3]; m3_eigen_values(covariance_v4(p, n, com), ev);
    const float scl = 1.0f / (ev[0] + ev[1] + ev[2]);
    out[0] = (ev[0] - ev[1]) * scl; out[1] = 2.0f * (ev[1] - ev[2]) * scl; out[2] = 3.0f * ev[2] * scl;
    free(p);
}

/* _angle :4099-4114 */
float mdo_angle_pos(const float a[3], const float b[3], const float c[3]) {
    float v0[3] = { a[0] - b[0], a[1] - b[1], a[2] - b[2] }, v1[3] = { c[0] - b[0], c[1] - b[1], c[2] - b[2] };
    normalize3(v0); normalize3(v1);
    return acosf(v0[0] * v1[0] + v0[1] * v1[1] + v0[2] * v1[2]);
}
float mdo_angle(const float* x, const float* y, const float* z, int32_t a, int32_t b, int32_t c) {
    const float pa[3] = { x[a], y[a], z[a] }, pb[3] = { x[b], y[b], z[b] }, pc[3] = { x[c], y[c], z[c] };
    return mdo_angle_pos(pa, pb, pc);
}

/* _dihedral :4171-4196 */
float mdo_dihedral_pos(const float p[4][3], const mdo_unitcell_t* cell) {
    float dx[3][3];
    for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) dx[k][i] = p[k + 1][i] - p[k][i];
    if (cell->flags & MDO_CELL_ORTHO) {   /* md_util_min_image_vec3 -> min_image_ortho md_util.c:8424-8436 */
        const float ext[3] = { (float)cell->x, (float)cell->y, (float)cell->z };
        for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) {
            const float half = ext[i] * 0.5f;
            if (ext[i] > 0.0f) { while (dx[k][i] > half) dx[k][i] -= ext[i]; while (dx[k][i] <= -half) dx[k][i] += ext[i]; }
        }
    } else if (cell->flags & MDO_CELL_TRICLINIC) {   /* min_image_triclinic with the half diagonal md_util.c:8360-8423: zone reduction along c, b, a, then the 27 images */
        double Ad[3][3]; cell_A(Ad, cell);
        float box[3][3]; for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) box[c][r] = (float)Ad[c][r];
        const float half[3] = { box[0][0] * 0.5f, box[1][1] * 0.5f, box[2][2] * 0.5f };
        for (int k = 0; k < 3; ++k) {
            for (int i = 2; i >= 0; i--) if (half[i] > 0.0f) {
                while (dx[k][i] > half[i]) for (int j = i; j >= 0; j--) dx[k][j] -= box[i][j];
                while (dx[k][i] <= -half[i]) for (int j = i; j >= 0; j--) dx[k][j] += box[i][j];
            }
            min_image_triclinic(dx[k], (const float (*)[3])box);
        }
    }
    /* vec3_dihedral_angle core/md_vec_math.h:558-567 */
    const float* d1 = dx[0]; const float* d2 = dx[1]; const float* d3 = dx[2];
    const float v1[3] = { d1[1] * d2[2] - d1[2] * d2[1], d1[2] * d2[0] - d1[0] * d2[2], d1[0] * d2[1] - d1[1] * d2[0] };
    const float v2[3] = { d2[1] * d3[2] - d2[2] * d3[1], d2[2] * d3[0] - d2[0] * d3[2], d2[0] * d3[1] - d2[1] * d3[0] };
    const float w[3] = { v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0] };
    const float wl = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const float s = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
    float angle = atan2f(wl, s);
    const float dot = d1[0] * v2[0] + d1[1] * v2[1] + d1[2] * v2[2];
    if (dot < 0) angle = -angle;
    return angle;
}
float mdo_dihedral(const float* x, const float* y, const float* z, int32_t a, int32_t b, int32_t c, int32_t d, const mdo_unitcell_t* cell) {
    const int32_t id[4] = { a, b, c, d }; float p[4][3];
    for (int k = 0; k < 4; ++k) { p[k][0] = x[id[k]]; p[k][1] = y[id[k]]; p[k][2] = z[id[k]]; }
    return mdo_dihedral_pos((const float (*)[3])p, cell);
}

/* extract_com md_script_functions.inl:857-874: vec4 sum += (x,y,z,1) * w in atom order; w == 0 -> 1; xyz / w (vec3_div1 md_vec_math.h:471) */
void mdo_group_com(const float* x, const float* y, const float* z, const float* mass,
                   const int32_t* idx, const uint32_t* off, size_t n_groups, float* out) {
    for (size_t g = 0; g < n_groups; ++g) {
        volatile float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;   /* volatile: keep every intermediate rounded to float */
        for (uint32_t k = off[g]; k < off[g + 1]; ++k) {
            const int32_t a = idx[k]; const float w = mass ? mass[a] : 1.0f;
            volatile float px = x[a] * w, py = y[a] * w, pz = z[a] * w, pw = 1.0f * w;
            sx = sx + px; sy = sy + py; sz = sz + pz; sw = sw + pw;
        }
        if (sw == 0.0f) sw = 1.0f;
        out[3 * g + 0] = sx / sw; out[3 * g + 1] = sy / sw; out[3 * g + 2] = sz / sw;
    }
}

/* ------------------------------------------------------------------------------------------------
 * XTC frame decode: md_xtc_decode_frame_data_soa_scaled (md_xtc.c:747-931) as xtc_reader_load_frame calls it (:947-993, scale 10:
 * nm -> Angstrom), restated with a plain bit reader and 128-bit integers. The packed integers are transmitted as whole bytes,
 * least significant first, followed by the remaining high bits (ext/xtc/xdrfile.c: sendints/receiveints; md_xtc.c:304-315 reads
 * the same thing with a byte swap).
 */
static const uint32_t xtc_magicints[] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096,
    5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287,
    660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216 };
#define XTC_FIRSTIDX 9
#define XTC_LASTIDX ((int)(sizeof(xtc_magicints) / sizeof(xtc_magicints[0])))

typedef struct { const uint8_t* p; size_t nbits_total; size_t pos; } xtc_bits_t;
static uint32_t xtc_get(xtc_bits_t* b, unsigned n) {   /* n <= 32 bits, most significant first */
    uint32_t v = 0;
    for (unsigned i = 0; i < n; ++i, ++b->pos) {
        const unsigned bit = b->pos < b->nbits_total ? (b->p[b->pos >> 3] >> (7 - (b->pos & 7))) & 1u : 0u;
        v = (v << 1) | bit;
    }
    return v;
}
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static float bef32(const uint8_t* p) { return f_from_bits(be32(p)); }

static int xtc_sizeofint(uint32_t size) {   /* md_xtc.c:157-166 */
    uint32_t num = 1; int nb = 0;
    while ((int)size >= (int)num && nb < 32) { nb++; num *= 2; }
    return nb;
}
static int xtc_sizeofints(const uint32_t sizes[3]) {   /* md_xtc.c:168-193: bits of the product of the three sizes */
    unsigned __int128 prod = (unsigned __int128)sizes[0] * sizes[1]; prod *= sizes[2];
    int nbytes = 0; unsigned __int128 t = prod; while (t > 0xff) { t >>= 8; nbytes++; }
    uint32_t top = (uint32_t)t, num = 1; int nb = 0;
    while (top >= num) { nb++; num *= 2; }
    return nb + nbytes * 8;
}
/* `nbits` wide packed field -> three integers with mixed radix (size_y, size_z) */
static void xtc_unpack3(xtc_bits_t* b, unsigned nbits, uint32_t size_y, uint32_t size_z, int32_t out[3]) {
    unsigned __int128 v = 0; unsigned shift = 0, left = nbits;
    while (left >= 8) { v |= (unsigned __int128)xtc_get(b, 8) << shift; shift += 8; left -= 8; }
    if (left) v |= (unsigned __int128)xtc_get(b, left) << shift;
    const unsigned __int128 zy = (unsigned __int128)size_z * size_y;
    const uint32_t x = (uint32_t)(v / zy);
    const uint64_t yz = (uint64_t)(v / size_z);
    out[0] = (int32_t)x; out[1] = (int32_t)(uint32_t)(yz - (uint64_t)x * size_y); out[2] = (int32_t)(uint32_t)((uint64_t)v - yz * size_z);
}
/* bitsize == 0 branch (an axis spans more than 0xffffff units): one plain big-endian integer per axis, as the format's writer emits them
 * (ext/xtc/xdrfile.c: sendbits / receivebits). NOTE: the reference's reader mis-decodes this branch (unpack_uint32 md_xtc.c:295 applies the
 * byte order of the packed fields and shifts a 32-bit word by a 64-bit alignment) and also the > 64-bit packed field (unpack_coord128 :317);
 * both only occur for boxes wider than ~2.6 um. Here they follow the file format, checked by round trip against the written coordinates. */
static uint32_t xtc_unpack1(xtc_bits_t* b, unsigned nbits) { return xtc_get(b, nbits); }

int mdo_xtc_decode_frame(const uint8_t* frame, size_t nbytes, size_t num_atoms, float* x, float* y, float* z,
                         mdo_unitcell_t* cell, int32_t* step, float* time) {
    const float scale = 10.0f;   /* xtc_reader_load_frame :976 */
    if (!frame || nbytes < 56 || be32(frame) != 1995u) return 0;
    const int32_t natoms = (int32_t)be32(frame + 4);
    if (step) *step = (int32_t)be32(frame + 8);
    if (time) *time = bef32(frame + 12);
    if (cell) {   /* box[i] *= scale (:765-768); md_unitcell_from_matrix_float md_unitcell.inl:109 -> from_basis_parameters :12-31 */
        float box[9]; for (int i = 0; i < 9; ++i) box[i] = bef32(frame + 16 + 4 * i) * scale;
        const double cx = box[0], cy = box[4], cz = box[8], xy = box[3], xz = box[6], yz = box[7];
        uint32_t flags = 0;
        if (xy == 0.0 && xz == 0.0 && yz == 0.0) { if (!(cx == 0.0 && cy == 0.0 && cz == 0.0) && !(cx == 1.0 && cy == 1.0 && cz == 1.0)) flags |= MDO_CELL_ORTHO; }
        else flags |= MDO_CELL_TRICLINIC;
        if (flags) { if (cx != 0.0) flags |= MDO_CELL_PBC_X; if (cy != 0.0) flags |= MDO_CELL_PBC_Y; if (cz != 0.0) flags |= MDO_CELL_PBC_Z; }
        cell->x = cx; cell->xy = xy; cell->xz = xz; cell->y = cy; cell->yz = yz; cell->z = cz; cell->flags = flags;
    }
    if (!x || !y || !z) return 1;
    if ((int32_t)be32(frame + 52) != (int32_t)num_atoms || natoms != (int32_t)num_atoms) return 0;
    size_t off = 56;
    if (natoms <= 9) {
        if (nbytes < off + 12u * (size_t)natoms) return 0;
        for (int i = 0; i < natoms; ++i) { x[i] = bef32(frame + off + 12 * i) * scale; y[i] = bef32(frame + off + 12 * i + 4) * scale; z[i] = bef32(frame + off + 12 * i + 8) * scale; }
        return 1;
    }
    if (nbytes < 92) return 0;
    const float precision = bef32(frame + off); off += 4;
    int32_t minint[3], maxint[3];
    for (int k = 0; k < 3; ++k) minint[k] = (int32_t)be32(frame + off + 4 * k);
    off += 12;
    for (int k = 0; k < 3; ++k) maxint[k] = (int32_t)be32(frame + off + 4 * k);
    off += 12;
    int smallidx = (int32_t)be32(frame + off); off += 4;
    if (smallidx < XTC_FIRSTIDX || smallidx >= XTC_LASTIDX) return 0;
    const uint32_t sizeint[3] = { (uint32_t)(maxint[0] - minint[0] + 1), (uint32_t)(maxint[1] - minint[1] + 1), (uint32_t)(maxint[2] - minint[2] + 1) };
    unsigned bitsize = 0, bitsizeint[3] = { 0, 0, 0 };
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) { for (int k = 0; k < 3; ++k) bitsizeint[k] = (unsigned)xtc_sizeofint(sizeint[k]); }
    else bitsize = (unsigned)xtc_sizeofints(sizeint);
    int smaller = (int)(xtc_magicints[smallidx - 1 > XTC_FIRSTIDX ? smallidx - 1 : XTC_FIRSTIDX] / 2);
    int smallnum = (int)(xtc_magicints[smallidx] / 2);
    const uint32_t data_bytes = be32(frame + off); off += 4;
    if (nbytes < off + data_bytes) return 0;
    xtc_bits_t br = { frame + off, (size_t)data_bytes * 8, 0 };
    const float coord_scale = scale / precision;
    int run = 0, run_count = 0, atom = 0;
    int32_t c[3];
    while (atom < natoms) {
        if (bitsize == 0) { for (int k = 0; k < 3; ++k) c[k] = (int32_t)xtc_unpack1(&br, bitsizeint[k]); }
        else xtc_unpack3(&br, bitsize, sizeint[1], sizeint[2], c);
        for (int k = 0; k < 3; ++k) c[k] += minint[k];
        const uint32_t flag = xtc_get(&br, 1);
        int is_smaller = 0;
        if (flag) { run = (int)xtc_get(&br, 5); run_count = run / 3; is_smaller = run % 3; run -= is_smaller; is_smaller--; }
        if (atom + run_count + 1 > natoms) return 0;
        if (run > 0) {
            const int32_t prev[3] = { c[0], c[1], c[2] };
            const uint32_t ss = xtc_magicints[smallidx];
            int32_t d[3];
            xtc_unpack3(&br, (unsigned)smallidx, ss, ss, d);
            for (int k = 0; k < 3; ++k) c[k] = d[k] + (c[k] - smallnum);
            x[atom] = (float)c[0] * coord_scale; y[atom] = (float)c[1] * coord_scale; z[atom] = (float)c[2] * coord_scale; ++atom;      /* first two swapped (:855-856) */
            x[atom] = (float)prev[0] * coord_scale; y[atom] = (float)prev[1] * coord_scale; z[atom] = (float)prev[2] * coord_scale; ++atom;
            for (int i = 1; i < run_count; ++i) {
                xtc_unpack3(&br, (unsigned)smallidx, ss, ss, d);
                for (int k = 0; k < 3; ++k) c[k] = d[k] + (c[k] - smallnum);
                x[atom] = (float)c[0] * coord_scale; y[atom] = (float)c[1] * coord_scale; z[atom] = (float)c[2] * coord_scale; ++atom;
            }
        } else {
            x[atom] = (float)c[0] * coord_scale; y[atom] = (float)c[1] * coord_scale; z[atom] = (float)c[2] * coord_scale; ++atom;
        }
        smallidx += is_smaller;
        if (is_smaller < 0) { smallnum = smaller; smaller = smallidx > XTC_FIRSTIDX ? (int)(xtc_magicints[smallidx - 1] / 2) : 0; }
        else if (is_smaller > 0) { smaller = smallnum; smallnum = (int)(xtc_magicints[smallidx] / 2); }
        if (smallidx < XTC_FIRSTIDX || smallidx >= XTC_LASTIDX) return 0;
    }
    return atom == natoms;
}

/* md_xtc_read_frame_offsets_and_times md_xtc.c:436-570 on a file image: offsets[0..n] (offsets[n] = end); returns n */
size_t mdo_xtc_frame_offsets(const uint8_t* file, size_t nbytes, int64_t* offsets, size_t cap) {
    if (nbytes < 56 || be32(file) != 1995u || cap < 2) return 0;
    const int32_t natoms = (int32_t)be32(file + 4);
    if (natoms <= 0) return 0;
    size_t n = 0, pos = 0;
    if (natoms <= 9) {
        const size_t fb = 56 + 12u * (size_t)natoms;
        while (pos + fb <= nbytes && n + 1 < cap && be32(file + pos) == 1995u) { offsets[n++] = (int64_t)pos; pos += fb; }
        offsets[n] = (int64_t)(n * fb);
        return n;
    }
    while (pos != nbytes && n + 1 < cap) {
        if (pos + 92 > nbytes || be32(file + pos) != 1995u) break;
        const size_t fb = ((size_t)be32(file + pos + 88) + 3u) & ~(size_t)3;
        if (pos + 92 + fb > nbytes) break;
        offsets[n++] = (int64_t)pos; pos += 92 + fb;
    }
    offsets[n] = (int64_t)pos;
    return n;
}
