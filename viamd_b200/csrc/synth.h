/* synth.h — deterministic synthetic trajectories for the BASELINE.json configs.
 *
 * Plain C99 that also compiles as CUDA device code (MDSYNTH_HD). The same
 * functions are used by
 *   - the GPU frame generator (csrc/synth.cu, frames produced directly in HBM),
 *   - the host frame source of the C-ABI (csrc/host_api.cpp),
 *   - the reference harness (oracle/ref_harness.c, in-memory md_trajectory_i),
 * so that every consumer sees bit-identical coordinates.
 *
 * Bit-reproducibility rule: the per-frame path uses only integer hashing,
 * exact int->float conversion, multiplication by a power of two (exact) and a
 * single float add per coordinate, so FMA contraction cannot change a result.
 * Everything that needs sin/cos/sqrt (molecule orientation) lives in the
 * one-off base-configuration builder and is executed on the host only; the
 * resulting float arrays are what is shipped to the device / the harness.
 *
 * Workloads follow SURVEY.md §8(d):
 *   water: n^3 lattice of rigid 3-site waters (OW,HW1,HW2), spacing 3.104 A,
 *          O jitter U(-0.6,0.6) A, random orientation, cubic ortho cell L=n*3.104.
 *          Frame f = base + per-molecule rigid displacement d(seed,f,mol),
 *          d = (sum of 4 hash bytes - 510) * 2^-9 A per axis (sigma ~0.289 A),
 *          then each ATOM is wrapped into [0,L) on its own (molecules may be
 *          split across the boundary, as in GROMACS "atom" pbc output; the SDF
 *          unwrap path must cope with that).
 */
#ifndef MDSYNTH_H
#define MDSYNTH_H

#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define MDSYNTH_HD __host__ __device__ static inline
#else
#define MDSYNTH_HD static inline
#endif

/* 32-bit avalanche mixer (lowbias32-style constants); integer only. */
MDSYNTH_HD uint32_t mdsynth_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU;
    x ^= x >> 15; x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

MDSYNTH_HD uint32_t mdsynth_hash4(uint32_t seed, uint32_t a, uint32_t b, uint32_t c) {
    uint32_t h = mdsynth_mix(seed ^ 0x9e3779b9U);
    h = mdsynth_mix(h ^ (a + 0x85ebca6bU));
    h = mdsynth_mix(h ^ (b + 0xc2b2ae35U));
    h = mdsynth_mix(h ^ (c + 0x27d4eb2fU));
    return h;
}

/* Displacement of molecule `mol` along `axis` in frame `frame`, in Angstrom.
 * Exactly representable: integer in [-510,510] times 2^-9. */
MDSYNTH_HD float mdsynth_disp(uint32_t seed, uint32_t frame, uint32_t mol, uint32_t axis) {
    uint32_t h = mdsynth_hash4(seed, frame, mol, axis);
    int s = (int)(h & 0xff) + (int)((h >> 8) & 0xff) + (int)((h >> 16) & 0xff) + (int)(h >> 24) - 510;
    return (float)s * 0.001953125f; /* 2^-9 */
}

/* Wrap a coordinate into [0,L). One conditional add each way; inputs never
 * leave (-L, 2L) in these workloads. */
MDSYNTH_HD float mdsynth_wrap(float x, float L) {
    if (x < 0.0f) x = x + L;
    if (x >= L)   x = x - L;
    /* x + L can round to exactly L for tiny negative x */
    if (x >= L)   x = 0.0f;
    return x;
}

/* Frame coordinate of one atom: base + disp(mol), wrapped. */
MDSYNTH_HD float mdsynth_frame_coord(float base, uint32_t seed, uint32_t frame, uint32_t mol, uint32_t axis, float L) {
    float d = mdsynth_disp(seed, frame, mol, axis);
    float x = base + d;
    return mdsynth_wrap(x, L);
}

#include <math.h>
#include <stdio.h>

typedef struct mdsynth_water_t {
    uint32_t seed;
    uint32_t n;            /* lattice points per axis */
    uint32_t num_mol;      /* n^3 */
    uint32_t num_atoms;    /* 3 n^3 */
    float    L;            /* cubic cell edge, Angstrom */
} mdsynth_water_t;

static inline mdsynth_water_t mdsynth_water_desc(uint32_t n, uint32_t seed) {
    mdsynth_water_t w;
    w.seed = seed; w.n = n; w.num_mol = n * n * n; w.num_atoms = 3 * w.num_mol;
    w.L = (float)n * 3.104f;
    return w;
}

static inline double mdsynth_u01(uint32_t h) { return ((double)(h >> 8) + 0.5) * (1.0 / 16777216.0); }

/* Base configuration. Writes two sets of float SoA arrays of length 3*n^3:
 *   base_*  : every atom wrapped into [0,L)           (what frames are built from)
 *   whole_* : O wrapped, H placed relative to O        (unbroken molecules, for the .gro topology file)
 * Atom order per molecule: OW, HW1, HW2. Either output set may be NULL. */
static inline void mdsynth_water_base(const mdsynth_water_t* w,
                                      float* base_x, float* base_y, float* base_z,
                                      float* whole_x, float* whole_y, float* whole_z) {
    const double spacing = 3.104;
    const double r_oh = 0.9572;
    const double half = 0.5 * 104.52 * (3.14159265358979323846 / 180.0);
    const double hx = r_oh * cos(half), hy = r_oh * sin(half);
    const float L = w->L;
    for (uint32_t iz = 0; iz < w->n; ++iz)
    for (uint32_t iy = 0; iy < w->n; ++iy)
    for (uint32_t ix = 0; ix < w->n; ++ix) {
        const uint32_t m = (iz * w->n + iy) * w->n + ix;
        const uint32_t lat[3] = { ix, iy, iz };
        double o[3];
        for (uint32_t a = 0; a < 3; ++a) {
            double u = mdsynth_u01(mdsynth_hash4(w->seed, 0xBA5E0000u + a, m, 1));
            o[a] = ((double)lat[a] + 0.5) * spacing + (u * 1.2 - 0.6);
        }
        /* uniform random rotation from a normalised 4-vector */
        double q[4], nq = 0.0;
        for (uint32_t a = 0; a < 4; ++a) {
            /* approx. normal via sum of uniforms; exact distribution is irrelevant */
            double s = 0.0;
            for (uint32_t k = 0; k < 4; ++k) s += mdsynth_u01(mdsynth_hash4(w->seed, 0x0A170000u + a * 4 + k, m, 2));
            q[a] = s - 2.0; nq += q[a] * q[a];
        }
        if (nq < 1e-12) { q[0] = 1; q[1] = q[2] = q[3] = 0; nq = 1; }
        nq = 1.0 / sqrt(nq);
        const double qw = q[0] * nq, qx = q[1] * nq, qy = q[2] * nq, qz = q[3] * nq;
        const double R[3][3] = {
            { 1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw),     2 * (qx * qz + qy * qw) },
            { 2 * (qx * qy + qz * qw),     1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw) },
            { 2 * (qx * qz - qy * qw),     2 * (qy * qz + qx * qw),     1 - 2 * (qx * qx + qy * qy) },
        };
        const double hloc[2][3] = { { hx, hy, 0.0 }, { hx, -hy, 0.0 } };
        float* bout[3] = { base_x, base_y, base_z };
        float* wout[3] = { whole_x, whole_y, whole_z };
        for (uint32_t a = 0; a < 3; ++a) {
            const float of = mdsynth_wrap((float)o[a], L);
            if (bout[a]) bout[a][3 * m] = of;
            if (wout[a]) wout[a][3 * m] = of;
            for (uint32_t h = 0; h < 2; ++h) {
                const double d = R[a][0] * hloc[h][0] + R[a][1] * hloc[h][1] + R[a][2] * hloc[h][2];
                const float hf = (float)((double)of + d);
                if (wout[a]) wout[a][3 * m + 1 + h] = hf;
                if (bout[a]) bout[a][3 * m + 1 + h] = mdsynth_wrap(hf, L);
            }
        }
    }
}

/* One frame on the host (same arithmetic as the device generator). */
static inline void mdsynth_water_frame(const mdsynth_water_t* w, uint32_t frame,
                                       const float* base_x, const float* base_y, const float* base_z,
                                       float* x, float* y, float* z) {
    for (uint32_t i = 0; i < w->num_atoms; ++i) {
        const uint32_t mol = i / 3;
        x[i] = mdsynth_frame_coord(base_x[i], w->seed, frame, mol, 0, w->L);
        y[i] = mdsynth_frame_coord(base_y[i], w->seed, frame, mol, 1, w->L);
        z[i] = mdsynth_frame_coord(base_z[i], w->seed, frame, mol, 2, w->L);
    }
}

/* GROMACS .gro topology/coordinate file (nm, 3 decimals) of the unbroken base
 * configuration; residue SOL, atoms OW/HW1/HW2. Returns 0 on success. */
static inline int mdsynth_water_write_gro(const mdsynth_water_t* w, const char* path,
                                          const float* whole_x, const float* whole_y, const float* whole_z) {
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "synthetic water n=%u seed=%u\n%u\n", w->n, w->seed, w->num_atoms);
    static const char* names[3] = { "OW", "HW1", "HW2" };
    for (uint32_t i = 0; i < w->num_atoms; ++i) {
        const uint32_t m = i / 3;
        fprintf(f, "%5u%-5s%5s%5u%8.3f%8.3f%8.3f\n", (m + 1) % 100000, "SOL", names[i % 3], (i + 1) % 100000,
                whole_x[i] * 0.1, whole_y[i] * 0.1, whole_z[i] * 0.1);
    }
    fprintf(f, "%10.5f%10.5f%10.5f\n", w->L * 0.1, w->L * 0.1, w->L * 0.1);
    fclose(f);
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * membrane (BASELINE config 4): coarse-grained bilayer slab + solvent beads in an orthorhombic cell.
 *   nl x nl lipids per leaflet on an 8 A grid (Lx = Ly = 8*nl), two leaflets, 12 beads per lipid
 *   (NC3 PO4 GL1 GL2 | C1A C2A C3A C4A | C1B C2B C3B C4B; tails 4.7 A bead spacing along -/+z), one residue "LIP" per lipid;
 *   nwz layers of nw_xy x nw_xy solvent beads "W" (one residue "SOLW" each) above and below the slab; Lz from the content.
 *   Frame f = base + per-molecule rigid displacement (same law as the water box), each atom wrapped per axis.
 * name('C2*') selects the two C2 tail beads of every lipid (the "lipid-tail" selection of config 4). */
#define MDSYNTH_LIPID_BEADS 12
typedef struct mdsynth_membrane_t {
    uint32_t seed, nl, nw_xy, nwz;
    uint32_t num_lipids, num_water, num_atoms, num_mol;
    float Lx, Ly, Lz;
} mdsynth_membrane_t;

static inline mdsynth_membrane_t mdsynth_membrane_desc(uint32_t nl, uint32_t nw_xy, uint32_t nwz, uint32_t seed) {
    mdsynth_membrane_t m;
    m.seed = seed; m.nl = nl; m.nw_xy = nw_xy; m.nwz = nwz;
    m.num_lipids = 2 * nl * nl; m.num_water = 2 * nwz * nw_xy * nw_xy;
    m.num_atoms = m.num_lipids * MDSYNTH_LIPID_BEADS + m.num_water; m.num_mol = m.num_lipids + m.num_water;
    m.Lx = m.Ly = 8.0f * (float)nl;
    m.Lz = 48.0f + 2.0f * (float)nwz * 4.5f;   /* 2 x 24 A leaflets + solvent layers of 4.5 A */
    return m;
}

static const char* const mdsynth_lipid_names[MDSYNTH_LIPID_BEADS] = { "NC3", "PO4", "GL1", "GL2", "C1A", "C2A", "C3A", "C4A", "C1B", "C2B", "C3B", "C4B" };

/* base_xyz: [3][num_atoms] wrapped; whole_xyz: unbroken molecules (for the .gro); mol_id: [num_atoms] molecule of each atom. */
static inline void mdsynth_membrane_base(const mdsynth_membrane_t* m, float* base_xyz, float* whole_xyz, uint32_t* mol_id) {
    const size_t N = m->num_atoms;
    const float L[3] = { m->Lx, m->Ly, m->Lz };
    const double zc = 0.5 * (double)m->Lz;
    /* bead offsets along the leaflet normal (towards the bilayer centre) and a small in-plane fan for the two tails */
    static const double dn[MDSYNTH_LIPID_BEADS] = { 0.0, 3.5, 7.0, 7.0, 10.5, 15.2, 19.9, 24.0 - 0.4, 10.5, 15.2, 19.9, 24.0 - 0.4 };
    static const double dxp[MDSYNTH_LIPID_BEADS] = { 0.0, 0.0, -1.8, 1.8, -2.2, -2.2, -2.2, -2.2, 2.2, 2.2, 2.2, 2.2 };
    size_t a = 0; uint32_t mol = 0;
    for (uint32_t leaf = 0; leaf < 2; ++leaf)
    for (uint32_t iy = 0; iy < m->nl; ++iy)
    for (uint32_t ix = 0; ix < m->nl; ++ix, ++mol) {
        const double jx = mdsynth_u01(mdsynth_hash4(m->seed, 0x11D00001u, mol, 3)) * 3.0 - 1.5;
        const double jy = mdsynth_u01(mdsynth_hash4(m->seed, 0x11D00002u, mol, 3)) * 3.0 - 1.5;
        const double ang = mdsynth_u01(mdsynth_hash4(m->seed, 0x11D00003u, mol, 3)) * 6.283185307179586;
        const double ca = cos(ang), sa = sin(ang);
        const double hx = ((double)ix + 0.5) * 8.0 + jx, hy = ((double)iy + 0.5) * 8.0 + jy;
        const double hz = leaf ? zc + 24.0 : zc - 24.0;   /* head plane; normal points to the centre */
        const double sgn = leaf ? -1.0 : 1.0;
        for (uint32_t b = 0; b < MDSYNTH_LIPID_BEADS; ++b, ++a) {
            const double p[3] = { hx + ca * dxp[b], hy + sa * dxp[b], hz + sgn * dn[b] };
            for (uint32_t k = 0; k < 3; ++k) {
                const float v = (float)p[k];
                if (whole_xyz) whole_xyz[k * N + a] = v;
                if (base_xyz) base_xyz[k * N + a] = mdsynth_wrap(v, L[k]);
            }
            if (mol_id) mol_id[a] = mol;
        }
    }
    const double wsp = (double)m->Lx / (double)m->nw_xy;
    for (uint32_t side = 0; side < 2; ++side)
    for (uint32_t iz = 0; iz < m->nwz; ++iz)
    for (uint32_t iy = 0; iy < m->nw_xy; ++iy)
    for (uint32_t ix = 0; ix < m->nw_xy; ++ix, ++mol, ++a) {
        const double jx = mdsynth_u01(mdsynth_hash4(m->seed, 0x3A7E0001u, mol, 4)) * 1.6 - 0.8;
        const double jy = mdsynth_u01(mdsynth_hash4(m->seed, 0x3A7E0002u, mol, 4)) * 1.6 - 0.8;
        const double jz = mdsynth_u01(mdsynth_hash4(m->seed, 0x3A7E0003u, mol, 4)) * 1.6 - 0.8;
        const double zoff = 26.5 + ((double)iz + 0.5) * 4.5;
        const double p[3] = { ((double)ix + 0.5) * wsp + jx, ((double)iy + 0.5) * wsp + jy, (side ? zc + zoff : zc - zoff) + jz };
        for (uint32_t k = 0; k < 3; ++k) {
            const float v = mdsynth_wrap((float)p[k], L[k]);
            if (whole_xyz) whole_xyz[k * N + a] = v;
            if (base_xyz) base_xyz[k * N + a] = v;
        }
        if (mol_id) mol_id[a] = mol;
    }
}

static inline void mdsynth_membrane_frame(const mdsynth_membrane_t* m, uint32_t frame, const float* base_xyz, const uint32_t* mol_id,
                                          float* x, float* y, float* z) {
    const size_t N = m->num_atoms;
    for (size_t i = 0; i < N; ++i) {
        x[i] = mdsynth_frame_coord(base_xyz[i], m->seed, frame, mol_id[i], 0, m->Lx);
        y[i] = mdsynth_frame_coord(base_xyz[N + i], m->seed, frame, mol_id[i], 1, m->Ly);
        z[i] = mdsynth_frame_coord(base_xyz[2 * N + i], m->seed, frame, mol_id[i], 2, m->Lz);
    }
}

static inline int mdsynth_membrane_write_gro(const mdsynth_membrane_t* m, const char* path, const float* whole_xyz) {
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    const size_t N = m->num_atoms;
    fprintf(f, "synthetic membrane nl=%u nw=%u x %u seed=%u\n%zu\n", m->nl, m->nw_xy, m->nwz, m->seed, N);
    for (size_t i = 0; i < N; ++i) {
        const size_t nlip = (size_t)m->num_lipids * MDSYNTH_LIPID_BEADS;
        const int is_lip = i < nlip;
        const size_t res = is_lip ? i / MDSYNTH_LIPID_BEADS : m->num_lipids + (i - nlip);
        fprintf(f, "%5zu%-5s%5s%5zu%8.3f%8.3f%8.3f\n", (res + 1) % 100000, is_lip ? "LIP" : "SOLW", is_lip ? mdsynth_lipid_names[i % MDSYNTH_LIPID_BEADS] : "W",
                (i + 1) % 100000, whole_xyz[i] * 0.1, whole_xyz[N + i] * 0.1, whole_xyz[2 * N + i] * 0.1);
    }
    fprintf(f, "%10.5f%10.5f%10.5f\n", m->Lx * 0.1, m->Ly * 0.1, m->Lz * 0.1);
    fclose(f);
    return 0;
}
#endif /* MDSYNTH_H */
