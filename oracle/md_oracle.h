/* md_oracle.h — plain-C CPU restatement of the reference hot path (TEST INFRASTRUCTURE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this. The product
 * (viamd_b200/) never links or calls it. Every function cites the reference file:line it follows
 * (paths relative to /root/reference/ext/mdlib/src). Parity status: PINNED — each entry point is
 * checked against outputs of the unmodified reference (oracle/_ref/ref_harness_strict) by
 * tests/test_oracle_vs_ref.py and against the committed vectors in tests/golden/.
 */
#ifndef MD_ORACLE_H
#define MD_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDO_DIST_BINS 1024   /* MD_DIST_BINS md_script_functions.inl:5 */
#define MDO_VOL_DIM   128    /* MD_VOL_DIM   md_script_functions.inl:9 */

enum { MDO_CELL_ORTHO = 1, MDO_CELL_TRICLINIC = 2, MDO_CELL_PBC_X = 4, MDO_CELL_PBC_Y = 8, MDO_CELL_PBC_Z = 16, MDO_CELL_PBC_ALL = 28 };

/* same field order as md_unitcell_t (md_types.h:254-259) */
typedef struct mdo_unitcell_t { double x, xy, xz, y, yz, z; uint32_t flags; } mdo_unitcell_t;

/* rdf(ref, trg, [min,]max): compute_rdf md_script_functions.inl:5263-5338 for atom-index streams.
 * bins/weights: 1024 floats each; returns total pair count. ref_excl_struct: optional, per-ref structure id whose
 * atoms (excl_off/excl_idx CSR) are excluded (rdf_cb_excl_mask :5243) — pass NULL for plain rdf_cb.
 * ref_pos: optional AoS xyz positions replacing ref_idx (COM references). */
uint64_t mdo_rdf_frame(const float* x, const float* y, const float* z,
                       const int32_t* ref_idx, const float* ref_pos_aos, size_t n_ref,
                       const int32_t* trg_idx, size_t n_trg,
                       const mdo_unitcell_t* cell, float min_cutoff, float max_cutoff,
                       const uint32_t* excl_off, const int32_t* excl_idx,
                       float* bins, float* weights);

/* coordinate_extract() for an ARRAY of bitfields (md_script_functions.inl:1496-1507): one centre of mass per group through
 * extract_com :857-874 — sequential float sums in ascending atom order, no periodic treatment. out: AoS xyz [n_groups][3]. */
void mdo_group_com(const float* x, const float* y, const float* z, const float* mass,
                   const int32_t* idx, const uint32_t* off, size_t n_groups, float* out_aos);

/* sdf(ref_structures[], target, cutoff): _sdf md_script_functions.inl:5699-5856.
 * struct_idx: n_struct * struct_size ascending atom indices; conn_*: bond connectivity CSR (md_bond_conn_data_t);
 * init_*: frame-0 coordinates; mass: per atom. vol: 128^3 floats, incremented. Returns number of voxel increments. */
uint64_t mdo_sdf_frame(const float* x, const float* y, const float* z,
                       const float* init_x, const float* init_y, const float* init_z, const float* mass,
                       const int32_t* struct_idx, size_t n_struct, size_t struct_size,
                       const int32_t* trg_idx, size_t n_trg,
                       const uint32_t* conn_off, const int32_t* conn_idx, size_t conn_off_count,
                       const mdo_unitcell_t* cell, float cutoff, float* vol, float* out_matrices /* n_struct*16 or NULL */);

/* density_x/_y/_z: _internal_density md_script_functions.inl:4825-4947 (axis 0..2). init_cell = frame-0 unit cell. */
void mdo_density_frame(const float* x, const float* y, const float* z, const float* mass,
                       const int32_t* idx, size_t n, const mdo_unitcell_t* init_cell, int axis, float* bins, float* weights);

/* distance(a,b) single atom indices: _distance md_script_functions.inl:3851-3890, md_util_deperiodize_vec4 md_util.c:8971 */
float mdo_distance(const float* x, const float* y, const float* z, int32_t a, int32_t b, const mdo_unitcell_t* cell);
/* angle(a,b,c) :4099-4114 and dihedral(a,b,c,d) :4171-4196 for single atom indices */
float mdo_angle(const float* x, const float* y, const float* z, int32_t a, int32_t b, int32_t c);
float mdo_dihedral(const float* x, const float* y, const float* z, int32_t a, int32_t b, int32_t c, int32_t d, const mdo_unitcell_t* cell);

/* Arguments that are selections (or several indices): coordinate_extract_com md_script_functions.inl:1717 -> md_util_com_compute
 * md_util.c:8163 (AVX2 build: 8 float lanes + double remainder; periodic cells use the trigonometric centre of mass _com_pbc_iw :7850
 * with md_mm256_sincos_ps core/md_simd.h:1177). direct != 0 and count == 1: the atom's own position (:1755). */
void mdo_com(const float* x, const float* y, const float* z, const float* mass, const int32_t* idx, size_t count, const mdo_unitcell_t* cell, float out[3]);
void mdo_arg_position(const float* x, const float* y, const float* z, const float* mass, const int32_t* idx, size_t count, int direct,
                      const mdo_unitcell_t* cell, float out[3]);
float mdo_distance_pos(const float a[3], const float b[3], const mdo_unitcell_t* cell);
float mdo_angle_pos(const float a[3], const float b[3], const float c[3]);
float mdo_dihedral_pos(const float p[4][3], const mdo_unitcell_t* cell);

/* XTC frame decode as the reference's trajectory reader does it (md_xtc.c:747-931 with scale 10 nm -> Angstrom, :947-993): coordinates,
 * unit cell (md_unitcell_from_matrix_float), step and time of ONE frame starting at `frame`. Returns 1 on success.
 * mdo_xtc_frame_offsets scans a whole file image for the frame starts (md_xtc.c:436-570): offsets[0..n], returns n. */
int mdo_xtc_decode_frame(const uint8_t* frame, size_t nbytes, size_t num_atoms, float* x, float* y, float* z,
                         mdo_unitcell_t* cell, int32_t* step, float* time);
size_t mdo_xtc_frame_offsets(const uint8_t* file, size_t nbytes, int64_t* offsets, size_t cap);

/* distance_min / distance_max over two atom selections (md_script_functions.inl:3892-3968; both evaluate md_util_min_distance md_util.c:8242) */
float mdo_min_distance(const float* x, const float* y, const float* z, const int32_t* ia, size_t na, const int32_t* ib, size_t nb, const mdo_unitcell_t* cell);

/* building blocks exposed for unit tests */
void mdo_svd3(const float A[3][3], float U[3][3], float S[3][3], float V[3][3]); /* ext/svd3/svd3.c */
uint64_t mdo_count_pairs(const float* x, const float* y, const float* z, const int32_t* ref_idx, size_t n_ref,
                         const int32_t* trg_idx, size_t n_trg, const mdo_unitcell_t* cell, double cell_ext, double cutoff);

#ifdef __cplusplus
}
#endif
#endif
