/* ref_harness.c — drives the UNMODIFIED reference (scanberg/mdlib) through its public API.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under viamd_b200/ may link, import or execute this.
 * It exists so that (a) the plain-C oracle restatement (oracle/md_oracle.c) can be pinned
 * against the real reference, (b) golden vectors under tests/golden/ can be generated, and
 * (c) bench.py can time the reference's own CPU md_script_eval_frame_range on the host cores.
 *
 * It is linked against objects compiled by oracle/Makefile straight from the sources under
 * /root/reference/ext/mdlib (never copied into this repository); outputs go to oracle/_ref/.
 *
 * Calls used (all public, mdlib/src/md_script.h:171-253, md_trajectory.h:49-67):
 *   md_gro_system_init_from_file / md_pdb_system_init_from_file, md_util_system_postprocess,
 *   md_script_ir_create / _compile_from_source, md_script_eval_create / _frame_range /
 *   _property_data / _clear_data.
 * The trajectory is an in-memory md_trajectory_i (no file I/O in the timed region), exactly
 * the shape SURVEY.md §8(d) "CPU timing" prescribes; threads take contiguous frame slices the
 * way VIAMD's enkiTS range task does (reference src/main.cpp:993-997, src/task_system.cpp:73-87).
 *
 * usage:
 *   ref_harness sysinfo --sys F --out O
 *   ref_harness eval    --sys F --traj SPEC --script S --out O [--perframe B:E] [--full B:E] [--threads T]
 *   ref_harness time    --sys F --traj SPEC --script S --frames B:E --threads T [--repeat R]
 * traj SPEC:  raw:<file> | synthwater:<n>:<seed>:<nframes> | sys (trajectory attached by the loader, e.g. multi-model PDB)
 */
#include <md_script.h>
#include <md_system.h>
#include <md_trajectory.h>
#include <md_gro.h>
#include <md_pdb.h>
#include <md_util.h>
#include <md_xtc.h>
#include <xdrfile_xtc.h>
#include <core/md_allocator.h>
#include <core/md_arena_allocator.h>
#include <core/md_str.h>
#include <core/md_os.h>
#include <core/md_log.h>
#include <core/md_bitfield.h>

#include <pthread.h>

#include "harness_common.h"

static void wr(FILE* f, const void* p, size_t n) { if (fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(3); } }
static void wr_u32(FILE* f, uint32_t v) { wr(f, &v, 4); }
static void wr_i64(FILE* f, int64_t v) { wr(f, &v, 8); }
static void wr_u64(FILE* f, uint64_t v) { wr(f, &v, 8); }

/* record: kind(0 perframe,1 full,2 aggregates of a full evaluation) prop beg end  min_value max_value min_range[2] max_range[2]  storage(0 dense,1 sparse) count payload */
static void write_record(FILE* f, uint32_t kind, uint32_t prop, int64_t beg, int64_t end, const md_script_property_data_t* d, const float* vals, size_t n) {
    wr_u32(f, kind); wr_u32(f, prop); wr_i64(f, beg); wr_i64(f, end);
    wr(f, &d->min_value, 4); wr(f, &d->max_value, 4); wr(f, d->min_range, 8); wr(f, d->max_range, 8);
    size_t nnz = 0; for (size_t i = 0; i < n; ++i) nnz += (vals[i] != 0.0f);
    if (n > 65536 && nnz * 2 < n) {
        wr_u32(f, 1); wr_u64(f, nnz);
        for (size_t i = 0; i < n; ++i) if (vals[i] != 0.0f) { uint32_t ii = (uint32_t)i; wr(f, &ii, 4); wr(f, &vals[i], 4); }
    } else { wr_u32(f, 0); wr_u64(f, n); wr(f, vals, n * 4); }
}

typedef struct { md_script_eval_t* eval; const md_script_ir_t* ir; const md_system_t* sys; const md_trajectory_i* traj; uint32_t beg, end; bool ok; } job_t;
static void* job_main(void* p) { job_t* j = p; j->ok = md_script_eval_frame_range(j->eval, j->ir, j->sys, j->traj, j->beg, j->end); return NULL; }

static bool run_threads(md_script_eval_t* eval, const md_script_ir_t* ir, const md_system_t* sys, const md_trajectory_i* traj, long beg, long end, int T) {
    if (T <= 1) return md_script_eval_frame_range(eval, ir, sys, traj, (uint32_t)beg, (uint32_t)end);
    pthread_t* th = calloc(T, sizeof(pthread_t)); job_t* jobs = calloc(T, sizeof(job_t));
    long n = end - beg; bool ok = true;
    for (int t = 0; t < T; ++t) {
        jobs[t] = (job_t){ eval, ir, sys, traj, (uint32_t)(beg + n * t / T), (uint32_t)(beg + n * (t + 1) / T), false };
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) { pthread_join(th[t], NULL); ok = ok && jobs[t].ok; }
    free(th); free(jobs); return ok;
}

/* ---------------------------------------------------------------- modes */

static int mode_sysinfo(int argc, char** argv) {
    md_allocator_i* alloc = md_vm_arena_create(GIGABYTES(8));
    md_system_t sys; if (!load_system(&sys, arg_val(argc, argv, "--sys", ""), alloc)) return 2;
    FILE* f = fopen(arg_val(argc, argv, "--out", "sysinfo.bin"), "wb"); if (!f) return 2;
    const size_t n = sys.atom.count;
    wr(f, "MDSYSINF", 8); wr_u64(f, n);
    float* mass = malloc(n * 4); md_atom_extract_masses(mass, 0, n, &sys.atom); wr(f, mass, n * 4);
    for (size_t i = 0; i < n; ++i) { uint32_t z = md_atom_atomic_number(&sys.atom, i); wr_u32(f, z); }
    for (size_t i = 0; i < n; ++i) { str_t s = md_atom_name(&sys.atom, i); char b[8] = {0}; memcpy(b, s.ptr, s.len < 8 ? s.len : 7); wr(f, b, 8); }
    /* component (residue) atom offsets */
    wr_u64(f, sys.component.count);
    wr(f, sys.component.atom_offset, (sys.component.count + 1) * 4);
    /* bond connectivity CSR */
    wr_u64(f, sys.bond.conn.offset_count); wr_u64(f, sys.bond.conn.count);
    wr(f, sys.bond.conn.offset, sys.bond.conn.offset_count * 4);
    wr(f, sys.bond.conn.atom_idx, sys.bond.conn.count * 4);
    wr(f, &sys.unitcell, sizeof(md_unitcell_t));
    wr(f, sys.atom.x, n * 4); wr(f, sys.atom.y, n * 4); wr(f, sys.atom.z, n * 4);
    fclose(f);
    printf("{\"atoms\": %zu, \"components\": %zu, \"bonds\": %zu}\n", n, sys.component.count, sys.bond.count);
    return 0;
}

static int mode_eval(int argc, char** argv, bool timing) {
    md_allocator_i* alloc = md_vm_arena_create(GIGABYTES(16));
    md_system_t sys; if (!load_system(&sys, arg_val(argc, argv, "--sys", ""), alloc)) return 2;
    md_trajectory_i traj = {0}; mem_traj_t mt;
    if (!make_traj(&traj, &mt, arg_val(argc, argv, "--traj", "sys"), &sys)) return 2;
    const size_t num_frames = md_trajectory_num_frames(&traj);
    if (md_trajectory_num_atoms(&traj) != sys.atom.count) { fprintf(stderr, "atom count mismatch traj %zu sys %zu\n", md_trajectory_num_atoms(&traj), sys.atom.count); return 2; }

    const char* src = arg_val(argc, argv, "--script", "");
    md_script_ir_t* ir = md_script_ir_create(alloc);
    if (!md_script_ir_compile_from_source(ir, (str_t){ src, strlen(src) }, &sys, &traj, NULL) || !md_script_ir_valid(ir)) {
        fprintf(stderr, "script failed to compile: %s\n", src);
        const md_log_token_t* e = md_script_ir_errors(ir);
        for (size_t i = 0; i < md_script_ir_num_errors(ir); ++i) fprintf(stderr, "  error: %.*s\n", (int)e[i].text.len, e[i].text.ptr);
        return 2;
    }
    const int T = atoi(arg_val(argc, argv, "--threads", "1"));
    md_script_eval_t* eval = md_script_eval_create(num_frames, ir, alloc);
    if (!eval) { fprintf(stderr, "no properties in script\n"); return 2; }
    const size_t np = md_script_ir_property_count(ir);
    const str_t* names = md_script_ir_property_names(ir);

    if (timing) {
        long b = 0, e = (long)num_frames; parse_range(arg_val(argc, argv, "--frames", NULL), &b, &e);
        const int R = atoi(arg_val(argc, argv, "--repeat", "1"));
        const int W = atoi(arg_val(argc, argv, "--warmup", "1"));   /* untimed full passes inside this process (page faults, arena commits, caches) */
        if (traj.inst == (struct md_trajectory_o*)&mt) mt_materialize(&mt, (size_t)b, (size_t)e);   /* the frames are in memory before the clock starts */
        for (int w = 0; w < W; ++w) { md_script_eval_clear_data(eval); run_threads(eval, ir, &sys, &traj, b, e, T); }
        double best = 1e300, sum = 0; double* times = calloc(R > 0 ? R : 1, sizeof(double));
        for (int r = 0; r < R; ++r) {
            md_script_eval_clear_data(eval);
            double t0 = now_s(); bool ok = run_threads(eval, ir, &sys, &traj, b, e, T); double dt = now_s() - t0;
            if (!ok) { fprintf(stderr, "evaluation failed\n"); return 2; }
            if (dt < best) best = dt; sum += dt; times[r] = dt;
        }
        double chk = 0; for (size_t p = 0; p < np; ++p) { const md_script_property_data_t* d = md_script_eval_property_data(eval, names[p]); for (size_t i = 0; i < d->num_values; ++i) chk += d->values[i]; }
        printf("{\"frames\": %ld, \"threads\": %d, \"repeat\": %d, \"warmup\": %d, \"best_s\": %.6f, \"mean_s\": %.6f, \"frames_per_s\": %.3f, \"frames_per_s_mean\": %.3f, \"checksum\": %.6f, \"times_s\": [",
               e - b, T, R, W, best, sum / R, (double)(e - b) / best, (double)(e - b) * R / sum, chk);
        for (int r = 0; r < R; ++r) printf("%s%.6f", r ? ", " : "", times[r]);
        printf("]}\n");
        return 0;
    }

    FILE* f = fopen(arg_val(argc, argv, "--out", "refout.bin"), "wb"); if (!f) return 2;
    long pb = 0, pe = 0, fb = 0, fe = 0;
    const bool perframe = parse_range(arg_val(argc, argv, "--perframe", NULL), &pb, &pe);
    const bool full = parse_range(arg_val(argc, argv, "--full", NULL), &fb, &fe);
    wr(f, "MDREFOUT", 8); wr_u32(f, 1); wr_u32(f, (uint32_t)np);
    for (size_t p = 0; p < np; ++p) {
        const md_script_property_data_t* d = md_script_eval_property_data(eval, names[p]);
        char nm[64] = {0}; memcpy(nm, names[p].ptr, names[p].len < 63 ? names[p].len : 63); wr(f, nm, 64);
        wr_u32(f, (uint32_t)md_script_ir_property_flags(ir, names[p])); wr(f, d->dim, 16); wr_u64(f, d->num_values);
    }
    if (perframe) {
        /* raw per-frame outputs: a cleared eval + a single-frame range makes the cumulative
         * moving average return the frame's own values (n = 0), md_script.c:5912-5921 */
        for (long fr = pb; fr < pe; ++fr) {
            md_script_eval_clear_data(eval);
            if (!md_script_eval_frame_range(eval, ir, &sys, &traj, (uint32_t)fr, (uint32_t)fr + 1)) { fprintf(stderr, "frame %ld failed\n", fr); return 2; }
            for (size_t p = 0; p < np; ++p) {
                const md_script_property_data_t* d = md_script_eval_property_data(eval, names[p]);
                const uint32_t fl = (uint32_t)md_script_ir_property_flags(ir, names[p]);
                if (fl & MD_SCRIPT_PROPERTY_FLAG_TEMPORAL) {
                    const size_t w = d->num_values / num_frames;
                    write_record(f, 0, (uint32_t)p, fr, fr + 1, d, d->values + (size_t)fr * w, w);
                } else write_record(f, 0, (uint32_t)p, fr, fr + 1, d, d->values, d->num_values);
            }
        }
    }
    if (full) {
        md_script_eval_clear_data(eval);
        if (!run_threads(eval, ir, &sys, &traj, fb, fe, T)) { fprintf(stderr, "full evaluation failed\n"); return 2; }
        for (size_t p = 0; p < np; ++p) {
            const md_script_property_data_t* d = md_script_eval_property_data(eval, names[p]);
            write_record(f, 1, (uint32_t)p, fb, fe, d, d->values, d->num_values);
            if (d->aggregate && d->aggregate->num_values) {   /* kind 2: per-frame mean | variance | (min,max) of a multi-valued temporal (md_script.c:5886-5890) */
                const size_t na = d->aggregate->num_values; float* agg = malloc(na * 16);
                memcpy(agg, d->aggregate->population_mean, na * 4); memcpy(agg + na, d->aggregate->population_var, na * 4);
                memcpy(agg + 2 * na, d->aggregate->population_ext, na * 8);
                wr_u32(f, 2); wr_u32(f, (uint32_t)p); wr_i64(f, fb); wr_i64(f, fe);
                wr(f, &d->min_value, 4); wr(f, &d->max_value, 4); wr(f, d->min_range, 8); wr(f, d->max_range, 8);
                wr_u32(f, 0); wr_u64(f, na * 4); wr(f, agg, na * 16); free(agg);
            }
        }
    }
    fclose(f);
    printf("{\"properties\": %zu, \"frames\": %zu}\n", np, num_frames);
    return 0;
}

/* dumptraj: write frames [B,E) of any trajectory spec as an MDRAWTRJ container (fixture generation) */
static int mode_dumptraj(int argc, char** argv) {
    md_allocator_i* alloc = md_vm_arena_create(GIGABYTES(8));
    md_system_t sys; if (!load_system(&sys, arg_val(argc, argv, "--sys", ""), alloc)) return 2;
    md_trajectory_i traj = {0}; mem_traj_t mt;
    if (!make_traj(&traj, &mt, arg_val(argc, argv, "--traj", "sys"), &sys)) return 2;
    long b = 0, e = (long)md_trajectory_num_frames(&traj); parse_range(arg_val(argc, argv, "--frames", NULL), &b, &e);
    const size_t n = sys.atom.count;
    FILE* f = fopen(arg_val(argc, argv, "--out", "traj.raw"), "wb"); if (!f) return 2;
    wr(f, "MDRAWTRJ", 8); wr_u64(f, (uint64_t)(e - b)); wr_u64(f, n);
    float* xyz = malloc(n * 12);
    md_trajectory_reader_i rd = {0}; md_trajectory_reader_init(&rd, &traj);
    for (long fr = b; fr < e; ++fr) {
        md_trajectory_frame_header_t h = {0};
        if (!md_trajectory_reader_load_frame(rd, fr, &h, xyz, xyz + n, xyz + 2 * n)) return 2;
        double cell[6] = { h.unitcell.x, h.unitcell.xy, h.unitcell.xz, h.unitcell.y, h.unitcell.yz, h.unitcell.z };
        uint32_t fl[2] = { (uint32_t)h.unitcell.flags, 0 };
        wr(f, cell, 48); wr(f, fl, 8); wr(f, xyz, n * 12);
    }
    md_trajectory_reader_free(&rd); fclose(f);
    return 0;
}

/* shapespace: the per-frame loop body of VIAMD's shape-space component (src/components/shapespace/shapespace.cpp:404-431) and of
 * _shape_weights (md_script_functions.inl:6005-6050), written against the reference's own functions: for every frame and every structure
 * (here: components [R0,R1) of the system, i.e. residues) xyzw -> md_util_com_compute_vec4 (periodic) -> md_util_deperiodize_vec4 ->
 * mat3_covariance_matrix_vec4 -> md_util_shape_weights. Output: MDSHAPES, frames, structures, then weights[frame][structure][3]. */
static int mode_shapespace(int argc, char** argv) {
    md_allocator_i* alloc = md_vm_arena_create(GIGABYTES(8));
    md_system_t sys; if (!load_system(&sys, arg_val(argc, argv, "--sys", ""), alloc)) return 2;
    md_trajectory_i traj = {0}; mem_traj_t mt;
    if (!make_traj(&traj, &mt, arg_val(argc, argv, "--traj", "sys"), &sys)) return 2;
    long b = 0, e = (long)md_trajectory_num_frames(&traj); parse_range(arg_val(argc, argv, "--frames", NULL), &b, &e);
    long r0 = 0, r1 = (long)sys.component.count; parse_range(arg_val(argc, argv, "--res", NULL), &r0, &r1);
    const bool use_mass = atoi(arg_val(argc, argv, "--mass", "1")) != 0;
    const size_t n = sys.atom.count;
    float* xyz = malloc(n * 12); float* x = xyz; float* y = xyz + n; float* z = xyz + 2 * n;
    float* w = malloc(n * 4); md_atom_extract_masses(w, 0, n, &sys.atom);
    vec4_t* xyzw = malloc(n * sizeof(vec4_t));
    FILE* f = fopen(arg_val(argc, argv, "--out", "shapes.bin"), "wb"); if (!f) return 2;
    wr(f, "MDSHAPES", 8); wr_u64(f, (uint64_t)(e - b)); wr_u64(f, (uint64_t)(r1 - r0));
    md_trajectory_reader_i rd = {0}; md_trajectory_reader_init(&rd, &traj);
    for (long fr = b; fr < e; ++fr) {
        md_trajectory_frame_header_t h = {0};
        if (!md_trajectory_reader_load_frame(rd, fr, &h, x, y, z)) return 2;
        for (long r = r0; r < r1; ++r) {
            const size_t a0 = sys.component.atom_offset[r], a1 = sys.component.atom_offset[r + 1], count = a1 - a0;
            for (size_t k = 0; k < count; ++k) xyzw[k] = vec4_set(x[a0 + k], y[a0 + k], z[a0 + k], use_mass ? w[a0 + k] : 1.0f);
            const vec3_t com = md_util_com_compute_vec4(xyzw, 0, count, &h.unitcell);
            md_util_deperiodize_vec4(xyzw, count, com, &h.unitcell);
            const mat3_t M = mat3_covariance_matrix_vec4(xyzw, 0, count, com);
            const vec3_t weights = md_util_shape_weights(&M);
            wr(f, &weights, 12);
        }
    }
    md_trajectory_reader_free(&rd); fclose(f);
    return 0;
}

/* xtcwrite: frames [B,E) of any trajectory spec -> an .xtc file through the reference's bundled xdrfile writer (ext/xtc/xdrfile_xtc.c:
 * write_xtc), coordinates Angstrom -> nm. Fixture generation for the XTC decode path. */
/* backbone angles per frame through the reference's md_util_backbone_angles_compute (VIAMD's "Backbone Operations" pass, src/viamd.cpp:488-520):
 * MDBACKBN | u64 F | u64 nseg | i32 atoms[nseg][5] = C(i-1), N, CA, C, N(i+1) (-1: the segment has no angles) | f32 angles[F][nseg][2] */
static int mode_backbone(int argc, char** argv) {
    md_allocator_i* alloc = md_vm_arena_create(GIGABYTES(8));
    md_system_t sys; if (!load_system(&sys, arg_val(argc, argv, "--sys", ""), alloc)) return 2;
    md_trajectory_i traj = {0}; mem_traj_t mt;
    if (!make_traj(&traj, &mt, arg_val(argc, argv, "--traj", "sys"), &sys)) return 2;
    long b = 0, e = (long)md_trajectory_num_frames(&traj); parse_range(arg_val(argc, argv, "--frames", NULL), &b, &e);
    const md_protein_backbone_data_t* bb = &sys.protein_backbone;
    const size_t nseg = bb->segment.count; if (!nseg) { fprintf(stderr, "system has no protein backbone\n"); return 2; }
    FILE* f = fopen(arg_val(argc, argv, "--out", "backbone.bin"), "wb"); if (!f) return 2;
    wr(f, "MDBACKBN", 8); wr_u64(f, (uint64_t)(e - b)); wr_u64(f, (uint64_t)nseg);
    int32_t* five = calloc(nseg * 5, sizeof(int32_t));
    for (size_t i = 0; i < nseg * 5; ++i) five[i] = -1;
    for (size_t r = 0; r < bb->range.count; ++r) {   /* the loop bounds of md_util_backbone_angles_compute (md_util.c:2584-2592) */
        const size_t rb = bb->range.offset[r], re = bb->range.offset[r + 1];
        if (re - rb < 4) continue;
        for (size_t i = rb + 1; i + 1 < re; ++i) {
            five[5 * i + 0] = bb->segment.atoms[i - 1].c; five[5 * i + 1] = bb->segment.atoms[i].n; five[5 * i + 2] = bb->segment.atoms[i].ca;
            five[5 * i + 3] = bb->segment.atoms[i].c; five[5 * i + 4] = bb->segment.atoms[i + 1].n;
        }
    }
    wr(f, five, nseg * 5 * sizeof(int32_t));
    const size_t n = sys.atom.count; float* x = malloc(n * 4); float* y = malloc(n * 4); float* z = malloc(n * 4);
    md_backbone_angles_t* ang = calloc(nseg, sizeof(md_backbone_angles_t));
    for (long fr = b; fr < e; ++fr) {
        md_trajectory_frame_header_t hdr = {0};
        if (!md_trajectory_load_frame(&traj, fr, &hdr, x, y, z)) return 2;
        md_util_backbone_angles_compute(ang, nseg, x, y, z, &hdr.unitcell, bb);
        wr(f, ang, nseg * sizeof(md_backbone_angles_t));
    }
    fclose(f);
    printf("{\"frames\": %ld, \"segments\": %zu}\n", e - b, nseg);
    return 0;
}

static int mode_xtcwrite(int argc, char** argv) {
    md_allocator_i* alloc = md_vm_arena_create(GIGABYTES(8));
    md_system_t sys; if (!load_system(&sys, arg_val(argc, argv, "--sys", ""), alloc)) return 2;
    md_trajectory_i traj = {0}; mem_traj_t mt;
    if (!make_traj(&traj, &mt, arg_val(argc, argv, "--traj", "sys"), &sys)) return 2;
    long b = 0, e = (long)md_trajectory_num_frames(&traj); parse_range(arg_val(argc, argv, "--frames", NULL), &b, &e);
    const float prec = (float)atof(arg_val(argc, argv, "--precision", "1000"));
    const size_t n = sys.atom.count;
    XDRFILE* xd = xdrfile_open(arg_val(argc, argv, "--out", "traj.xtc"), "w"); if (!xd) return 2;
    float* xyz = malloc(n * 12); rvec* r = malloc(n * sizeof(rvec));
    md_trajectory_reader_i rd = {0}; md_trajectory_reader_init(&rd, &traj);
    for (long fr = b; fr < e; ++fr) {
        md_trajectory_frame_header_t h = {0};
        if (!md_trajectory_reader_load_frame(rd, fr, &h, xyz, xyz + n, xyz + 2 * n)) return 2;
        for (size_t i = 0; i < n; ++i) { r[i][0] = xyz[i] * 0.1f; r[i][1] = xyz[n + i] * 0.1f; r[i][2] = xyz[2 * n + i] * 0.1f; }
        matrix box = { { (float)(h.unitcell.x * 0.1), 0, 0 }, { (float)(h.unitcell.xy * 0.1), (float)(h.unitcell.y * 0.1), 0 },
                       { (float)(h.unitcell.xz * 0.1), (float)(h.unitcell.yz * 0.1), (float)(h.unitcell.z * 0.1) } };
        if (write_xtc(xd, (int)n, (int)fr, (float)fr, box, r, prec) != exdrOK) return 2;
    }
    md_trajectory_reader_free(&rd); xdrfile_close(xd);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "xtcwrite") == 0) return mode_xtcwrite(argc, argv);
    if (argc < 2) { fprintf(stderr, "usage: ref_harness sysinfo|eval|time|dumptraj ...\n"); return 1; }
    if (strcmp(argv[1], "dumptraj") == 0) return mode_dumptraj(argc, argv);
    if (strcmp(argv[1], "shapespace") == 0) return mode_shapespace(argc, argv);
    if (strcmp(argv[1], "backbone") == 0) return mode_backbone(argc, argv);
    if (strcmp(argv[1], "sysinfo") == 0) return mode_sysinfo(argc, argv);
    if (strcmp(argv[1], "eval") == 0) return mode_eval(argc, argv, false);
    if (strcmp(argv[1], "time") == 0) return mode_eval(argc, argv, true);
    fprintf(stderr, "unknown mode %s\n", argv[1]);
    return 1;
}
