/* shim_harness.c — TEST INFRASTRUCTURE: proves the drop-in boundary end to end.
 *
 * This translation unit is the reference's md_script.c (included verbatim from /root/reference, exactly as its own white-box tests do,
 * mdlib/unittest/test_script.c:20) followed by integration/md_script_mdgpu.inl — i.e. what a maintainer's build would contain.
 *   shim_harness lower --sys F --script S --out O
 *        compile S with the UNMODIFIED md_script front-end, lower the IR with the shim, dump the descriptors (no GPU needed)
 *   shim_harness eval  --sys F --traj SPEC --script S [--frames B:E]
 *        evaluate S twice through the md_script API: md_script_eval_frame_range (reference CPU path) and
 *        md_script_gpu_eval_frame_range (libmdgpu), then compare the md_script_property_data_t contents. Exit code 0 = parity.
 *   shim_harness dropin --sys F --traj SPEC --script S [--threads T] [--chunk C] [--interrupt-at K]
 *        VIAMD's call pattern, unmodified: T threads pull disjoint frame ranges of C frames and call the PUBLIC md_script_eval_frame_range
 *        on ONE eval (src/main.cpp:993-997 through src/task_system.cpp:73-87; mdlib/unittest/test_script.c:1352-1417) — which is the
 *        dispatcher of integration/md_script_mdgpu.inl — while the main thread polls the frame mask as the UI does; the result is compared with
 *        the reference's own md_script_eval_frame_range (`__cpu` in this TU). Then md_script_eval_interrupt mid-run, clear, re-evaluate.
 */
#include "../integration/md_script_mdgpu_pre.h"   /* exactly integration/md_script_mdgpu.c: pre.h, the reference's md_script.c, the .inl */
#include <md_script.c>
#include <md_gro.h>
#include <md_pdb.h>
#include "harness_common.h"
#include "../integration/md_script_mdgpu.inl"
#include <pthread.h>

/* relative tolerance of the value comparisons; --tol overrides it for runs of thousands of frames, where the reference's own float moving average
 * (md_script.c:5912) has drifted 1 - 3e-5 from the exact mean of the same per-frame values */
static double g_rel_tol = 1e-5;
static int mode_lower(int argc, char** argv) {
    md_allocator_i* alloc = md_vm_arena_create(GIGABYTES(8));
    md_system_t sys; if (!load_system(&sys, arg_val(argc, argv, "--sys", ""), alloc)) return 2;
    const char* src = arg_val(argc, argv, "--script", "");
    md_script_ir_t* ir = md_script_ir_create(alloc);
    if (!md_script_ir_compile_from_source(ir, (str_t){ src, strlen(src) }, &sys, NULL, NULL) || !md_script_ir_valid(ir)) { fprintf(stderr, "script failed to compile\n"); return 2; }
    md_script_gpu_lowered_t low = {0};
    if (!md_script_gpu_lower_sys(&low, ir, &sys, alloc)) return 3;
    FILE* f = fopen(arg_val(argc, argv, "--out", "lowered.bin"), "wb"); if (!f) return 2;
    uint64_t np = low.num_props; fwrite("MDLOWER3", 1, 8, f); fwrite(&np, 8, 1, f);
    for (size_t i = 0; i < low.num_props; ++i) {
        const mdgpu_property_desc_t* p = &low.props[i];
        uint64_t v[3] = { p->op, p->num_structures, p->structure_size };
        fwrite(low.names[i], 1, 64, f); fwrite(v, 8, 3, f); fwrite(&p->cutoff_min, 4, 1, f); fwrite(&p->cutoff_max, 4, 1, f);
        for (int k = 0; k < 4; ++k) { uint64_t c = p->idx_count[k]; fwrite(&c, 8, 1, f); if (c) fwrite(p->idx[k], 4, c, f); }
        for (int k = 0; k < 4; ++k) {   /* dynamic arguments; rdf's round-1 spelling of a dynamic reference is reported as dyn[0] too */
            mdgpu_dynamic_arg_t dy = p->dyn[k];
            if (k == 0 && p->op == MDGPU_OP_RDF && p->ref_within_radius > 0.0f) { dy.radius_min = p->ref_within_min; dy.radius_max = p->ref_within_radius; dy.has_and = p->com_args & 1u; dy.and_idx = p->idx[2]; dy.and_count = p->idx_count[2]; }
            uint64_t c = dy.has_and ? dy.and_count : 0, h = dy.has_and;
            fwrite(&dy.radius_min, 4, 1, f); fwrite(&dy.radius_max, 4, 1, f); fwrite(&h, 8, 1, f); fwrite(&c, 8, 1, f); if (c) fwrite(dy.and_idx, 4, c, f);
        }
        { uint64_t nb = p->num_structures_b; fwrite(&nb, 8, 1, f); if (nb && p->structure_offsets_b) fwrite(p->structure_offsets_b, 4, nb + 1, f); }
        for (int k = 0; k < 4; ++k) { uint64_t np_k = p->arg_offsets[k] ? p->arg_parts[k] : 0; fwrite(&np_k, 8, 1, f); if (np_k) fwrite(p->arg_offsets[k], 4, np_k + 1, f); }   /* arrays of selections as one argument */
    }
    fclose(f);
    printf("{\"properties\": %zu}\n", low.num_props);
    return 0;
}

static int mode_eval(int argc, char** argv) {
    md_allocator_i* alloc = md_vm_arena_create(GIGABYTES(16));
    md_system_t sys; if (!load_system(&sys, arg_val(argc, argv, "--sys", ""), alloc)) return 2;
    md_trajectory_i traj = {0}; mem_traj_t mt;
    if (!make_traj(&traj, &mt, arg_val(argc, argv, "--traj", "sys"), &sys)) return 2;
    const size_t num_frames = md_trajectory_num_frames(&traj);
    const char* src = arg_val(argc, argv, "--script", "");
    md_script_ir_t* ir = md_script_ir_create(alloc);
    if (!md_script_ir_compile_from_source(ir, (str_t){ src, strlen(src) }, &sys, &traj, NULL) || !md_script_ir_valid(ir)) { fprintf(stderr, "script failed to compile\n"); return 2; }
    long b = 0, e = (long)num_frames; parse_range(arg_val(argc, argv, "--frames", NULL), &b, &e);

    md_script_eval_t* cpu = md_script_eval_create(num_frames, ir, alloc);
    md_script_eval_t* gpu = md_script_eval_create(num_frames, ir, alloc);
    if (!cpu || !gpu) return 2;
    double t0 = now_s();
    if (!md_script_eval_frame_range__cpu(cpu, ir, &sys, &traj, (uint32_t)b, (uint32_t)e)) { fprintf(stderr, "reference evaluation failed\n"); return 2; }
    double t_cpu = now_s() - t0;

    mdgpu_plan* plan = md_script_gpu_plan_create(ir, &sys, num_frames, 0, alloc);
    if (!plan) { fprintf(stderr, "plan creation failed: %s\n", mdgpu_last_error()); return 4; }
    t0 = now_s();
    if (!md_script_gpu_eval_frame_range(plan, gpu, ir, &traj, (uint32_t)b, (uint32_t)e, 4)) return 4;
    double t_gpu = now_s() - t0;

    int bad = 0;
    const size_t np = md_script_ir_property_count(ir); const str_t* names = md_script_ir_property_names(ir);
    printf("{\"frames\": %ld, \"cpu_s\": %.4f, \"gpu_s\": %.4f, \"properties\": [", e - b, t_cpu, t_gpu);
    for (size_t p = 0; p < np; ++p) {
        const md_script_property_data_t* a = md_script_eval_property_data(cpu, names[p]);
        const md_script_property_data_t* g = md_script_eval_property_data(gpu, names[p]);
        double maxrel = 0, maxabs = 0; size_t nbad = 0;
        for (size_t i = 0; i < a->num_values; ++i) {
            const double d = fabs((double)a->values[i] - (double)g->values[i]);
            const double tol = g_rel_tol * fabs((double)a->values[i]) + 1e-6;   /* north_star: 1e-5 relative for averaged floats (--tol) */
            if (d > maxabs) maxabs = d;
            if (fabs((double)a->values[i]) > 0 && d / fabs((double)a->values[i]) > maxrel) maxrel = d / fabs((double)a->values[i]);
            if (d > tol) nbad++;
        }
        const bool mask_ok = md_bitfield_popcount(&cpu->frame_mask) == md_bitfield_popcount(&gpu->frame_mask);
        if (nbad || !mask_ok) bad = 1;
        printf("%s{\"name\": \"%.*s\", \"num_values\": %zu, \"max_abs\": %.3g, \"max_rel\": %.3g, \"out_of_tol\": %zu, \"frame_mask_equal\": %s}",
               p ? ", " : "", (int)names[p].len, names[p].ptr, a->num_values, maxabs, maxrel, nbad, mask_ok ? "true" : "false");
    }
    printf("], \"parity\": %s}\n", bad ? "false" : "true");
    mdgpu_plan_destroy(plan);
    return bad ? 5 : 0;
}


/* ---- drop-in mode: the public API, called the way VIAMD calls it ---- */
typedef struct dropin_job_t {
    md_script_eval_t* eval; const md_script_ir_t* ir; const md_system_t* sys; const md_trajectory_i* traj;
    uint32_t num_frames, chunk; volatile uint32_t next; volatile int failed; volatile int finished;
} dropin_job_t;

static void* dropin_worker(void* arg) {
    dropin_job_t* j = (dropin_job_t*)arg;
    for (;;) {
        const uint32_t beg = __sync_fetch_and_add(&j->next, j->chunk);
        if (beg >= j->num_frames) break;
        const uint32_t end = beg + j->chunk < j->num_frames ? beg + j->chunk : j->num_frames;
        if (!md_script_eval_frame_range(j->eval, j->ir, j->sys, j->traj, beg, end)) j->failed = 1;   /* the PUBLIC entry point */
    }
    __sync_fetch_and_add(&j->finished, 1);
    return NULL;
}

/* runs the job on T threads; the calling thread polls the frame mask like VIAMD's UI does; returns the number of distinct partial counts seen */
static int dropin_run(dropin_job_t* j, int T, long interrupt_at, size_t* out_final) {
    pthread_t th[64]; if (T > 64) T = 64;
    j->next = 0; j->failed = 0; j->finished = 0;
    for (int t = 0; t < T; ++t) pthread_create(&th[t], NULL, dropin_worker, j);
    int distinct = 0; size_t last = 0; bool interrupted = false;
    while (j->finished < T) {
        md_mutex_lock(&j->eval->frame_lock); const size_t done = md_bitfield_popcount(&j->eval->frame_mask); md_mutex_unlock(&j->eval->frame_lock);
        if (done != last) { if (done < j->num_frames) distinct++; last = done; }
        if (interrupt_at >= 0 && !interrupted && done >= (size_t)interrupt_at) { md_script_eval_interrupt(j->eval); interrupted = true; }
        struct timespec ts = { 0, 200000 }; nanosleep(&ts, NULL);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
    md_mutex_lock(&j->eval->frame_lock); *out_final = md_bitfield_popcount(&j->eval->frame_mask); md_mutex_unlock(&j->eval->frame_lock);
    return distinct;
}

static int compare_evals(const md_script_ir_t* ir, md_script_eval_t* cpu, md_script_eval_t* gpu, const char* tag) {
    int bad = 0;
    const size_t np = md_script_ir_property_count(ir); const str_t* names = md_script_ir_property_names(ir);
    printf("\"%s\": [", tag);
    for (size_t p = 0; p < np; ++p) {
        const md_script_property_data_t* a = md_script_eval_property_data(cpu, names[p]);
        const md_script_property_data_t* g = md_script_eval_property_data(gpu, names[p]);
        double maxrel = 0, maxabs = 0; size_t nbad = 0;
        for (size_t i = 0; i < a->num_values; ++i) {
            const double d = fabs((double)a->values[i] - (double)g->values[i]);
            const double tol = g_rel_tol * fabs((double)a->values[i]) + 1e-6;
            if (d > maxabs) maxabs = d;
            if (fabs((double)a->values[i]) > 0 && d / fabs((double)a->values[i]) > maxrel) maxrel = d / fabs((double)a->values[i]);
            if (d > tol) nbad++;
        }
        const bool scal_ok = fabs((double)a->min_value - (double)g->min_value) <= 1e-5 * fabs((double)a->min_value) + 1e-6 && fabs((double)a->max_value - (double)g->max_value) <= 1e-5 * fabs((double)a->max_value) + 1e-6;
        if (nbad || !scal_ok || a->fingerprint == 0 || g->fingerprint == 0) bad = 1;
        printf("%s{\"name\": \"%.*s\", \"max_abs\": %.3g, \"max_rel\": %.3g, \"out_of_tol\": %zu, \"min_max_equal\": %s}", p ? ", " : "", (int)names[p].len, names[p].ptr, maxabs, maxrel, nbad, scal_ok ? "true" : "false");
    }
    printf("], ");
    return bad;
}

static int mode_dropin(int argc, char** argv) {
    md_allocator_i* alloc = md_vm_arena_create(GIGABYTES(16));
    md_system_t sys; if (!load_system(&sys, arg_val(argc, argv, "--sys", ""), alloc)) return 2;
    md_trajectory_i traj = {0}; mem_traj_t mt;
    if (!make_traj(&traj, &mt, arg_val(argc, argv, "--traj", "sys"), &sys)) return 2;
    const size_t num_frames = md_trajectory_num_frames(&traj);
    const char* src = arg_val(argc, argv, "--script", "");
    md_script_ir_t* ir = md_script_ir_create(alloc);
    if (!md_script_ir_compile_from_source(ir, (str_t){ src, strlen(src) }, &sys, &traj, NULL) || !md_script_ir_valid(ir)) { fprintf(stderr, "script failed to compile\n"); return 2; }
    const int T = atoi(arg_val(argc, argv, "--threads", "4"));
    long chunk = atol(arg_val(argc, argv, "--chunk", "0")); if (chunk <= 0) chunk = (long)(num_frames / (size_t)(T * (T > 1 ? T - 1 : 1))); if (chunk < 1) chunk = 1;
    const long interrupt_at = atol(arg_val(argc, argv, "--interrupt-at", "-1"));
    g_rel_tol = atof(arg_val(argc, argv, "--tol", "1e-5"));

    md_script_eval_t* cpu = md_script_eval_create(num_frames, ir, alloc);
    md_script_eval_t* gpu = md_script_eval_create(num_frames, ir, alloc);
    if (!cpu || !gpu) return 2;
    double t0 = now_s();
    if (!md_script_eval_frame_range__cpu(cpu, ir, &sys, &traj, 0, (uint32_t)num_frames)) { fprintf(stderr, "reference evaluation failed\n"); return 2; }
    const double t_cpu = now_s() - t0;

    dropin_job_t job = { gpu, ir, &sys, &traj, (uint32_t)num_frames, (uint32_t)chunk, 0, 0, 0 };
    size_t done = 0;
    t0 = now_s();
    const int partial_views = dropin_run(&job, T, -1, &done);
    const double t_gpu = now_s() - t0;
    int bad = job.failed || done != num_frames;
    printf("{\"frames\": %zu, \"threads\": %d, \"chunk\": %ld, \"cpu_s\": %.4f, \"dropin_s\": %.4f, \"partial_mask_views\": %d, \"frames_done\": %zu, ", num_frames, T, chunk, t_cpu, t_gpu, partial_views, done);
    bad |= compare_evals(ir, cpu, gpu, "properties");

    /* interrupt mid-run (md_script_eval_interrupt :6663), then VIAMD's restart sequence: clear + evaluate again (src/main.cpp:982-997) */
    size_t after_interrupt = num_frames, after_restart = 0; int bad2 = 0;
    if (interrupt_at >= 0) {
        md_script_eval_clear_data(gpu);
        dropin_run(&job, T, interrupt_at, &after_interrupt);
        md_script_eval_clear_data(gpu);
        md_mutex_lock(&gpu->frame_lock); const size_t cleared = md_bitfield_popcount(&gpu->frame_mask); md_mutex_unlock(&gpu->frame_lock);
        dropin_run(&job, T, -1, &after_restart);
        bad2 = job.failed || cleared != 0 || after_restart != num_frames;
        bad2 |= compare_evals(ir, cpu, gpu, "after_restart");
        printf("\"frames_done_at_interrupt\": %zu, \"frames_done_after_restart\": %zu, ", after_interrupt, after_restart);
    }
    md_script_eval_free(gpu); md_script_eval_free(cpu);
    printf("\"parity\": %s}\n", (bad || bad2) ? "false" : "true");
    return (bad || bad2) ? 5 : 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: shim_harness lower|eval ...\n"); return 1; }
    if (strcmp(argv[1], "lower") == 0) return mode_lower(argc, argv);
    if (strcmp(argv[1], "eval") == 0) return mode_eval(argc, argv);
    if (strcmp(argv[1], "dropin") == 0) return mode_dropin(argc, argv);
    return 1;
}
