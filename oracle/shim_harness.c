/* shim_harness.c — TEST INFRASTRUCTURE: proves the drop-in boundary end to end.
 *
 * This translation unit is the reference's md_script.c (included verbatim from /root/reference, exactly as its own white-box tests do,
 * mdlib/unittest/test_script.c:20) followed by integration/md_script_mdgpu.inl — i.e. what a maintainer's build would contain.
 *   shim_harness lower --sys F --script S --out O
 *        compile S with the UNMODIFIED md_script front-end, lower the IR with the shim, dump the descriptors (no GPU needed)
 *   shim_harness eval  --sys F --traj SPEC --script S [--frames B:E]
 *        evaluate S twice through the md_script API: md_script_eval_frame_range (reference CPU path) and
 *        md_script_gpu_eval_frame_range (libmdgpu), then compare the md_script_property_data_t contents. Exit code 0 = parity.
 */
#include <md_script.c>
#include <md_gro.h>
#include <md_pdb.h>
#include "harness_common.h"
#include "../integration/md_script_mdgpu.inl"

static int mode_lower(int argc, char** argv) {
    md_allocator_i* alloc = md_vm_arena_create(GIGABYTES(8));
    md_system_t sys; if (!load_system(&sys, arg_val(argc, argv, "--sys", ""), alloc)) return 2;
    const char* src = arg_val(argc, argv, "--script", "");
    md_script_ir_t* ir = md_script_ir_create(alloc);
    if (!md_script_ir_compile_from_source(ir, (str_t){ src, strlen(src) }, &sys, NULL, NULL) || !md_script_ir_valid(ir)) { fprintf(stderr, "script failed to compile\n"); return 2; }
    md_script_gpu_lowered_t low = {0};
    if (!md_script_gpu_lower(&low, ir, alloc)) return 3;
    FILE* f = fopen(arg_val(argc, argv, "--out", "lowered.bin"), "wb"); if (!f) return 2;
    uint64_t np = low.num_props; fwrite("MDLOWER1", 1, 8, f); fwrite(&np, 8, 1, f);
    for (size_t i = 0; i < low.num_props; ++i) {
        const mdgpu_property_desc_t* p = &low.props[i];
        uint64_t v[3] = { p->op, p->num_structures, p->structure_size };
        fwrite(low.names[i], 1, 64, f); fwrite(v, 8, 3, f); fwrite(&p->cutoff_min, 4, 1, f); fwrite(&p->cutoff_max, 4, 1, f);
        for (int k = 0; k < 4; ++k) { uint64_t c = p->idx_count[k]; fwrite(&c, 8, 1, f); if (c) fwrite(p->idx[k], 4, c, f); }
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
    if (!md_script_eval_frame_range(cpu, ir, &sys, &traj, (uint32_t)b, (uint32_t)e)) { fprintf(stderr, "reference evaluation failed\n"); return 2; }
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
            const double tol = 1e-5 * fabs((double)a->values[i]) + 1e-6;   /* north_star: 1e-5 relative for averaged floats */
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

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: shim_harness lower|eval ...\n"); return 1; }
    if (strcmp(argv[1], "lower") == 0) return mode_lower(argc, argv);
    if (strcmp(argv[1], "eval") == 0) return mode_eval(argc, argv);
    return 1;
}
