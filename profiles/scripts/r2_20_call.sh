set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_20_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_20_gpu_tests.log
# soak of the zero-edit drop-in: 16 threads x disjoint ranges on ONE eval, 4000 frames of a 12288-atom box, interrupt + clear + full re-evaluation, compared with the reference's CPU path
oracle/build/synth_tool water-gro 16 2024 /tmp/w16.gro
S="r = rdf(element('O'), element('O'), 8.0); d = distance(1,10); dz = density_z(element('O')); dp = distance_pair(residue(1:4), residue(10:15)); v = sdf(residue(1:200), element('O'), 6.0); rw = rdf(within(4.0, residue(1:20)), element('O'), 6.0); aa = angle(residue(1:2), residue(5:7), 30); rt = rdf(element('O'), residue(10:60), 6.0); dm = distance_min(residue(1:4), residue(100:130)); cc = contact_count(residue(1:5), residue(10:400), 4.0);"
for t in 16 3; do
timeout 1200 oracle/_ref/shim_harness dropin --sys /tmp/w16.gro --traj synthwater:16:2024:4000 --script "$S" --threads $t --interrupt-at 1500 > gpurun_out/r2_20_soak_t$t.json 2> gpurun_out/r2_20_soak_t$t.err; echo rc=$?; tail -1 gpurun_out/r2_20_soak_t$t.json | cut -c1-1200; tail -2 gpurun_out/r2_20_soak_t$t.err
done
K="one_position_argument or as_target or beyond_half"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_new_ops.py -q -x -k "$K" > gpurun_out/r2_20_memcheck_newforms.log 2>&1; echo memcheck rc=$?; tail -3 gpurun_out/r2_20_memcheck_newforms.log
