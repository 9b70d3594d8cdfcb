set -x
mkdir -p gpurun_out
# compute-sanitizer over the golden-fixture GPU tests (small systems): out-of-bounds / misaligned accesses, then shared-memory hazards
K="golden_water or golden_triclinic or golden_config1 or empty_and_error or variants or overflow_pass or xtc_device_decode"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > gpurun_out/r2_14_memcheck_parity.log 2>&1; echo memcheck_parity rc=$?; tail -6 gpurun_out/r2_14_memcheck_parity.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_new_ops.py -q -x -k "not shim" > gpurun_out/r2_14_memcheck_newops.log 2>&1; echo memcheck_newops rc=$?; tail -6 gpurun_out/r2_14_memcheck_newops.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden_water_rdf_per_frame or golden_water_sdf or golden_triclinic or golden_water_density" > gpurun_out/r2_14_racecheck.log 2>&1; echo racecheck rc=$?; tail -8 gpurun_out/r2_14_racecheck.log
grep -c "ERROR SUMMARY" gpurun_out/r2_14_*.log; grep -h "ERROR SUMMARY" gpurun_out/r2_14_*.log
