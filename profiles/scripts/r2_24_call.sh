set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_24_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_24_gpu_tests.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_new_ops.py -q -x -k "one_position_argument or as_target" > gpurun_out/r2_24_memcheck_arrays.log 2>&1; echo memcheck rc=$?; tail -2 gpurun_out/r2_24_memcheck_arrays.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2_24_bench.json 2> gpurun_out/r2_24_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_24_bench.json')); print(round(d['value']), round(d['e2e']['value']), d['cpu_baseline']['value'], d['roofline']['kernels_alone_ms_per_launch'])"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
