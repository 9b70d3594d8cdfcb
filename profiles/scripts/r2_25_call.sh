set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_25_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_25_gpu_tests.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_new_ops.py -q -x -k "one_position_argument or contexts" > gpurun_out/r2_25_memcheck_ctx.log 2>&1; echo memcheck rc=$?; tail -2 gpurun_out/r2_25_memcheck_ctx.log
timeout 900 python tests/golden/fuzz_gpu.py 30 606 > gpurun_out/r2_25_fuzz_gpu.log 2>&1; tail -1 gpurun_out/r2_25_fuzz_gpu.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_25_bench.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_25_bench.json')); print(round(d['value']), round(d['e2e']['value']))"
