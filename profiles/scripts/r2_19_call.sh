set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_19_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_19_gpu_tests.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_19_bench.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_19_bench.json')); print(round(d['value']), round(d['e2e']['value']))"
