set -x
mkdir -p gpurun_out
timeout 600 python profiles/leak_loop.py 60 > gpurun_out/r2_28_leak_loop.log 2>&1; tail -4 gpurun_out/r2_28_leak_loop.log
timeout 900 compute-sanitizer --leak-check full --error-exitcode 9 python profiles/leak_loop.py 3 > gpurun_out/r2_28_leakcheck.log 2>&1; echo leakcheck rc=$?; grep -E "LEAK SUMMARY|ERROR SUMMARY|Leaked" gpurun_out/r2_28_leakcheck.log | head -5
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-iso > gpurun_out/r2_28_bench_20steps.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_28_bench_20steps.json')); print(round(d['value']), round(d['e2e']['value']), d['clocks'])"
