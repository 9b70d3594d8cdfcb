set -x
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-iso > gpurun_out/r2_18_ramp_$i.json 2>/dev/null; python - <<PY
import json; d=json.load(open('gpurun_out/r2_18_ramp_$i.json')); print('ramp$i', round(d['value']), round(d['e2e']['value']), d['gpu_launches'])
PY
done
timeout 300 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline --no-iso > gpurun_out/r2_18_ramp_c3.json 2>/dev/null; python - <<PY
import json; d=json.load(open('gpurun_out/r2_18_ramp_c3.json')); print('c3', round(d['value']), round(d['e2e']['value']))
PY
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
