set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_29_ab_$tag.json 2>/dev/null; python - <<PY
import json; d=json.load(open('gpurun_out/r2_29_ab_$tag.json')); print('$tag', round(d['value']), {k: round(v,4) for k,v in d['roofline']['kernels_alone_ms_per_launch'].items()}, d['checks'])
PY
}
run default X=1
run flat6 MDGPU_CULL=flat
run flat8 MDGPU_CULL=flat MDGPU_CULL_OCC=8
run flat4 MDGPU_CULL=flat MDGPU_CULL_OCC=4
run default2 X=1
MDGPU_CULL=flat timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
