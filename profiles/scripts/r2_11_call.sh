set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_11_ab_$tag.json 2>/dev/null; python - <<PY
import json; d=json.load(open('gpurun_out/r2_11_ab_$tag.json')); print('$tag', round(d['value']), {k: round(v,4) for k,v in d['roofline']['kernels_alone_ms_per_launch'].items()})
PY
}
run t0 MDGPU_PAIR_TUNE=0
run t1 MDGPU_PAIR_TUNE=1
run t2 MDGPU_PAIR_TUNE=2
run t3 MDGPU_PAIR_TUNE=3
run t0b MDGPU_PAIR_TUNE=0
run t2b MDGPU_PAIR_TUNE=2
MDGPU_PAIR_TUNE=3 timeout 600 python -m pytest tests -m gpu -q -k "rdf or parity or variants" 2>&1 | tail -2
