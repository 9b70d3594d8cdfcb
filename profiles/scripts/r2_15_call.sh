set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-iso $EXTRA > gpurun_out/r2_15_ab_$tag.json 2>/dev/null; python - <<PY
import json; d=json.load(open('gpurun_out/r2_15_ab_$tag.json')); print('$tag', round(d['value']))
PY
}
EXTRA=""; run v0 X=1
EXTRA="--rdf-variant 2"; run v2 X=1
EXTRA="--rdf-variant 2"; run v2_cull8 MDGPU_CULL_OCC=8
EXTRA=""; run v0_cull8 MDGPU_CULL_OCC=8
EXTRA="--rdf-variant 2 --streams 6"; run v2_s6 X=1
EXTRA="--streams 6"; run v0_s6 X=1
EXTRA="--streams 2"; run v0_s2 X=1
EXTRA=""; run v0b X=1
