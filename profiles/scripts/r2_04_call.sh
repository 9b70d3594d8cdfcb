set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e $EXTRA > gpurun_out/r2_04_ab_$tag.json 2>/dev/null; python - <<PY
import json; d=json.load(open('gpurun_out/r2_04_ab_$tag.json')); print('$tag', round(d['value']), {k: round(v,4) for k,v in d['roofline']['kernels_alone_ms_per_launch'].items()})
PY
}
EXTRA="" run default X=1
EXTRA="" run default_again X=1
EXTRA="" run cull_half MDGPU_CULL=half
EXTRA="" run sdf_ring MDGPU_SDF=ring
EXTRA="--rdf-variant 2" run pairs_3cta X=1
EXTRA="--rdf-variant 2" run pairs_3cta_sdf_ring MDGPU_SDF=ring
EXTRA="--streams 4" run streams4 X=1
EXTRA="--streams 2" run streams2 X=1
EXTRA="--batch-frames 296" run batch296 X=1
EXTRA="--batch-frames 74" run batch74 X=1
