set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_07_ab_$tag.json 2>/dev/null; python - <<PY
import json; d=json.load(open('gpurun_out/r2_07_ab_$tag.json')); print('$tag', round(d['value']), {k: round(v,4) for k,v in d['roofline']['kernels_alone_ms_per_launch'].items()})
PY
}
run default X=1
run sdf_occ4 MDGPU_SDF_OCC=4
run sdf_occ5 MDGPU_SDF_OCC=5
run cull_occ6 MDGPU_CULL_OCC=6
run cull_occ8 MDGPU_CULL_OCC=8
run both MDGPU_SDF_OCC=4 MDGPU_CULL_OCC=6
run default2 X=1
