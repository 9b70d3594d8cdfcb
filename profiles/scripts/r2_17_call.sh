set -x
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-iso "$@" > gpurun_out/r2_17_ab_$tag.json 2>/dev/null; python - <<PY
import json; d=json.load(open('gpurun_out/r2_17_ab_$tag.json')); print('$tag', round(d['value']), round(d['e2e']['value']))
PY
}
run b148_s4
run b148_s6 --streams 6
run b148_s8 --streams 8
run b74_s4 --batch-frames 74
run b74_s8 --batch-frames 74 --streams 8
run b296_s4 --batch-frames 296
run b148_s4_t32 --ingest-threads 32
run b148_s6_t24 --streams 6 --ingest-threads 24
