set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_02_gpu_tests.log 2>&1; tail -6 gpurun_out/r2_02_gpu_tests.log
python bench.py --steps 6 --warmup 3 > gpurun_out/r2_02_bench.json 2> gpurun_out/r2_02_bench.err; cat gpurun_out/r2_02_bench.json; tail -3 gpurun_out/r2_02_bench.err
for v in 2 4; do python bench.py --steps 6 --warmup 3 --rdf-variant $v --no-cpu-baseline --no-e2e > gpurun_out/r2_02_bench_variant$v.json 2>/dev/null; cat gpurun_out/r2_02_bench_variant$v.json; done
python bench.py --steps 4 --warmup 3 --ingest-mode 1 --no-cpu-baseline --no-iso > gpurun_out/r2_02_bench_wholeframe_ingest.json 2>/dev/null; cat gpurun_out/r2_02_bench_wholeframe_ingest.json
for t in 4 8 32; do python bench.py --steps 4 --warmup 3 --ingest-threads $t --no-cpu-baseline --no-iso > gpurun_out/r2_02_bench_ingest_t$t.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/r2_02_bench_ingest_t$t.json'));print($t,d['value'],d['e2e']['value'])"; done
for c in 2 3 4; do python bench.py --config $c --steps 4 --warmup 3 > gpurun_out/r2_02_bench_config$c.json 2> gpurun_out/r2_02_bench_config$c.err; cat gpurun_out/r2_02_bench_config$c.json; tail -2 gpurun_out/r2_02_bench_config$c.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-iso > /dev/null 2>&1
for v in 0 4; do ncu --set full --clock-control none --import-source on -k regex:"k_rdf_pairs_v2" -c 1 -o gpurun_out/r2_02_pairs_variant$v python bench.py --steps 1 --warmup 1 --rdf-variant $v --no-e2e --no-cpu-baseline --no-iso > /dev/null 2>&1; done
ls -la gpurun_out
