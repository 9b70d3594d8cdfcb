set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_03_gpu_tests.log 2>&1; tail -4 gpurun_out/r2_03_gpu_tests.log
python bench.py --steps 6 --warmup 3 > gpurun_out/r2_03_bench.json 2> gpurun_out/r2_03_bench.err; cat gpurun_out/r2_03_bench.json; tail -3 gpurun_out/r2_03_bench.err
python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_03_bench_config3.json 2>/dev/null; cat gpurun_out/r2_03_bench_config3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_03_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-iso > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_rdf_cull|k_sdf_scatter|k_rdf_pairs_v2" -c 3 -o gpurun_out/r2_03_kernels python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-iso > /dev/null 2>&1
ls -la gpurun_out | tail -8
