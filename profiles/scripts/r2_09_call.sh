set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_09_gpu_tests.log 2>&1; tail -6 gpurun_out/r2_09_gpu_tests.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_09_bench.json 2> gpurun_out/r2_09_bench.err; cat gpurun_out/r2_09_bench.json; tail -3 gpurun_out/r2_09_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rdf_pairs_v2" -c 2 -o gpurun_out/r2_09_pairs python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-iso > /dev/null 2>&1
ls -la gpurun_out | tail -6
