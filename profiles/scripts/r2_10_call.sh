set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_10_gpu_tests.log 2>&1; tail -6 gpurun_out/r2_10_gpu_tests.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_10_bench.json 2> gpurun_out/r2_10_bench.err; cat gpurun_out/r2_10_bench.json; tail -3 gpurun_out/r2_10_bench.err
timeout 900 python tests/golden/fuzz_gpu.py 40 404 > gpurun_out/r2_10_fuzz_gpu.log 2>&1; tail -3 gpurun_out/r2_10_fuzz_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_10_smoke.log 2>&1; tail -2 gpurun_out/r2_10_smoke.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rdf_pairs_v2" -c 2 -o gpurun_out/r2_10_pairs python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-iso > /dev/null 2>&1
ls -la gpurun_out | tail -8
