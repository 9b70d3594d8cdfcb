set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_16_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_16_gpu_tests.log
timeout 900 python tests/golden/fuzz_gpu.py 30 505 > gpurun_out/r2_16_fuzz_gpu.log 2>&1; tail -1 gpurun_out/r2_16_fuzz_gpu.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2_16_bench.json 2> gpurun_out/r2_16_bench.err; cat gpurun_out/r2_16_bench.json | cut -c1-300
