set -x
mkdir -p gpurun_out
timeout 1500 python tests/golden/fuzz_gpu.py 150 707 > gpurun_out/r2_31_fuzz_gpu.log 2>&1; tail -1 gpurun_out/r2_31_fuzz_gpu.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
