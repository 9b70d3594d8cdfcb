set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_26_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_26_gpu_tests.log
