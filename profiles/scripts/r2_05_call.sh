set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2_05_bench_2gpu.json 2> gpurun_out/r2_05_bench_2gpu.err; cat gpurun_out/r2_05_bench_2gpu.json; tail -3 gpurun_out/r2_05_bench_2gpu.err
timeout 600 python profiles/multi_device_e2e.py --devices 2 > gpurun_out/r2_05_multi_device_e2e.json 2> gpurun_out/r2_05_multi_device_e2e.err; cat gpurun_out/r2_05_multi_device_e2e.json; tail -5 gpurun_out/r2_05_multi_device_e2e.err
MDGPU_DEVICES=0,1 timeout 600 python -m pytest tests/test_integration_shim.py -m gpu -q -x > gpurun_out/r2_05_dropin_2gpu.log 2>&1; tail -3 gpurun_out/r2_05_dropin_2gpu.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_05_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_05_gpu_tests.log
