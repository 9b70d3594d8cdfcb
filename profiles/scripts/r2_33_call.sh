set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --frames-per-step 296 --dist-backend gloo --one-device > gpurun_out/r2_33_bench_n8_rehearsal.json 2> gpurun_out/r2_33_bench_n8_rehearsal.err; echo rc=$?; cut -c1-400 gpurun_out/r2_33_bench_n8_rehearsal.json; tail -3 gpurun_out/r2_33_bench_n8_rehearsal.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_33_ref_n8.json 2> gpurun_out/r2_33_ref_n8.err; echo rc=$?; cut -c1-300 gpurun_out/r2_33_ref_n8.json
