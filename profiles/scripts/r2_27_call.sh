set -x
mkdir -p gpurun_out
nvidia-smi -L | head -2; lscpu | grep -E "Model name|^CPU\(s\)|NUMA" 
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_27_gpu_tests.log 2>&1; tail -4 gpurun_out/r2_27_gpu_tests.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2_27_bench.json 2> gpurun_out/r2_27_bench.err; cat gpurun_out/r2_27_bench.json; tail -3 gpurun_out/r2_27_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_27_bench_reference_arm.json 2> gpurun_out/r2_27_bench_reference_arm.err; cat gpurun_out/r2_27_bench_reference_arm.json | cut -c1-600
for c in 2 3 4; do timeout 600 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_27_bench_config$c.json 2> gpurun_out/r2_27_bench_config$c.err; done
timeout 600 python profiles/newops_workload.py > gpurun_out/r2_27_newops_timings.jsonl 2>&1; tail -3 gpurun_out/r2_27_newops_timings.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rdf_cull_full|k_sdf_scatter|k_rdf_pairs_v2|k_bin_points|k_scatter_points|k_sdf_fit" -c 12 -o gpurun_out/r2_27_kernels python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-iso > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_density|k_rdf_pairs_v2|k_rdf_cull_full" -c 4 -o gpurun_out/r2_27_config4_kernels python bench.py --config 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-iso > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_27_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-iso > /dev/null 2>&1
ls -la gpurun_out | tail -14
