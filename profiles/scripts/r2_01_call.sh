set -x
mkdir -p gpurun_out
nvidia-smi -L | head -2; lscpu | grep -E "Model name|^CPU\(s\)|NUMA" 
python -m pytest tests -m gpu -q -s > gpurun_out/r2_01_gpu_tests.log 2>&1; tail -5 gpurun_out/r2_01_gpu_tests.log
timeout 900 python tests/golden/fuzz_gpu.py 40 101 > gpurun_out/r2_01_fuzz_gpu.log 2>&1; tail -3 gpurun_out/r2_01_fuzz_gpu.log
python bench.py --steps 6 --warmup 3 > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err; cat gpurun_out/r2_01_bench.json
python profiles/newops_workload.py > gpurun_out/r2_01_newops_timings.jsonl 2>&1; cat gpurun_out/r2_01_newops_timings.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_01_newops_launches.csv python profiles/newops_workload.py --reps 1 --warm 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_01_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_rdf_cull|k_bin_points|k_scatter_points|k_scan_cells|k_sdf_fit|k_sdf_scatter|k_rdf_pairs_v2" -c 12 -o gpurun_out/r2_01_fullset python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
