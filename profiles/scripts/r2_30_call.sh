set -x
mkdir -p gpurun_out
timeout 600 python profiles/e2e_phases.py > gpurun_out/r2_30_e2e_phases.log 2>&1; tail -8 gpurun_out/r2_30_e2e_phases.log
