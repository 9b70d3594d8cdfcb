set -x
mkdir -p gpurun_out
timeout 300 python profiles/config5_million_frames.py --frames 40000 > gpurun_out/r2_32_config5_40k.json 2> gpurun_out/r2_32.err; tail -1 gpurun_out/r2_32_config5_40k.json | cut -c1-600; tail -3 gpurun_out/r2_32.err
timeout 900 python profiles/config5_million_frames.py > gpurun_out/r2_32_config5_1M.json 2>> gpurun_out/r2_32.err; tail -1 gpurun_out/r2_32_config5_1M.json | cut -c1-700
