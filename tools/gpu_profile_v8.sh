#!/bin/bash
# pipe statistics (stats build), then the evidence pass of tools/gpu_profile_final.sh
mkdir -p gpurun_out
AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/libaudioflow_gpu_stats.so timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline --pipe-stats > gpurun_out/pipe_stats.json 2> gpurun_out/pipe_stats.err
cat gpurun_out/pipe_stats.err
bash tools/gpu_profile_final.sh
