#!/bin/bash
# 44.1 kHz-only batch (512 x 30 s, VAD on): per-role wait statistics (stats build) and one ncu --set full capture of the fused kernel
mkdir -p gpurun_out
export AF_CFG3_STREAMS=512 AF_CFG3_RATE=44100
AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/libaudioflow_gpu_stats.so timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 2 --pipe-stats > gpurun_out/s441_stats.json 2> gpurun_out/s441_stats.err
grep pipe-stats gpurun_out/s441_stats.err
unset AF_CFG3_STREAMS AF_CFG3_RATE
bash tools/gpu_ncu_441.sh
