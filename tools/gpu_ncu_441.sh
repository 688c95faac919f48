#!/bin/bash
# ncu --set full of the fused kernel on a 44.1 kHz-only batch (512 x 30 s, VAD on): where do the resampler warps wait?
mkdir -p gpurun_out
LIB=${1:-libaudioflow_gpu.so}
export AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$LIB AF_CFG3_STREAMS=512 AF_CFG3_RATE=44100
timeout 200 python bench.py --workload cfg3 --steps 3 --warmup 1 > gpurun_out/plain_441.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_fused -s 3 -c 1 -f -o gpurun_out/prof_441 python bench.py --workload cfg3 --steps 3 --warmup 1 > gpurun_out/ncu_441.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/plain_441.log | cut -c1-300
