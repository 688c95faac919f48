#!/bin/bash
# full ncu capture of the fused kernel, VAD off (launch 4 of the bench) [+ optional VAD on]
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_fused -s 3 -c 1 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
if [ "$1" = "vad" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_fused -s 9 -c 1 -f -o gpurun_out/prof_fused_vad $CMD > gpurun_out/ncu_full_vad.log 2>&1
fi
