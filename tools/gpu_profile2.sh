#!/bin/bash
# full ncu capture of the fused kernel with VAD off (launch 4) and VAD on (launch 10) + scan kernel
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:af_fused -s 3 -c 1 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:af_fused -s 9 -c 1 -f -o gpurun_out/prof_fused_vad $CMD > gpurun_out/ncu_full_vad.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:af_vad_scan -s 3 -c 1 -f -o gpurun_out/prof_scan $CMD > gpurun_out/ncu_full_scan.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
