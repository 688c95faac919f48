#!/bin/bash
# Evidence pass for profiles/: (1) plain run, (2) ncu --set full of the fused kernel with the VAD off (launch 4 of the
# bench) and on (launch 10), (3) full capture of the parallel scan kernel, (4) the launch list of the same command.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_fused -s 3 -c 1 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_fused -s 9 -c 1 -f -o gpurun_out/prof_fused_vad $CMD > gpurun_out/ncu_full_vad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_vad_scan -s 3 -c 1 -f -o gpurun_out/prof_scan $CMD > gpurun_out/ncu_full_scan.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
