#!/bin/bash
# Evidence pass for profiles/: (1) plain run, (2) ncu --set full of the fused kernel with the VAD off and on, (3) full
# capture of the parallel scan kernel, (4) the launch list of the same command, (5) the fused kernel on cfg3.
# AF_BENCH_PRE_MS=0 makes the launch order fixed: with --steps 3 --warmup 3 --quick the bench launches the fused kernel
# 7 times with the VAD off (3 warm-up, 1 pre, 3 timed: launches 0-6) and 7 times with the VAD on (7-13, each followed by the scan).
mkdir -p gpurun_out
export AF_BENCH_PRE_MS=0
CMD="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --quick"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_fused -s 4 -c 1 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_fused -s 11 -c 1 -f -o gpurun_out/prof_fused_vad $CMD > gpurun_out/ncu_full_vad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_vad_scan -s 3 -c 1 -f -o gpurun_out/prof_scan $CMD > gpurun_out/ncu_full_scan.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
if [ -n "$WITH_CFG3" ]; then
CMD3="python bench.py --workload cfg3 --steps 2 --warmup 3"
timeout 300 $CMD3 > gpurun_out/plain_cfg3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:af_fused -s 3 -c 1 -f -o gpurun_out/prof_cfg3 $CMD3 > gpurun_out/ncu_full_cfg3.log 2>&1
fi
tail -1 gpurun_out/plain.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
