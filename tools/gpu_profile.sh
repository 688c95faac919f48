#!/bin/bash
# ncu passes on the bench command (after a plain run of the same command exited 0), outputs in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:af_fused -s 3 -c 1 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/plain.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
