#!/bin/bash
# diagnosis of one build: per-role wait statistics (stats build) on cfg2 and one ncu --set full capture of the fused kernel (VAD off)
mkdir -p gpurun_out
bash tools/gpu_stats.sh
export AF_BENCH_PRE_MS=0
CMD="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --quick"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:af_fused -s 4 -c 1 -f -o gpurun_out/prof_diag $CMD > gpurun_out/ncu_diag.log 2>&1
echo "ncu exit $?"
