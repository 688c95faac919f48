#!/bin/bash
# the non-headline workloads (cfg3 / cfg4 / cfg5) on one GPU + the default bench line with e2e and CPU baseline
mkdir -p gpurun_out
for w in cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "$w exit $?"; tail -1 gpurun_out/bench_$w.json | cut -c1-700
done
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -1 gpurun_out/bench.json | cut -c1-2500
