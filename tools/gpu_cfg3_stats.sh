#!/bin/bash
# per-role wait statistics of the fused kernel on cfg3 (stats build) + the plain cfg3 line
mkdir -p gpurun_out
AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/libaudioflow_gpu_stats.so timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 2 --pipe-stats > gpurun_out/cfg3_stats.json 2> gpurun_out/cfg3_stats.err
grep pipe-stats gpurun_out/cfg3_stats.err
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 2 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -1 gpurun_out/bench_cfg3.json | cut -c1-400
