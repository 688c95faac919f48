#!/bin/bash
# pass 2: the C host test again, per-role wait statistics on a 44.1 kHz-only batch and on cfg2, ncu of the 44.1 kHz batch
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short --timeout 200 2>&1 | tail -5
export AF_CFG3_STREAMS=512 AF_CFG3_RATE=44100
AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/libaudioflow_gpu_stats.so timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 2 --pipe-stats > gpurun_out/s441_stats.json 2> gpurun_out/s441_stats.err
grep pipe-stats gpurun_out/s441_stats.err
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 2 > gpurun_out/bench_441.json 2> gpurun_out/bench_441.err; tail -1 gpurun_out/bench_441.json | cut -c1-700
unset AF_CFG3_STREAMS AF_CFG3_RATE
bash tools/gpu_stats.sh
bash tools/gpu_ncu_441.sh
