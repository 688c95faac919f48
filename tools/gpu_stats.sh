#!/bin/bash
# pipe statistics of the fused kernel (stats build) for cfg2, VAD off and on
mkdir -p gpurun_out
AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/libaudioflow_gpu_stats.so timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --no-cpu-baseline --pipe-stats > gpurun_out/pipe_stats.json 2> gpurun_out/pipe_stats.err
cat gpurun_out/pipe_stats.err
python -c "
import json
d=json.loads(open('gpurun_out/pipe_stats.json').read().strip().split('\n')[-1]); print(d['ms_per_step'], d['with_vad']['ms_per_step'])"
