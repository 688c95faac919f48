#!/bin/bash
# A/B of library builds on cfg2 (VAD off / on), cfg3 and a 44.1 kHz-only batch of 512 streams:
#   tools/gpu_ab3.sh libA.so libB.so ...     (files under audio-flow-rs_b200/lib/)
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "$@"; do
  AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$lib timeout 200 python bench.py --quick --steps 30 --warmup 5 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$lib cfg2', round(d['ms_per_step'],4), 'vad', round(d['with_vad']['ms_per_step'],4))"
  AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$lib timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$lib cfg3', round(d['ms_per_step_without_gather'],3))"
  AF_CFG3_STREAMS=512 AF_CFG3_RATE=44100 AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$lib timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$lib 44.1k x512', round(d['ms_per_step_without_gather'],3))"
done
done
