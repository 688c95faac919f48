#!/bin/bash
# cfg3 and the 44.1 kHz-only batch, alternating repetitions of each library build: tools/gpu_ab_cfg3.sh libA.so libB.so ...
for rep in 1 2; do
for lib in "$@"; do
  AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$lib timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$lib cfg3', round(d['ms_per_step_without_gather'],3))"
  AF_CFG3_STREAMS=512 AF_CFG3_RATE=44100 AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$lib timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$lib 44.1k x512', round(d['ms_per_step_without_gather'],3))"
done
done
