#!/bin/bash
# cfg2 only (VAD off / on), three alternating repetitions of each library build: tools/gpu_ab_cfg2.sh libA.so libB.so ...
for rep in 1 2 3; do
for lib in "$@"; do
  AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$lib timeout 200 python bench.py --quick --steps 50 --warmup 5 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$lib cfg2', round(d['ms_per_step'],4), 'vad', round(d['with_vad']['ms_per_step'],4))"
done
done
