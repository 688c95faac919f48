#!/bin/bash
# A/B timing of two builds of the library on cfg2 (kernel-only): tools/gpu_ab.sh libA.so libB.so
mkdir -p gpurun_out
for lib in "$@"; do
  for rep in 1 2; do
    AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$lib timeout 300 python bench.py --steps 30 --warmup 5 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$lib', round(d['ms_per_step'],4), round(d['with_vad']['ms_per_step'],4))"
  done
done
