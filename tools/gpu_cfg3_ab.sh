#!/bin/bash
# cfg3 (4096 x 30 s mixed 44.1/48 kHz, VAD on, one GPU) for several builds of the library: tools/gpu_cfg3_ab.sh libA.so libB.so
for lib in "$@"; do
  AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$lib timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('cfg3 $lib', round(d['ms_per_step'],3), round(d['hbm_frac_per_gpu'],4))"
done
