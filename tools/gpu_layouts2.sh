#!/bin/bash
# warp-to-role layouts (AF_LAYOUT) on a 44.1 kHz-only batch and on cfg3, for one library build
LIB=${1:-libaudioflow_gpu.so}
for l in 0 1 2 3; do
  AF_LAYOUT=$l AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$LIB AF_CFG3_STREAMS=512 AF_CFG3_RATE=44100 timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('layout $l 44.1k x512', round(d['ms_per_step_without_gather'],3))"
  AF_LAYOUT=$l AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$LIB timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('layout $l cfg3', round(d['ms_per_step_without_gather'],3))"
done
