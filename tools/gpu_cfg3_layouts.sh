#!/bin/bash
# cfg3 (4096 x 30 s mixed 44.1/48 kHz, VAD on, one GPU) under each warp layout
for l in 0 2 3; do
  AF_LAYOUT=$l timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('cfg3 layout $l', round(d['ms_per_step'],3), round(d['hbm_frac_per_gpu'],4))"
done
