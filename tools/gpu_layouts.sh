#!/bin/bash
# cfg2 kernel time under each warp layout (AF_LAYOUT override), VAD off / on
for l in 0 1 2 3; do
  AF_LAYOUT=$l timeout 300 python bench.py --steps 30 --warmup 5 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('layout $l', round(d['ms_per_step'],4), round(d['with_vad']['ms_per_step'],4))"
done
