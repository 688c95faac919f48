#!/bin/bash
# e2e leg (host buffers, H2D + D2H inside the timed region) for several sub-batch sizes / slot counts
for cfg in "96 3" "32 3" "16 4" "48 3"; do
  set -- $cfg
  AF_HOST_SUB_MB=$1 AF_HOST_SLOTS=$2 timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('sub_mb $1 slots $2 e2e ms', round(d['e2e']['ms_per_step'],2))"
done
