#!/bin/bash
# parity of the current build on the resampler-heavy tests, then A/B of library builds on cfg2 (VAD off / on) and cfg3:
#   tools/gpu_ab2.sh libA.so libB.so ...     (files under audio-flow-rs_b200/lib/)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_session_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -3
fi
for rep in 1 2; do
for lib in "$@"; do
  AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$lib timeout 200 python bench.py --quick --steps 30 --warmup 5 --e2e-steps 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$lib cfg2', round(d['ms_per_step'],4), 'vad', round(d['with_vad']['ms_per_step'],4))"
  AF_GPU_LIB=$PWD/audio-flow-rs_b200/lib/$lib timeout 200 python bench.py --workload cfg3 --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$lib cfg3', round(d['ms_per_step_without_gather'],3))"
done
done
