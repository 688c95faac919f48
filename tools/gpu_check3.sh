#!/bin/bash
# smoke (tight timeout: a pipeline deadlock must not hang the box) -> parity tests -> bench
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
grep -q "smoke exit 0" gpurun_out/smoke.log || exit 1
timeout 240 python -m pytest tests -m gpu -q --tb=short --timeout 60 -x -v > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit $?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['with_vad'], d['clocks'])"
