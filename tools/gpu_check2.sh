#!/bin/bash
# Quick parity subset + bench + ncu of the fused kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 600 -k "not full_size and not mixed_rate_device" > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
for v in tma sync; do
  timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline --variant $v > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  echo "bench $v exit $?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$v.json').read().strip().split('\n')[-1])
print('$v', d['value'], d['ms_per_step'], d['roofline']['frac'], d['with_vad']['ms_per_step'], d['clocks'])"
done
CMD="python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:af_fused -s 3 -c 1 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
grep -E "af_" gpurun_out/launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -12
