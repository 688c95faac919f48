#!/bin/bash
# round-2 first pass: smoke, the whole -m gpu suite (incl. the full-size cfg4 / cfg5 tests), baseline lines of every config
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 400 -v --durations=15 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -40 gpurun_out/pytest.log
for w in cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "$w exit $?"; tail -1 gpurun_out/bench_$w.json | cut -c1-600
done
timeout 300 python bench.py --steps 50 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit $?"; tail -1 gpurun_out/bench_quick.json | cut -c1-1500
