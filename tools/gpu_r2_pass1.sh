#!/bin/bash
# round-2 re-entry pass: smoke, the -m gpu suite, the default bench line, then the ncu evidence pass
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 400 -v --durations=15 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
grep -n "FAILED\|ERROR\|passed\|failed\|exit" gpurun_out/pytest.log | tail -20
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench N=1 exit $?"; tail -3 gpurun_out/bench.err
tail -1 gpurun_out/bench.json | cut -c1-3000
bash tools/gpu_profile_final.sh
