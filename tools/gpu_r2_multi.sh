#!/bin/bash
# round-2: full gpu suite on a 2-GPU box (the multi-GPU tests need it) + the C host
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu2.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 400 -v > gpurun_out/pytest2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest2.log
grep -n "FAILED\|PASSED\|SKIPPED\|ERROR\|passed\|failed" gpurun_out/pytest2.log | tail -90
timeout 120 audio-flow-rs_b200/lib/multi_gpu_host 2 64 30 | tee gpurun_out/multi_gpu_host.log
