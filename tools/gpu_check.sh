#!/bin/bash
# One GPU-box pass: parity tests, smoke, short bench.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 600 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/pytest.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json | tail -2; tail -3 gpurun_out/bench.err
