#!/bin/bash
# gpu suite + the default bench line at N = 1 and under torchrun at N = $1 (default 2)
N=${1:-2}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 600 python -m pytest tests -m gpu -q --tb=short --timeout 200 > gpurun_out/pytest2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest2.log
tail -5 gpurun_out/pytest2.log
fi
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench N=1 exit $?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
def p(k,v): print(k, json.dumps(v)[:900])
for k in ('value','ms_per_step','roofline','sustained','with_vad','e2e','e2e_variants','cpu_baseline','parity','cfg3','cfg4','cfg5','clocks'):
    p(k,d.get(k))
PY
if [ "$N" -gt 1 ]; then
TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench N=$N exit $?"; tail -5 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().split('\n')[-1])
for k in ('value','ms_per_step','per_rank_ms_per_step','with_vad','e2e','e2e_variants','cfg3'):
    print(k, json.dumps(d.get(k))[:900])
PY
fi
