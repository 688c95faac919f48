#!/bin/bash
# the default bench line under torchrun at N = $1 GPUs (cfg2 weak scaling, VAD + gather, e2e, cfg3 strong scaling)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/gpus_${N}.txt
TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench N=$N exit $?"; tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().split('\n')[-1])
for k in ('value','ms_per_step','per_rank_ms_per_step','sustained','with_vad','e2e','e2e_variants','cfg3'):
    print(k, json.dumps(d.get(k))[:1200])
PY
