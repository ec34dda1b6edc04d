#!/bin/bash
# N GPUs (N = visible devices): DDP parity worker (torchrun) + the N-GPU bench line
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 tests/ddp_worker.py > gpurun_out/ddp$N.log 2>&1; echo "ddp_worker world $N exit $?"; grep "ddp_worker:" gpurun_out/ddp$N.log | sort | uniq; tail -2 gpurun_out/ddp$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench$N exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print(d['e2e']['value']); print(d['clocks'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_${N}gpu.err').read()[-3000:])
PY
