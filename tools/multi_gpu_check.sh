#!/bin/bash
# 8 GPUs: the driver's scaling line (bench at N=8) + a DDP parity pass at world 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_8gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print(d['e2e']); print(d['clocks'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_8gpu.err').read()[-3000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 tests/ddp_worker.py > gpurun_out/ddp8.log 2>&1; echo "ddp_worker world 8 exit $?"; grep "ddp_worker" gpurun_out/ddp8.log | tail -4; tail -3 gpurun_out/ddp8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "bench4 exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_4gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print(d['e2e']['value'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_4gpu.err').read()[-3000:])
PY
