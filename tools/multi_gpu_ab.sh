#!/bin/bash
# A/B of the two exchange forms at N = visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 tests/ddp_worker.py > gpurun_out/ddp$N.log 2>&1; echo "ddp_worker world $N (dma) exit $?"; grep "ddp_worker:" gpurun_out/ddp$N.log | sort | uniq; tail -2 gpurun_out/ddp$N.log
for dma in 1 0; do
B2_DDP_DMA=$dma timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$dma bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${N}gpu_dma$dma.json 2> gpurun_out/bench_${N}gpu_dma$dma.err; echo "bench$N dma=$dma exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${N}gpu_dma$dma.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['e2e']['value'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_${N}gpu_dma$dma.err').read()[-3000:])
PY
done
