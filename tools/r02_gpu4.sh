#!/bin/bash
# round-2 GPU call (2 GPUs): the 2-rank DDP tests, bench A at N=2 (parity block, varlen leg, exchange block)
timeout 1200 python -m pytest tests/test_ddp.py -q --timeout 900 --tb=short --durations=5 > gpurun_out/pytest_ddp.log 2>&1
echo "ddp pytest rc $?"; tail -15 gpurun_out/pytest_ddp.log | cut -c1-300
timeout 600 python bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench rc $?"; python -c "
import json
txt=open('gpurun_out/bench_2gpu.json').read(); d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['pass'], d['exchange'], d['varlen'], d['torch_eager'])"
