#!/bin/bash
# round-2 GPU call 4 (2 GPUs): the 2-rank DDP tests, bench A at N=2 (parity block, varlen leg), kernel-form exchange A/B
timeout 1200 python -m pytest tests/test_ddp.py -q --timeout 900 --tb=short --durations=5 > gpurun_out/pytest_ddp.log 2>&1
echo "ddp pytest rc $?"; tail -40 gpurun_out/pytest_ddp.log
for v in "B2_DDP_DMA=1" "B2_DDP_DMA=0"; do
  env $v timeout 600 python bench.py --gpus 2 --steps 30 --warmup 5 --no-torch-eager > "gpurun_out/bench4_$(echo $v | tr ' =' '__').json" 2> gpurun_out/bench4.err
  echo "$v rc $?"; tail -3 gpurun_out/bench4.err; python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'], d.get('varlen'))" "gpurun_out/bench4_$(echo $v | tr ' =' '__').json"
done
