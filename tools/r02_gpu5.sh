#!/bin/bash
# round-2 GPU call 5 (1 GPU): how much of AdamW / the weight-gradient stream is exposed; varlen leg at N=1
for v in "B2_DEBUG_SKIP_ADAMW=0" "B2_DEBUG_SKIP_ADAMW=1" "B2_DEBUG_SKIP_ADAMW=1 B2_WGRAD_STREAM=0" "B2_WGRAD_STREAM=0"; do
  env $v timeout 300 python bench.py --steps 30 --warmup 5 --no-torch-eager --no-cpu-baseline --no-parity --no-varlen > "gpurun_out/bench5_$(echo $v | tr ' =' '__').json" 2> gpurun_out/bench5.err
  echo "$v rc $?"; python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(d['value'], d['ms_per_step'], d['e2e']['value'])" "gpurun_out/bench5_$(echo $v | tr ' =' '__').json"
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-torch-eager --no-cpu-baseline --no-parity > gpurun_out/bench5_varlen.json 2> gpurun_out/bench5.err
python -c "
import json
d=json.load(open('gpurun_out/bench5_varlen.json')); print(d['value'], d['ms_per_step'], d['varlen'])"
