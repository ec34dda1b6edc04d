#!/bin/bash
# round-2 GPU call 7 (1 GPU): background AdamW -- kernel test, then step time with / without it
timeout 300 python -m pytest tests/test_kernels.py -q --timeout 200 --tb=short -k "adamw" 2>&1 | tail -5
for v in "B2_ADAMW_BACKGROUND=1" "B2_ADAMW_BACKGROUND=0" "B2_ADAMW_BACKGROUND=1 B2_FUSED_LN=0"; do
  env $v timeout 300 python bench.py --steps 30 --warmup 5 --no-torch-eager --no-cpu-baseline --no-varlen > "gpurun_out/bench7_$(echo $v | tr ' =' '__').json" 2> gpurun_out/bench7.err
  echo "$v rc $?"; tail -2 gpurun_out/bench7.err | cut -c1-300; python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['pass'], d['parity']['max_dloss'], d['parity']['max_dweight'])" "gpurun_out/bench7_$(echo $v | tr ' =' '__').json"
done
