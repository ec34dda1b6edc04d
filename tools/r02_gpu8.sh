#!/bin/bash
timeout 200 python tools/coresidency_probe.py 2>&1 | grep -v Warn | tail -5
timeout 200 python -m pytest tests/test_kernels.py -q --timeout 200 --tb=short -k "adamw" > gpurun_out/pytest_adamw.log 2>&1; tail -12 gpurun_out/pytest_adamw.log | cut -c1-250
for v in "B2_ADAMW_BACKGROUND=1" ; do
  env $v timeout 300 python bench.py --steps 30 --warmup 5 --no-torch-eager --no-cpu-baseline --no-varlen > "gpurun_out/bench8_$(echo $v | tr ' =' '__').json" 2> gpurun_out/bench8.err
  echo "$v rc $?"; python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['pass'], d['parity']['max_dloss'], d['parity']['max_dweight'])" "gpurun_out/bench8_$(echo $v | tr ' =' '__').json"
done
