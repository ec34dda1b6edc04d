#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_quick.log | cut -c1-250
for v in "B2_ATTN_KEEP_BITS=0" "B2_ATTN_KEEP_BITS=1" "B2_ATTN_KEEP_BITS=0" "B2_ATTN_KEEP_BITS=1"; do
env $v timeout 300 python bench.py --steps 60 --warmup 5 --no-torch-eager --no-cpu-baseline --no-varlen --no-parity > gpurun_out/bench_quick_$v.json 2> gpurun_out/bench_quick.err
echo "$v bench rc $?"; python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(d['value'], d['ms_per_step'], d['e2e']['value'])" gpurun_out/bench_quick_$v.json
done
