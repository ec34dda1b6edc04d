#!/bin/bash
# development loop: kernel/model parity (no full-size oracle run) + one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "not full_size" > gpurun_out/pytest_quick.log 2>&1; echo "pytest quick exit $?"; tail -4 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_quick.err').read()[-3000:])
PY
