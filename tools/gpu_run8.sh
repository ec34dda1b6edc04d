#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model.py -q -m gpu -p no:cacheprovider > gpurun_out/test_model.log 2>&1; echo "test_model exit $?"; tail -3 gpurun_out/test_model.log
for ws in 1 0; do
B2_WGRAD_STREAM=$ws timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ws$ws.json 2> gpurun_out/bench.err; echo "bench wgrad_stream=$ws exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_ws$ws.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
done
