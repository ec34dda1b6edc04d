#!/bin/bash
mkdir -p gpurun_out
for f in test_gemm test_kernels test_model; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; tail -3 gpurun_out/$f.log
done
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e']); r=d['roofline']; print(r['achieved'],r['frac'],r['step_achieved_tflops_per_gpu'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
