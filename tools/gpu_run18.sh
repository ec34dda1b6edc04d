#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm.py tests/test_model.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest quick exit $?"; tail -3 gpurun_out/pytest_quick.log
for gw in 1 0; do
B2_GROUPED_WGRAD=$gw timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_gw$gw.json 2> gpurun_out/bench_gw$gw.err; echo "bench grouped=$gw exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_gw$gw.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_gw$gw.err').read()[-3000:])
PY
done
