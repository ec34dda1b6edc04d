#!/bin/bash
mkdir -p gpurun_out
for ew in 8 16; do
B2_GEMM_EPI_WARPS=$ew timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ew$ew.json 2> gpurun_out/bench_ew$ew.err; echo "bench ew=$ew exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_ew$ew.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','loss')}); print(d['e2e']['value'], d['roofline']['achieved'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_ew$ew.err').read()[-3000:])
PY
done
B2_GEMM_EPI_WARPS=8 timeout 600 python -m pytest tests/test_gemm.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
