#!/bin/bash
mkdir -p gpurun_out
for ew in 16 8; do
B2_GEMM_EPI_WARPS=$ew timeout 600 python -m pytest tests/test_gemm.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gemm_ew$ew.log 2>&1; echo "pytest gemm ew=$ew exit $?"; tail -3 gpurun_out/pytest_gemm_ew$ew.log
done
timeout 600 python -m pytest tests/test_kernels.py tests/test_model.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest quick exit $?"; tail -3 gpurun_out/pytest_quick.log
for ew in 16 8; do
B2_GEMM_EPI_WARPS=$ew timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ew$ew.json 2> gpurun_out/bench_ew$ew.err; echo "bench ew=$ew exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_ew$ew.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])
    for g in d['roofline']['detail']: print('   ', g['gemm'][:44].ljust(44), g['us'], g['tflops'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_ew$ew.err').read()[-3000:])
PY
done
