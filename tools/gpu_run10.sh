#!/bin/bash
# A/B of the in-place split-K dgrads + 2-CTA/SM attention backward; parity first
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm.py tests/test_kernels.py tests/test_model.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest quick exit $?"; tail -4 gpurun_out/pytest_quick.log
for acc in 1 0; do
B2_ACCUM_DGRAD=$acc timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_acc$acc.json 2> gpurun_out/bench_acc$acc.err; echo "bench acc=$acc exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_acc$acc.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e']); r=d['roofline']; print(r['achieved'],r['frac'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_acc$acc.err').read()[-3000:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 --csv --log-file gpurun_out/launches_r01e.csv python tools/profile_step.py 2 > gpurun_out/prof_e.log 2>&1; echo "ncu list exit $?"
