#!/bin/bash
mkdir -p gpurun_out
for f in test_gemm test_kernels test_model; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; tail -4 gpurun_out/$f.log
done
timeout 600 python tools/diag_full.py 12 2>/dev/null | awk '{print $1,$3,$5}' | sort -k2 -n -r | head -6
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e']); r=d['roofline']; print(r['achieved'],r['frac'],r['step_achieved_tflops_per_gpu'])
    for x in r['detail']: print(x)
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 800 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 4 > gpurun_out/profile_step.log 2>&1; echo "ncu list exit $?"
