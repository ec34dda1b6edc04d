#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels.py tests/test_model.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest quick exit $?"; tail -3 gpurun_out/pytest_quick.log
for v in 1 0; do
B2_ATTN_FWD128=$v timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_f128_$v.json 2> gpurun_out/bench_f128_$v.err; echo "bench fwd128=$v exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_f128_$v.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','loss')}); print(d['e2e']['value'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_f128_$v.err').read()[-3000:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"attention_fwd" -s 2 -c 2 --csv --log-file gpurun_out/attn_f128.csv python tools/profile_step.py 2 > gpurun_out/prof_f128.log 2>&1; echo "ncu exit $?"; grep -E "gpu__time|inst_exec|warps_active" gpurun_out/attn_f128.csv | cut -d, -f5,13- | cut -c1-30,90- | head
