#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels.py tests/test_model.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest quick exit $?"; tail -3 gpurun_out/pytest_quick.log
for st in 1 0; do
B2_LN_BWD_PAIR=$st timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ln$st.json 2> gpurun_out/bench_ln$st.err; echo "bench pair=$st exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_ln$st.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e']['value'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_ln$st.err').read()[-3000:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"layernorm_bwd" -s 4 -c 2 --csv --log-file gpurun_out/ln_pair.csv python tools/profile_step.py 2 > gpurun_out/prof_ln.log 2>&1; echo "ncu exit $?"; grep -E "gpu__time|inst_exec" gpurun_out/ln_pair.csv | cut -d, -f5,13- | head
