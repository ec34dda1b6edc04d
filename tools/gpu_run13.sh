#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm.py tests/test_kernels.py tests/test_model.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest quick exit $?"; tail -3 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_f.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_f.err').read()[-3000:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_bwd|layernorm_bwd" -s 4 -c 3 -o gpurun_out/prof_bwd python tools/profile_step.py 2 > gpurun_out/prof_bwd.log 2>&1; echo "ncu exit $?"
B2_LN_BWD_STAGED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"layernorm_bwd" -s 4 -c 1 -o gpurun_out/prof_ln_reg python tools/profile_step.py 2 > gpurun_out/prof_ln_reg.log 2>&1; echo "ncu exit $?"
