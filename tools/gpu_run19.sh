#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels.py tests/test_model.py tests/test_gemm.py tests/test_trainer.py -k "not full_size" -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest quick exit $?"; tail -3 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_i.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_i.err').read()[-3000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"head_bwd|layernorm_bwd|accum_finish" -s 0 -c 8 --csv --log-file gpurun_out/attn_i.csv python tools/profile_step.py 2 > gpurun_out/prof_attn_i.log 2>&1; echo "ncu exit $?"; grep -E "gpu__time|inst_exec" gpurun_out/attn_i.csv | cut -d, -f5,13- | cut -c1-40,100- | head
