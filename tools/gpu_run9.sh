#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x -s --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; grep -E "worst rel|passed|failed|Error" gpurun_out/pytest_gpu.log | tail -8
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','loss')}); print(d['e2e']); r=d['roofline']; print(r['achieved'],r['frac'],r['step_achieved_tflops_per_gpu'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 40 -c 8 -o gpurun_out/prof_gemm2 python tools/profile_step.py 2 > gpurun_out/prof_gemm2.log 2>&1; echo "ncu full exit $?"
