#!/bin/bash
# round-end style pass: full GPU suite, bench (both arms), smoke, ncu artefacts for profiles/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x -s --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; grep -E "worst rel|passed|failed|Error" gpurun_out/pytest_gpu.log | tail -8
timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','loss')}); print(d['e2e']); r=d['roofline']; print(r['achieved'],r['frac'],r['traffic'],r['step_achieved_tflops_per_gpu']); print(d['cpu_baseline']); print(d['clocks'])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"; tail -c 400 gpurun_out/bench_ref.json
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 222 -c 222 --csv --log-file gpurun_out/launches_r01n.csv python tools/profile_step.py 3 > gpurun_out/prof_n.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm2_bf16_kernel|gemm2_grouped" -s 30 -c 10 -o gpurun_out/prof_gemm_n python tools/profile_step.py 2 > gpurun_out/prof_gemm_n.log 2>&1; echo "ncu full exit $?"
