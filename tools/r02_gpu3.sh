#!/bin/bash
# round-2 GPU call 3 (1 GPU): new gemm_ln kernel (probe + tests), the full GPU suite with durations, bench A A/B
export B2_PARITY_REPORT=$PWD/gpurun_out/r02_parity_report3.jsonl
rm -f $B2_PARITY_REPORT
timeout 300 python tools/gemm_ln_probe.py > gpurun_out/gemm_ln_probe2.txt 2>&1; echo "probe rc $?"; tail -30 gpurun_out/gemm_ln_probe2.txt
timeout 600 python -m pytest tests/test_gemm.py::test_gemm_layernorm_cluster_kernel tests/test_packing.py tests/test_kernels.py -q --timeout 300 --tb=short -x > gpurun_out/pytest_gpu3a.log 2>&1
echo "targeted pytest rc $?"; tail -25 gpurun_out/pytest_gpu3a.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short --durations=12 > gpurun_out/pytest_gpu3.log 2>&1
echo "full pytest rc $?"; tail -70 gpurun_out/pytest_gpu3.log
for v in "B2_STEP_PRIORITY=0" "B2_STEP_PRIORITY=1" "B2_FUSED_LN=0"; do
  env $v timeout 300 python bench.py --steps 30 --warmup 5 --no-torch-eager --no-cpu-baseline > "gpurun_out/bench3_$(echo $v | tr ' =' '__').json" 2> gpurun_out/bench3.err
  echo "$v rc $?"; python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['pass'], d['parity']['max_dloss'], d['parity']['max_dweight'], d['parity']['max_dnorm_rel'])" "gpurun_out/bench3_$(echo $v | tr ' =' '__').json"
done
