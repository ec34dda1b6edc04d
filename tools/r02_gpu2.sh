#!/bin/bash
# round-2 GPU call 2 (1 GPU): gemm_ln probe, the failing tests with tracebacks, A/B of the fused LN and stream priority
export B2_PARITY_REPORT=$PWD/gpurun_out/r02_parity_report2.jsonl
rm -f $B2_PARITY_REPORT
timeout 300 python tools/gemm_ln_probe.py > gpurun_out/gemm_ln_probe.txt 2>&1; echo "probe rc $?"; cat gpurun_out/gemm_ln_probe.txt | tail -40
timeout 900 python -m pytest "tests/test_model.py::test_full_depth_step0_matches_ddp_fixture" "tests/test_trainer.py::test_config_a_loss_trajectory_vs_oracle" "tests/test_gemm.py::test_gemm_layernorm_cluster_kernel" -q --timeout 600 --tb=short --durations=8 > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest rc $?"; tail -60 gpurun_out/pytest_gpu2.log
for v in "B2_FUSED_LN=0" "B2_FUSED_LN=1" "B2_FUSED_LN=0 B2_STEP_PRIORITY=1"; do
  env $v timeout 300 python bench.py --steps 30 --warmup 5 --no-torch-eager --no-cpu-baseline --no-parity > "gpurun_out/bench_ab_$(echo $v | tr ' =' '__').json" 2> gpurun_out/bench_ab.err
  echo "$v rc $?"; python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(d['value'], d['ms_per_step'], d['e2e']['value'])" "gpurun_out/bench_ab_$(echo $v | tr ' =' '__').json"
done
