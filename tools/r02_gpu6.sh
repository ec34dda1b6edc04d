#!/bin/bash
# round-2 GPU call 6 (1 GPU): does a co-resident optimizer hide?  8-warp GEMM epilogue (frees 16 K registers per SM)
# with and without the optimizer
for v in "B2_GEMM_EPI_WARPS=8 B2_FUSED_LN=0" "B2_GEMM_EPI_WARPS=8 B2_FUSED_LN=0 B2_DEBUG_SKIP_ADAMW=1" "B2_FUSED_LN=0" "B2_FUSED_LN=0 B2_DEBUG_SKIP_ADAMW=1"; do
  env $v timeout 300 python bench.py --steps 30 --warmup 5 --no-torch-eager --no-cpu-baseline --no-parity --no-varlen > "gpurun_out/bench6_$(echo $v | tr ' =' '__').json" 2> gpurun_out/bench6.err
  echo "$v rc $?"; python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(d['value'], d['ms_per_step'], d['e2e']['value'])" "gpurun_out/bench6_$(echo $v | tr ' =' '__').json"
done
