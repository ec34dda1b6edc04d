#!/bin/bash
# first GPU bring-up: each test file in its own process (a trapped kernel poisons a CUDA context), logs in gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tools/gemm_debug.py > gpurun_out/gemm_debug.log 2>&1; echo "gemm_debug exit $?" | tee -a gpurun_out/summary.txt
for f in test_gemm test_kernels test_model; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log | tee -a gpurun_out/summary.txt
done
