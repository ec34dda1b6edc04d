#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_full.py 12 > gpurun_out/diag_full.log 2>&1; echo "diag exit $?"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 800 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 4 > gpurun_out/profile_step.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 60 -c 6 -o gpurun_out/prof_gemm python tools/profile_step.py 2 > gpurun_out/prof_gemm.log 2>&1; echo "ncu full exit $?"
