#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_fwd128|attention_bwd" -s 2 -c 2 -o gpurun_out/prof_attn2 python tools/profile_step.py 2 > gpurun_out/prof_attn2.log 2>&1; echo "ncu exit $?"
