#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model.py tests/test_kernels.py tests/test_trainer.py -q -m gpu -p no:cacheprovider -x -k "not full_size" > gpurun_out/pytest_quick.log 2>&1; echo "pytest quick exit $?"; tail -15 gpurun_out/pytest_quick.log
timeout 600 ncu --section Occupancy --section LaunchStats --section SpeedOfLight --clock-control none -k regex:attention_bwd_kernel -s 2 -c 2 --csv --log-file gpurun_out/attn_bwd_occ.csv python tools/profile_step.py 2 > gpurun_out/prof_occ.log 2>&1; echo "ncu occ exit $?"
grep -E "Theoretical Occupancy|Achieved Occupancy|Block Limit|Duration|Shared Memory Config|Dynamic Shared|Registers Per" gpurun_out/attn_bwd_occ.csv | cut -d, -f5,13- | head -40
timeout 600 python tests/context_torch_eager.py > gpurun_out/torch_eager.json 2> gpurun_out/torch_eager.err; echo "torch eager exit $?"; cat gpurun_out/torch_eager.json; tail -3 gpurun_out/torch_eager.err
