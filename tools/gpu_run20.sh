#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 --csv --log-file gpurun_out/launches_r01j.csv python tools/profile_step.py 2 > gpurun_out/prof_j.log 2>&1; echo "ncu list exit $?"
