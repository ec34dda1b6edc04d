#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_quick.log | cut -c1-200
timeout 300 python bench.py --steps 40 --warmup 5 --no-torch-eager --no-cpu-baseline --no-varlen > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench rc $?"; python -c "
import json
d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['pass'], d['parity']['max_dloss'], d['gpu_launches_per_step'])"
