#!/bin/bash
# last 1-GPU validation of the round: full GPU suite, smoke, default bench line (config A), short B / C lines
timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 300 --tb=short > gpurun_out/pytest_gpu_last.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu_last.log | cut -c1-200
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_last.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke_last.log
timeout -k 5 400 python bench.py > gpurun_out/bench_a_last.json 2> gpurun_out/bench_a_last.err; echo "bench A rc $?"
timeout -k 5 300 python bench.py --config B --steps 20 --no-cpu-baseline > gpurun_out/bench_b_last.json 2> gpurun_out/bench_b_last.err; echo "bench B rc $?"
timeout -k 5 300 python bench.py --config C --steps 20 --no-cpu-baseline > gpurun_out/bench_c_last.json 2> gpurun_out/bench_c_last.err; echo "bench C rc $?"
for n in a b c; do python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/bench_%s_last.json' % sys.argv[1]) if l.startswith('{')][-1]); print(sys.argv[1], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity']['pass'], d['parity']['max_dloss'], 'roof', d['roofline']['frac'], d['roofline']['frac_warm'], 'eager', (d.get('torch_eager') or {}).get('value'), 'varlen', (d.get('varlen') or {}).get('packed_samples_per_s'))" $n; done
