#!/bin/bash
# round-2 final 1-GPU validation: full GPU suite, smoke, bench lines of configs A / B / C, reference arm, ncu counters
export B2_PARITY_REPORT=$PWD/gpurun_out/r02_parity_report_final.jsonl
rm -f $B2_PARITY_REPORT
nvidia-smi > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --tb=short --durations=8 > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest rc $?"; tail -25 gpurun_out/pytest_gpu_final.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_a_final.json 2> gpurun_out/bench_a_final.err; echo "bench A rc $?"
timeout 600 python bench.py --config B --steps 20 > gpurun_out/bench_b_final.json 2> gpurun_out/bench_b_final.err; echo "bench B rc $?"
timeout 600 python bench.py --config C --steps 20 > gpurun_out/bench_c_final.json 2> gpurun_out/bench_c_final.err; echo "bench C rc $?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "bench ref rc $?"
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_counters_final.csv python tools/profile_step.py 3 > gpurun_out/r02_counters_final.log 2>&1
echo "ncu rc $?"
for n in a b c; do python -c "
import json,sys
d=json.load(open('gpurun_out/bench_%s_final.json' % sys.argv[1])); print(sys.argv[1], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity']['pass'], d['parity']['max_dloss'], 'roof', d['roofline']['frac'], d['roofline']['frac_warm'], 'eager', d['torch_eager'].get('value'), 'cpu', d['cpu_baseline']['value'], 'varlen', (d.get('varlen') or {}).get('packed_samples_per_s'))" $n; done
head -c 400 gpurun_out/bench_ref_final.json; echo
