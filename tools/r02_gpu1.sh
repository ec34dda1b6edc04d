#!/bin/bash
# round-2 GPU call 1 (1 GPU): full GPU test suite with the parity report, bench lines of configs A / B / C, and the
# per-kernel ncu counter pass of one training step.  Outputs under gpurun_out/.
export B2_PARITY_REPORT=$PWD/gpurun_out/r02_parity_report.jsonl
rm -f $B2_PARITY_REPORT
nvidia-smi > gpurun_out/gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench A rc $?"
timeout 900 python bench.py --config B --steps 20 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench B rc $?"
timeout 900 python bench.py --config C --steps 20 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench C rc $?"
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_counters.csv python tools/profile_step.py 3 > gpurun_out/r02_counters.log 2>&1
echo "ncu rc $?"
head -c 600 gpurun_out/bench_a.json; echo; tail -3 gpurun_out/bench_a.err
head -c 400 gpurun_out/bench_b.json; echo; tail -3 gpurun_out/bench_b.err
head -c 400 gpurun_out/bench_c.json; echo; tail -3 gpurun_out/bench_c.err
