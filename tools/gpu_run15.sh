#!/bin/bash
# 2 GPUs: DDP parity (torchrun + mp.spawn launchers; eager / amp / fused loops) and the 2-GPU bench line
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_ddp.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_ddp.log 2>&1; echo "pytest ddp exit $?"; tail -5 gpurun_out/pytest_ddp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 exit $?"; tail -c 1500 gpurun_out/bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "bench2 ref exit $?"; tail -c 600 gpurun_out/bench_2gpu_ref.json
