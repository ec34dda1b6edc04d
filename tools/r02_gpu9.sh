#!/bin/bash
# round-2 GPU call 9 (8 GPUs): config A (both exchange forms), B and C at 8 ranks, each with its parity block
run() { # name, env, args...
  name=$1; envs=$2; shift 2
  env $envs timeout 600 python bench.py --gpus 8 "$@" > gpurun_out/bench9_$name.json 2> gpurun_out/bench9_$name.err
  echo "$name rc $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench9_$name.err | tail -3 | cut -c1-300
  python - "$name" << 'PY'
import json, sys
name = sys.argv[1]
txt = open("gpurun_out/bench9_%s.json" % name).read()
lines = [l for l in txt.splitlines() if l.startswith("{")]
if lines:
    d = json.loads(lines[-1])
    print(name, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", {k: d["parity"].get(k) for k in ("pass", "max_dloss", "max_dloss_mean", "max_dweight", "shadow_identical")} if d.get("parity") else None)
    print("   varlen", d.get("varlen"), "eager", d.get("torch_eager"))
PY
}
run a_dma1 "B2_DDP_DMA=1" --steps 30 --warmup 5
run a_dma0 "B2_DDP_DMA=0" --steps 30 --warmup 5 --no-torch-eager --no-varlen
run b "B2_DDP_DMA=1" --config B --steps 20 --warmup 5
run c "B2_DDP_DMA=1" --config C --steps 20 --warmup 5
run c_dma0 "B2_DDP_DMA=0" --config C --steps 20 --warmup 5 --no-torch-eager
