#!/bin/bash
# round-2 8-GPU call: config A (default = fused peer-HBM exchange kernel) with every leg, config B at 8 ranks
run() { # name, env, args...
  name=$1; envs=$2; shift 2
  env $envs timeout 600 python bench.py --gpus 8 "$@" > gpurun_out/bench9_$name.json 2> gpurun_out/bench9_$name.err
  echo "$name rc $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|UserWarning\|detach\|last = float" gpurun_out/bench9_$name.err | tail -3 | cut -c1-300
  python - "$name" << 'PY'
import json, sys
name = sys.argv[1]
txt = open("gpurun_out/bench9_%s.json" % name).read()
lines = [l for l in txt.splitlines() if l.startswith("{")]
if lines:
    d = json.loads(lines[-1])
    print(name, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", {k: d["parity"].get(k) for k in ("pass", "max_dloss", "max_dloss_mean", "max_dweight", "shadow_identical")} if d.get("parity") else None)
    print("   varlen", d.get("varlen"), "eager", d.get("torch_eager"), "exchange", d.get("exchange"))
PY
}
run a_final "B2_DDP_DMA=0" --steps 40 --warmup 5
run b_final "B2_DDP_DMA=0" --config B --steps 20 --warmup 5 --no-torch-eager
