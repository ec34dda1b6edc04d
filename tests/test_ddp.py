"""GPU, >= 2 devices: the peer-HBM gradient exchange + partitioned AdamW against the oracle's DDP restatement."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_peer_hbm_ddp_matches_oracle(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "tests", "ddp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mode eager OK" in r.stdout and "mode amp OK" in r.stdout and "mode fused OK" in r.stdout


def test_mp_spawn_launcher(world=2):
    """the -mp scripts' launcher: torch.multiprocessing.spawn + tcp:// rendezvous"""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ddp_worker.py"), "--spawn", str(world)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mode fused OK" in r.stdout


def test_bench_parity_block_config_a_world2():
    """bench.py at N = 2, config A at FULL size: the `parity` block (3 dropout-off optimizer steps through Trainer ->
    FusedTrainStep -> peer-HBM exchange + partitioned AdamW -> one-sided state_dict) must pass against
    tests/golden/config_a_ddp.pt, and the line must carry the driver-contract keys."""
    import json
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29579", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "5",
           "--warmup", "3", "--no-torch-eager"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    par = line["parity"]
    assert par["pass"] and par["world"] == 2 and par["shadow_identical"] and par["max_dloss"] <= par["tol"]["loss"], par
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["e2e"]["value"] > 0 and line["gpu_launches"] > 0
