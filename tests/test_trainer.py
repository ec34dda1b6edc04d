"""GPU: the reference's Trainer surface (train / dev / test, multi-gpu-distributed-cls.py:157-239) on the b200 step,
and the full-size per-tensor gradient parity table."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from parity import (TOL_GRAD_REL, TOL_TRAJ, b2, bert_ref, full_config, grad_tol, is_qk, make_model, report,
                    state_from_hf_init, tiny_config, to_dev)

pytestmark = pytest.mark.gpu


class _Loader:
    """stands in for DataLoader(collate_fn=Collate.collate_fn): yields the dict of int64 host tensors [:88-97]"""

    def __init__(self, cfg, n, batch, seed, padded=True):
        self.batches = [bert_ref.synthetic_batch(cfg, batch, 128, seed + i, padded=padded) for i in range(n)]

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


class _Sampler:
    def set_epoch(self, e):
        self.epoch = e


@pytest.mark.parametrize("fused", [True, False])
def test_trainer_train_dev_test(cuda_dev, tmp_path, capsys, fused):
    cfg = tiny_config()
    state = state_from_hf_init(cfg)
    model = make_model(cfg, state, cuda_dev)
    args = b2.Args()
    args.local_rank, args.local_world_size, args.rank = 0, 1, 0
    args.ckpt_path = str(tmp_path / "ckpt.pt")
    args.fused = fused
    args.dev, args.eval_step = True, 3
    loader, dev_loader = _Loader(cfg, 6, 4, 100), _Loader(cfg, 12, 4, 900)
    args.total_step = len(loader)
    opt = b2.build_optimizer(model, args)
    tr = b2.Trainer(args, cfg, model, torch.nn.CrossEntropyLoss(), opt)
    tr.train(loader, dev_loader, _Sampler())
    out = capsys.readouterr().out
    assert out.count("【train】") == 6 and "【dev】" in out and "耗时" in out
    assert os.path.exists(args.ckpt_path)                       # best-accuracy checkpoint [:192]
    sd = torch.load(args.ckpt_path)
    assert "classifier.weight" in sd and sd["classifier.weight"].dtype == torch.float32
    # the step really trained: weights moved away from the initial state
    assert float((model.state_dict()["classifier.weight"].cpu() - state["classifier.weight"]).abs().max()) > 0
    loss_total, acc = tr.dev(dev_loader)
    assert 0.0 <= acc <= 1.0 and float(loss_total) > 0
    # eval is deterministic (dropout off) and matches the oracle's eval forward
    model.eval()
    b = dev_loader.batches[0]
    with torch.no_grad():
        logits, label = tr.on_step(b)
    ref_state = {k: v.cpu() for k, v in model.state_dict().items() if k in state}
    _, rz = bert_ref.forward(ref_state, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"], b["label"])
    assert float((logits.cpu() - rz).abs().max()) <= 1e-2
    # the graph-captured eval forward (used by dev/test when args.fused) is the eager one, bit for bit, on every replay
    ev = b2.FusedEvalStep(model, 4, 128)
    for i in (0, 1, 2, 3, 0):
        bi = dev_loader.batches[i]
        flog, flab, floss = ev(bi)
        with torch.no_grad():
            elog, elab = tr.on_step(bi)
        assert torch.equal(flog, elog) and torch.equal(flab, elab)
        assert abs(float(floss) - float(F.cross_entropy(elog, elab))) <= 1e-6
    # like the reference, Trainer.test hands sklearn the 6 label names: every class must occur in y_true U y_pred
    assert len({int(v) for b_ in dev_loader.batches for v in b_["label"]}) == 6
    report = tr.test(model, dev_loader, ["c%d" % i for i in range(6)])
    assert "precision" in report
    # reload the checkpoint into a fresh model (test.py:96-101 style) and into the optimizer-free eval path
    m2 = b2.BertForSequenceClassification(cfg)
    m2.load_state_dict(sd)
    m2.cuda().eval()
    with torch.no_grad():
        d = to_dev(b, cuda_dev)
        out2 = m2(input_ids=d["input_ids"], token_type_ids=d["token_type_ids"], attention_mask=d["attention_mask"])
    assert out2.logits.shape == (4, 6)


def test_full_size_gradients_per_tensor_vs_oracle(cuda_dev):
    """BASELINE config A, dropout off: every one of the 201 gradient tensors against the fp32 oracle (run here on the
    host cores).  Stated tolerance: rel-L2 <= 2e-2 for EVERY tensor.  The attention query / key
    projections are the sensitive ones (P * (dP - delta) with dP - delta = dO . (V_j - O_i): nearly collinear value rows
    at initialisation amplify every upstream rounding ~10x); they meet 2e-2 because the forward keeps the residual stream
    in fp32 (with a bf16 stream they reached 3.1e-2 in layer 11).  For context the same tensors of stock HF BERT (eager
    attention, cuBLAS) under torch.autocast(bf16) are measured against the same oracle on the same batch and reported."""
    from oracle import cpu_step
    cfg = full_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    state = state_from_hf_init(cfg)
    model = make_model(cfg, state, cuda_dev).train()
    batch = bert_ref.synthetic_batch(cfg, 32, 128, 1000, padded=True)
    d = to_dev(batch, cuda_dev)
    out = model(input_ids=d["input_ids"], token_type_ids=d["token_type_ids"], attention_mask=d["attention_mask"],
                labels=d["label"])
    F.cross_entropy(out[1], d["label"]).backward()
    torch.cuda.synchronize()
    torch.set_num_threads(cpu_step.usable_cores())   # affinity capped by the cgroup CPU quota
    _, _, ref = bert_ref.loss_and_grads(state, cfg, batch)
    got = model.grad_dict()
    # the comparison stack: HF's own module, bf16 autocast (fp32 parameters and gradients, bf16 matmuls, fp32 softmax)
    hf = cpu_step.build_hf_model(cfg, seed=123)
    hf.load_state_dict(state, strict=False)
    hf.to(cuda_dev).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ho = hf(input_ids=d["input_ids"], token_type_ids=d["token_type_ids"], attention_mask=d["attention_mask"])
        hl = F.cross_entropy(ho.logits.float(), d["label"])
    hl.backward()
    auto = {k: v.grad.detach() for k, v in hf.named_parameters()}
    scale = max(float(v.double().norm()) for v in ref.values())
    worst = {"qk": 0.0, "other": 0.0, "auto_qk": 0.0, "auto_other": 0.0}
    rows = []
    for k, r in ref.items():
        rn = float(r.double().norm())
        if rn < 1e-6 * scale:          # analytically zero (key.bias: softmax shift invariance)
            assert float(got[k].double().norm()) < 1e-3 * scale, k
            continue
        rel = float((got[k].cpu().double() - r.double()).norm()) / rn
        rel_a = float((auto[k].cpu().double() - r.double()).norm()) / rn
        grp = "qk" if is_qk(k) else "other"
        worst[grp] = max(worst[grp], rel)
        worst["auto_" + grp] = max(worst["auto_" + grp], rel_a)
        rows.append((k, round(rel, 5), round(rel_a, 5)))
        assert rel <= TOL_GRAD_REL, (k, rel)
    print("worst rel-L2 vs fp32 oracle: q/k %.4f (torch bf16 autocast: %.4f), others %.4f (autocast: %.4f)"
          % (worst["qk"], worst["auto_qk"], worst["other"], worst["auto_other"]))
    report("config_a_grads", {"worst": worst, "qk_rows": [r for r in rows if is_qk(r[0])]})
    assert worst["other"] <= 1.25 * worst["auto_other"], worst      # and no worse than autocast where it is not amplified


@pytest.mark.parametrize("dropout,steps", [(False, 20), (True, 8)])
def test_config_a_loss_trajectory_vs_oracle(cuda_dev, dropout, steps):
    """BASELINE.md §4 "loss trajectory, 20 steps": config A (B=32, S=128, HF AdamW lr 3e-5), the fused CUDA-graph step
    against the fp32 oracle stepping the same batches on the host cores.  dropout=True: the reference's training mode,
    the oracle replays the kernels' Philox keep-masks (fewer steps: the numpy replica draws 150 M mask bits per step).
    For dropout=False the same trajectory is also run by stock HF BERT under torch.autocast(bf16) with the restated HF
    AdamW -- the deviation bf16 compute itself causes on this chaotic early trajectory (AdamW's first updates are
    lr * sign(g) on every one of 102 M weights) -- and ours must stay within max(1e-2, 1.25x that)."""
    from oracle import adamw_ref, cpu_step
    from parity import oracle_masks
    kw = {} if dropout else dict(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    cfg = full_config(**kw)
    state = state_from_hf_init(cfg)
    B, S, seed = 32, 128, 777
    batches = [bert_ref.synthetic_batch(cfg, B, S, 8000 + i, padded=(i % 2 == 1)) for i in range(steps)]

    class A:
        weight_decay, learning_rate = 0.01, 3e-5

    model = make_model(cfg, state, cuda_dev).train()
    model._engine.seed_dropout(seed, 0)
    opt = b2.build_optimizer(model, A)
    step = b2.FusedTrainStep(model, opt, B, S)
    ours = []
    for b in batches:
        step(b)
        ours.append(step.loss_to_host())
    assert step.graph is not None

    auto = None
    if not dropout:
        hf = cpu_step.build_hf_model(cfg, seed=123)
        hf.load_state_dict(state, strict=False)
        hf.to(cuda_dev).train()
        hopt = cpu_step._HFOpt(hf, 3e-5, 0.01)
        auto = []
        for b in batches:
            d = to_dev(b, cuda_dev)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ho = hf(input_ids=d["input_ids"], token_type_ids=d["token_type_ids"],
                        attention_mask=d["attention_mask"])
                hl = F.cross_entropy(ho.logits.float(), d["label"])
            hopt.zero_grad()
            hl.backward()
            hopt.step()
            auto.append(float(hl))
        del hf, hopt

    torch.set_num_threads(cpu_step.usable_cores())   # affinity capped by the cgroup CPU quota
    params = {k: v.clone() for k, v in state.items()}
    ropt = adamw_ref.HFAdamW(params, lr=3e-5, weight_decay=0.01)
    ref = []
    for s, b in enumerate(batches):
        masks = oracle_masks(cfg, B, S, seed, s) if dropout else None
        l, _z, g = bert_ref.loss_and_grads(params, cfg, b, masks=masks)
        ropt.step(g)
        ref.append(float(l))
    dev_ours = [abs(a - r) for a, r in zip(ours, ref)]
    dev_auto = [abs(a - r) for a, r in zip(auto, ref)] if auto is not None else None
    report("config_a_trajectory", {"dropout": dropout, "steps": steps, "ref": ref, "ours": ours, "autocast": auto,
                                   "max_dev_ours": max(dev_ours), "max_dev_autocast": max(dev_auto) if dev_auto else None})
    print("trajectory (dropout=%s): max |dloss| ours %.4f%s" % (dropout, max(dev_ours),
          "" if dev_auto is None else ", torch bf16 autocast %.4f" % max(dev_auto)))
    bound = TOL_TRAJ if dev_auto is None else max(TOL_TRAJ, 1.25 * max(dev_auto))
    assert max(dev_ours) <= bound, (dev_ours, dev_auto)
