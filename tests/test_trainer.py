"""GPU: the reference's Trainer surface (train / dev / test, multi-gpu-distributed-cls.py:157-239) on the b200 step,
and the full-size per-tensor gradient parity table."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from parity import (TOL_GRAD_REL, b2, bert_ref, full_config, make_model, state_from_hf_init, tiny_config, to_dev)

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
    host cores).  Stated tolerance: rel-L2 <= 2e-2, except the attention query/key projections:
    <= 4e-2.  Their gradient is P * (dP - delta) with dP - delta = dO . (V_j - O_i): at initialisation the value rows of
    a sequence are nearly collinear, so the bf16 rounding of V (2^-9 relative) is amplified by |V| / |V_j - O_i| ~ 10.
    bf16 autocast of the reference has the same property; fp32 Q/K/V activations would be needed to remove it."""
    cfg = full_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    state = state_from_hf_init(cfg)
    model = make_model(cfg, state, cuda_dev).train()
    batch = bert_ref.synthetic_batch(cfg, 32, 128, 1000, padded=True)
    d = to_dev(batch, cuda_dev)
    out = model(input_ids=d["input_ids"], token_type_ids=d["token_type_ids"], attention_mask=d["attention_mask"],
                labels=d["label"])
    F.cross_entropy(out[1], d["label"]).backward()
    torch.cuda.synchronize()
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    _, _, ref = bert_ref.loss_and_grads(state, cfg, batch)
    got = model.grad_dict()
    scale = max(float(v.double().norm()) for v in ref.values())
    worst = {"qk": 0.0, "other": 0.0}
    for k, r in ref.items():
        rn = float(r.double().norm())
        if rn < 1e-6 * scale:          # analytically zero (key.bias: softmax shift invariance)
            assert float(got[k].double().norm()) < 1e-3 * scale, k
            continue
        rel = float((got[k].cpu().double() - r.double()).norm()) / rn
        qk = ".attention.self.query." in k or ".attention.self.key." in k
        worst["qk" if qk else "other"] = max(worst["qk" if qk else "other"], rel)
        assert rel <= (4e-2 if qk else TOL_GRAD_REL), (k, rel)
    print("worst rel-L2: q/k %.4f, others %.4f" % (worst["qk"], worst["other"]))
