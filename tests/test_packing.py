"""Token packing: the reference pads every row to max_seq_len = 128 (multi-gpu-distributed-cls.py:76)
while its data averages 18 tokens per row.  CPU: the packer's invariants.  GPU: a packed step equals the padded step it
replaces -- against the fp32 oracle run on the PADDED batch."""
import pytest
import torch
import torch.nn.functional as F

from parity import (TOL_LOGITS, TOL_LOSS, TOL_TRAJ, assert_grads_within_tolerance, b2, bert_ref, full_config, make_model,
                    state_from_hf_init, tiny_config, to_dev)
from pytorch_distributed_nlp_b200.packing import pack_batch


def short_batch(cfg, B, seed, lo=3, hi=40, S=128):
    """a batch shaped like the reference's: right-padded to 128, valid lengths ~ U{lo..hi}"""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(lo, hi + 1, (B,), generator=g)
    ids = torch.randint(1, cfg.vocab_size, (B, S), generator=g, dtype=torch.int64)
    ids[:, 0] = min(101, cfg.vocab_size - 1)
    mask = (torch.arange(S)[None] < lens[:, None]).to(torch.int64)
    tt = torch.randint(0, 2, (B, S), generator=g, dtype=torch.int64) * mask
    return {"input_ids": ids * mask, "token_type_ids": tt, "attention_mask": mask,
            "label": torch.randint(0, cfg.num_labels, (B,), generator=g, dtype=torch.int64)}


def test_pack_batch_invariants():
    cfg = tiny_config()
    b = short_batch(cfg, 37, 5)
    p = pack_batch(b["input_ids"], b["token_type_ids"], b["attention_mask"])
    lens = b["attention_mask"].sum(1)
    NB = p["bins"]
    assert NB == p["input_ids"].shape[0] and NB <= 37 and NB * 128 >= int(lens.sum())
    assert NB <= -(-int(lens.sum()) // 128) + 1                  # first-fit-decreasing wastes at most ~ one bin here
    seen = torch.zeros(NB, 128, dtype=torch.bool)
    for i in range(37):
        r = int(p["cls_index"][i])
        k, lo = divmod(r, 128)
        n = int(lens[i])
        assert lo + n <= 128 and not seen[k, lo:lo + n].any()     # inside one bin, no overlap
        seen[k, lo:lo + n] = True
        assert torch.equal(p["input_ids"][k, lo:lo + n], b["input_ids"][i, :n])
        assert torch.equal(p["token_type_ids"][k, lo:lo + n], b["token_type_ids"][i, :n])
        assert torch.equal(p["position_ids"][k, lo:lo + n], torch.arange(n))
        seg = p["segments"][k, lo:lo + n]
        assert bool(((seg & 0xffff) == lo).all()) and bool(((seg >> 16) == lo + n).all())
    # unused rows: pad id, a one-row segment of their own (so no softmax row is fully masked)
    un = ~seen
    assert bool((p["input_ids"][un] == 0).all())
    rows = torch.arange(128).repeat(NB, 1)[un]
    assert bool(((p["segments"][un] & 0xffff) == rows).all()) and bool(((p["segments"][un] >> 16) == rows + 1).all())
    # full-length rows pack one per bin
    full = bert_ref.synthetic_batch(cfg, 3, 128, 1)
    assert pack_batch(full["input_ids"], full["token_type_ids"], full["attention_mask"])["bins"] == 3
    bad = b["attention_mask"].clone()
    bad[0, 0] = 0
    with pytest.raises(ValueError, match="prefix"):
        pack_batch(b["input_ids"], b["token_type_ids"], bad)
    with pytest.raises(ValueError, match="more than"):
        pack_batch(torch.ones(1, 256, dtype=torch.int64), None, torch.ones(1, 256, dtype=torch.int64))


def _run(model, dev, batch, packed=None):
    if packed is None:
        d = to_dev(batch, dev)
        out = model(input_ids=d["input_ids"], token_type_ids=d["token_type_ids"], attention_mask=d["attention_mask"],
                    labels=d["label"])
    else:
        out = model(input_ids=packed["input_ids"].to(dev), token_type_ids=packed["token_type_ids"].to(dev),
                    labels=batch["label"].to(dev), position_ids=packed["position_ids"].to(dev),
                    segments=packed["segments"].to(dev), cls_index=packed["cls_index"].to(dev))
    loss = F.cross_entropy(out[1], batch["label"].to(dev))
    loss.backward()
    torch.cuda.synchronize()
    return out, loss


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["tiny", "config-A"])
def test_packed_step_equals_padded_step(cuda_dev, which):
    """dropout off: logits, loss and every gradient of the packed step against the fp32 oracle on the PADDED batch
    (stated tolerances), and against the padded CUDA step.  config-A runs the hidden-768 path (cluster LayerNorm)."""
    kw = dict(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    cfg, B = (tiny_config(**kw), 13) if which == "tiny" else (full_config(num_hidden_layers=3, **kw), 32)
    state = state_from_hf_init(cfg)
    batch = short_batch(cfg, B, 21)
    packed = pack_batch(batch["input_ids"], batch["token_type_ids"], batch["attention_mask"])
    assert packed["bins"] < B
    rl, rz, rg = bert_ref.loss_and_grads(state, cfg, batch)
    model = make_model(cfg, state, cuda_dev).train()
    out_p, loss_p = _run(model, cuda_dev, batch, packed)
    g_packed = model.grad_dict()
    assert out_p[1].shape == (B, cfg.num_labels)
    assert abs(float(loss_p) - float(rl)) <= TOL_LOSS and abs(float(out_p[0]) - float(rl)) <= TOL_LOSS
    assert float((out_p[1].detach().cpu() - rz).abs().max()) <= TOL_LOGITS
    assert_grads_within_tolerance(g_packed, rg)
    out_d, loss_d = _run(model, cuda_dev, batch)                 # the padded step on the same weights
    assert float((out_d[1].detach() - out_p[1].detach()).abs().max()) <= TOL_LOGITS
    assert abs(float(loss_d) - float(loss_p)) <= TOL_LOSS
    # eval forward, packed, without labels
    model.eval()
    with torch.no_grad():
        ev = model(input_ids=packed["input_ids"].to(cuda_dev), token_type_ids=packed["token_type_ids"].to(cuda_dev),
                   position_ids=packed["position_ids"].to(cuda_dev), segments=packed["segments"].to(cuda_dev),
                   cls_index=packed["cls_index"].to(cuda_dev))
    assert float((ev.logits.cpu() - rz).abs().max()) <= TOL_LOGITS
    with pytest.raises(ValueError, match="together"):
        model(input_ids=packed["input_ids"].to(cuda_dev), segments=packed["segments"].to(cuda_dev))


@pytest.mark.gpu
def test_trainer_with_packing_follows_the_oracle(cuda_dev):
    """Trainer(args.pack = True): the reference's padded host batches are packed on the host, each bin count gets its own
    captured step, and the 5-step loss trajectory is the oracle's on the PADDED batches."""
    from oracle import ddp_ref
    cfg = tiny_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    state = state_from_hf_init(cfg)
    batches = [short_batch(cfg, 16, 300 + i, hi=(30 if i % 2 else 90)) for i in range(5)]
    ref = {k: v.clone() for k, v in state.items()}
    hist = ddp_ref.train(ref, cfg, [[b] for b in batches])
    model = make_model(cfg, state, cuda_dev)
    args = b2.Args()
    args.local_rank, args.local_world_size, args.rank, args.pack = 0, 1, 0, True
    opt = b2.build_optimizer(model, args)
    tr = b2.Trainer(args, cfg, model, torch.nn.CrossEntropyLoss(), opt)
    for i, b in enumerate(batches):
        loss = float(tr.train_step(b))
        assert abs(loss - float(hist[i]["loss_mean"])) <= TOL_TRAJ, (i, loss, float(hist[i]["loss_mean"]))
    assert len(tr._packed) >= 2                                   # two length regimes -> two bin counts -> two graphs
    sd = model.state_dict()
    for k, v in ref.items():
        assert float((sd[k].cpu() - v).abs().max()) <= 2e-4, k
