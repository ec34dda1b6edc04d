"""ORACLE / CPU baseline (test + bench infrastructure): the reference's ``single-gpu-cls.py`` training-loop body on
host cores.

The reference script itself cannot be imported here: its first statement ``from transformers import ... AdamW``
(single-gpu-cls.py:11) fails on the installed transformers 5.5, and its ``Args.device`` is hard-wired to "cuda"
(:198).  What is timed is therefore a port ("kind": "port") of its loop body (:131-140):
    output = model(input_ids, token_type_ids, attention_mask, labels)   # HF BertForSequenceClassification, eager
    loss   = CrossEntropyLoss()(output[1], label)
    optimizer.zero_grad(); loss.backward(); optimizer.step()            # HF AdamW (restated, adamw_ref)
    loss.item()
with the real HF model class from the installed transformers, fp32, dropout ON (model.train()), all host threads.
"""
import time

import torch

from . import adamw_ref, bert_ref


def hf_config(cfg):
    from transformers import BertConfig
    return BertConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                      num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                      intermediate_size=cfg.intermediate_size, max_position_embeddings=cfg.max_position_embeddings,
                      type_vocab_size=cfg.type_vocab_size, hidden_dropout_prob=cfg.hidden_dropout_prob,
                      attention_probs_dropout_prob=cfg.attention_probs_dropout_prob,
                      layer_norm_eps=cfg.layer_norm_eps, hidden_act="gelu", num_labels=cfg.num_labels,
                      attn_implementation="eager")


def build_hf_model(cfg, seed=123):
    """The arithmetic the reference runs: HF BertForSequenceClassification, eager attention, fp32, HF init."""
    from transformers import BertForSequenceClassification
    torch.manual_seed(seed)
    return BertForSequenceClassification(hf_config(cfg))


def usable_cores():
    """Host cores this process may actually use: affinity mask capped by the cgroup CPU quota (a container that
    reports 128 CPUs but is throttled to a few thrashes when torch spawns 128 threads)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


class _HFOpt:
    """adamw_ref.HFAdamW driven from module parameters (the per-tensor python loop of the original)."""

    def __init__(self, model, lr, weight_decay):
        self.named = dict(model.named_parameters())
        self.inner = adamw_ref.HFAdamW({k: v.data for k, v in self.named.items()}, lr=lr, weight_decay=weight_decay)

    def zero_grad(self):
        for p in self.named.values():
            p.grad = None

    def step(self):
        self.inner.step({k: v.grad for k, v in self.named.items()})


def time_steps(cfg, batch_size, seq_len, steps, warmup, threads=None, seed=1000):
    """Returns dict(samples_per_s, ms_per_step, cores, losses)."""
    if threads is None:
        threads = usable_cores()
    torch.set_num_threads(threads)
    model = build_hf_model(cfg)
    model.train()
    opt = _HFOpt(model, 3e-5, 0.01)
    crit = torch.nn.CrossEntropyLoss()
    ring = [bert_ref.synthetic_batch(cfg, batch_size, seq_len, seed + i) for i in range(max(1, min(4, steps)))]
    losses, times = [], []
    for i in range(warmup + steps):
        b = ring[i % len(ring)]
        t0 = time.perf_counter()
        out = model(input_ids=b["input_ids"], token_type_ids=b["token_type_ids"],
                    attention_mask=b["attention_mask"], labels=b["label"])
        loss = crit(out[1], b["label"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        lv = loss.item()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            losses.append(lv)
    total = sum(times)
    return {"samples_per_s": batch_size * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "cores": threads, "losses": losses}
