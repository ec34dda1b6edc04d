"""ORACLE (test infrastructure): the data-parallel semantics of torch DistributedDataParallel, restated.

Reference: ``DistributedDataParallel(model, device_ids=[rank])`` (multi-gpu-distributed-cls.py:341).  Its Reducer
pre-divides each rank's gradient bucket by world_size and all-reduces with SUM
(SP/torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:18-33; reducer.hpp `div_factor_`), so every rank
steps with the mean of the per-rank gradients; parameters are broadcast from rank 0 at wrap time
(SP/torch/nn/parallel/distributed.py:879-889).  tests/golden/make_golden.py runs the real torch DDP on gloo
(world 2) to pin this restatement.
"""
import torch

from . import adamw_ref, bert_ref


def mean_grads(per_rank_grads):
    world = len(per_rank_grads)
    out = {}
    for k in per_rank_grads[0]:
        acc = torch.zeros_like(per_rank_grads[0][k])
        for g in per_rank_grads:           # pre-divide, then sum in rank order
            acc += g[k] / world
        out[k] = acc
    return out


def train(params, cfg, batches_per_step, lr=3e-5, weight_decay=0.01):
    """batches_per_step: list over steps of list over ranks of batch dicts.  Dropout off.  Updates `params` in place.
    Returns per step: dict(loss_per_rank, loss_mean, logits_per_rank, grads (averaged))."""
    opt = adamw_ref.HFAdamW(params, lr=lr, weight_decay=weight_decay)
    history = []
    for rank_batches in batches_per_step:
        losses, logits, grads = [], [], []
        for b in rank_batches:
            l, z, g = bert_ref.loss_and_grads(params, cfg, b)
            losses.append(l)
            logits.append(z)
            grads.append(g)
        avg = mean_grads(grads)
        history.append({"loss_per_rank": torch.stack(losses), "loss_mean": torch.stack(losses).mean(),
                        "logits_per_rank": logits, "grads": avg})
        opt.step(avg)
    return history
