"""Synthetic inputs of the benchmark (BASELINE.json `north_star`: no dataset, no tokenizer).

Stands in for what the reference's ``Collate.collate_fn`` yields (multi-gpu-distributed-cls.py:88-97): a dict of int64
host tensors ``input_ids / token_type_ids / attention_mask [batch, seq]`` and ``label [batch]``.
"""
import torch


def synthetic_batch(cfg, batch, seq, seed, padded=False, device="cpu"):
    """ids ~ U{0..vocab-1} with [:,0] = 101 (CLS), token types 0, labels ~ U{0..C-1}, generator seed `seed`;
    `padded` draws per-row valid lengths ~ U{8..seq} and zeroes ids / mask beyond them (the padded parity variant)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, cfg.vocab_size, (batch, seq), generator=g, dtype=torch.int64)
    ids[:, 0] = min(101, cfg.vocab_size - 1)
    lab = torch.randint(0, cfg.num_labels, (batch,), generator=g, dtype=torch.int64)
    mask = torch.ones(batch, seq, dtype=torch.int64)
    if padded:
        lens = torch.randint(8, seq + 1, (batch,), generator=g)
        ar = torch.arange(seq)[None]
        mask = (ar < lens[:, None]).to(torch.int64)
        ids = ids * mask
    tt = torch.zeros(batch, seq, dtype=torch.int64)
    return {"input_ids": ids.to(device), "token_type_ids": tt.to(device), "attention_mask": mask.to(device),
            "label": lab.to(device)}
