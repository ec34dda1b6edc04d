"""Token packing for the reference's real input shape (§8 f3 of SURVEY.md).

The reference tokenises with ``padding="max_length", max_length=128`` (multi-gpu-distributed-cls.py:76) while the rows
of data/train.json average 18 characters: ~85 % of every [batch, 128] input is padding that the step still pays full
price for.  `pack_batch` re-arranges such a batch into fewer 128-token BINS: the valid prefixes of several sequences
share one bin (first-fit, longest first), every token keeps the position id it had in its own sequence, and every bin
row carries the [lo, hi) range of its own sequence inside the bin.  The attention kernels then mask with that range
(a block-diagonal mask per bin) instead of the key-padding mask, every other kernel is token-wise and simply sees
fewer rows, and the pooler reads each sequence's first token through `cls_index`.  Per sequence the arithmetic is
exactly that of the padded batch (a padded key contributes exp(-3.4e38 - m) = 0 to its softmax row, like a key of
another sequence here).
"""
import numpy as np
import torch

BIN = 128


def pack_batch(input_ids, token_type_ids, attention_mask, bin_len=BIN):
    """input_ids / token_type_ids / attention_mask: int64 [B, S] host tensors as the reference's Collate yields them
    (valid tokens first: the tokenizer pads on the right).  Returns a dict of host tensors:
      input_ids, token_type_ids, position_ids   int64 [NB, bin_len]   (unused bin rows: pad id 0, position 0)
      segments                                   int32 [NB, bin_len]   lo | hi << 16: the row's own sequence is
                                                                       rows [lo, hi) of its bin (an unused row: itself)
      cls_index                                  int64 [B]             flat row (bin * bin_len + lo) of sequence b's
                                                                       first token, in the ORIGINAL batch order
      lengths                                    int64 [B]
    NB <= B; a sequence longer than bin_len is not supported (the reference truncates to max_seq_len = 128)."""
    if input_ids.dim() != 2:
        raise ValueError("input_ids must be [batch, seq]")
    B, S = input_ids.shape
    ids_np = input_ids.numpy()
    mask_np = (attention_mask.numpy() != 0) if attention_mask is not None else np.ones((B, S), dtype=bool)
    lens = mask_np.sum(1).astype(np.int64)
    # valid tokens must form a prefix (right padding): row i is exactly `lens[i]` ones followed by zeros
    if not np.array_equal(mask_np, np.arange(S)[None, :] < lens[:, None]):
        raise ValueError("pack_batch: attention_mask is not a right-padded prefix mask")
    if int(lens.min()) < 1:
        raise ValueError("pack_batch: empty sequence (no valid token)")
    if int(lens.max()) > bin_len:
        raise ValueError("pack_batch: a sequence has %d valid tokens, more than the %d-token bin"
                         % (int(lens.max()), bin_len))
    # first-fit, longest first (ties in batch order): a few dozen integers -- plain Python
    ll = lens.tolist()
    order = sorted(range(B), key=lambda i: (-ll[i], i))
    free, bin_of, lo_of = [], [0] * B, [0] * B
    for i in order:
        n = ll[i]
        for k in range(len(free)):
            if free[k] >= n:
                bin_of[i], lo_of[i] = k, bin_len - free[k]
                free[k] -= n
                break
        else:
            free.append(bin_len - n)
            bin_of[i], lo_of[i] = len(free) - 1, 0
    NB = len(free)
    # one vectorised gather / scatter for all tokens
    lo = np.asarray(lo_of, dtype=np.int64)
    first = np.asarray(bin_of, dtype=np.int64) * bin_len + lo          # flat destination row of every sequence's [CLS]
    total = int(lens.sum())
    seq_of_tok = np.repeat(np.arange(B, dtype=np.int64), lens)
    starts = np.cumsum(lens) - lens
    within = np.arange(total, dtype=np.int64) - np.repeat(starts, lens)
    src = seq_of_tok * S + within
    dst = first[seq_of_tok] + within
    ids = np.zeros(NB * bin_len, dtype=np.int64)
    tts = np.zeros(NB * bin_len, dtype=np.int64)
    pos = np.zeros(NB * bin_len, dtype=np.int64)
    ar = np.arange(bin_len, dtype=np.int32)
    seg = np.tile(ar | ((ar + 1) << 16), NB)                             # unused rows: a segment of their own
    ids[dst] = ids_np.reshape(-1)[src]
    if token_type_ids is not None:
        tts[dst] = token_type_ids.numpy().reshape(-1)[src]
    pos[dst] = within
    seg[dst] = (lo[seq_of_tok] | ((lo[seq_of_tok] + lens[seq_of_tok]) << 16)).astype(np.int32)
    t = torch.from_numpy
    return {"input_ids": t(ids).view(NB, bin_len), "token_type_ids": t(tts).view(NB, bin_len),
            "position_ids": t(pos).view(NB, bin_len), "segments": t(seg).view(NB, bin_len), "cls_index": t(first),
            "lengths": t(lens), "bins": NB}
