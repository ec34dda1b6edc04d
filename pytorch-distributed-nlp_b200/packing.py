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
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    mask = attention_mask.to(torch.int64)
    lengths = mask.sum(1)
    # valid tokens must form a prefix (right padding): cumulative product of the mask == the mask
    if not torch.equal(torch.cumprod(mask, 1), mask):
        raise ValueError("pack_batch: attention_mask is not a right-padded prefix mask")
    if int(lengths.min()) < 1:
        raise ValueError("pack_batch: empty sequence (no valid token)")
    if int(lengths.max()) > bin_len:
        raise ValueError("pack_batch: a sequence has %d valid tokens, more than the %d-token bin"
                         % (int(lengths.max()), bin_len))
    order = sorted(range(B), key=lambda i: (-int(lengths[i]), i))     # longest first; ties in batch order
    free, where = [], [None] * B                                         # free[k]: tokens left in bin k
    for i in order:
        n = int(lengths[i])
        for k in range(len(free)):
            if free[k] >= n:
                where[i] = (k, bin_len - free[k])
                free[k] -= n
                break
        else:
            free.append(bin_len - n)
            where[i] = (len(free) - 1, 0)
    NB = len(free)
    ids = torch.zeros(NB, bin_len, dtype=torch.int64)
    tts = torch.zeros(NB, bin_len, dtype=torch.int64)
    pos = torch.zeros(NB, bin_len, dtype=torch.int64)
    ar = torch.arange(bin_len, dtype=torch.int32)
    seg = (ar | ((ar + 1) << 16)).repeat(NB, 1).contiguous()             # unused rows: a segment of their own
    cls_index = torch.zeros(B, dtype=torch.int64)
    for i in range(B):
        k, lo = where[i]
        n = int(lengths[i])
        ids[k, lo:lo + n] = input_ids[i, :n]
        tts[k, lo:lo + n] = token_type_ids[i, :n]
        pos[k, lo:lo + n] = torch.arange(n)
        seg[k, lo:lo + n] = lo | ((lo + n) << 16)
        cls_index[i] = k * bin_len + lo
    return {"input_ids": ids, "token_type_ids": tts, "position_ids": pos, "segments": seg, "cls_index": cls_index,
            "lengths": lengths, "bins": NB}
