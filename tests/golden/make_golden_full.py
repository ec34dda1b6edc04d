"""Full-size multi-rank parity fixtures for BASELINE.json's configs (A: chinese-bert-wwm-ext B=32 S=128,
B: bert-base B=16 S=512, C: bert-large B=16 S=128), generated HERE on CPU from the reference's arithmetic
dependencies: HF transformers BertForSequenceClassification (eager attention, fp32, dropout off) for every rank's
forward/backward, the DDP mean of the per-rank gradients (default_hooks.py:18-33 semantics; the restatement
oracle/ddp_ref.mean_grads is pinned against REAL torch DDP on gloo by tiny_ddp_w2.pt / tests/test_oracle.py), and the
restated HF AdamW.

    python tests/golden/make_golden_full.py [A] [B] [C]      # writes tests/golden/config_{a,b,c}_ddp.pt

One process plays all ranks in turn (the eight-rank BERT-large job does not fit eight CPU processes).  Step 0 is
shared by all world sizes (same initial weights, rank r's batch does not depend on the world size).

Per config and world size the fixture holds, for `steps` optimizer steps: the loss of every rank at every step, the
logits of every rank at every step, and after the last step 64 evenly strided values + the norm of every fp32 weight
tensor.  For rank 0 at step 0 it also holds the norm and a 64-value strided sample of every gradient tensor.  Inputs
are regenerated from seeds (`batch_seed`), the initial weights from `set_seed(123)` + the package's CPU initialiser;
`init_checksum` and rank 0's step-0 input ids guard against a drifting initialiser / RNG.  bench.py checks its N-GPU run against these at every N (the `parity` block of its JSON line),
tests/test_ddp.py and tests/test_model.py use them on the GPU, tests/test_oracle.py pins the oracle to them on CPU.
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import adamw_ref, bert_ref, cpu_step, ddp_ref  # noqa: E402
import pytorch_distributed_nlp_b200 as b2  # noqa: E402

CONFIGS = {
    # name: (config factory, per-GPU batch, seq, optimizer steps, world sizes)
    "A": (b2.chinese_bert_wwm_ext_config, 32, 128, 3, (1, 2, 4, 8)),
    "B": (b2.bert_base_config, 16, 512, 2, (1, 2, 8)),
    "C": (b2.bert_large_config, 16, 128, 2, (1, 2, 8)),
}
NSAMPLE = 64


def batch_seed(step, rank):
    return 5000 + 100 * step + rank


def make_batch(cfg, B, S, step, rank):
    # odd steps use the padded variant (per-row valid length ~ U{8..S}), even steps full-length rows
    return bert_ref.synthetic_batch(cfg, B, S, batch_seed(step, rank), padded=(step % 2 == 1))


def strided_sample(t, n=NSAMPLE):
    f = t.detach().flatten()
    if f.numel() <= n:
        return f.clone()
    idx = torch.arange(n, dtype=torch.int64) * (f.numel() - 1) // (n - 1)      # exact integer stride
    return f[idx].clone()


def hf_loss_and_grads(model, batch):
    model.zero_grad(set_to_none=True)
    out = model(input_ids=batch["input_ids"], token_type_ids=batch["token_type_ids"],
                attention_mask=batch["attention_mask"], labels=batch["label"])
    loss = torch.nn.CrossEntropyLoss()(out[1], batch["label"])
    loss.backward()
    grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v))
             for k, v in model.named_parameters()}
    return loss.detach().clone(), out[1].detach().clone(), grads


def generate(name):
    factory, B, S, steps, worlds = CONFIGS[name]
    cfg = factory(num_labels=6, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    # Initial weights: the PACKAGE's own initialiser under set_seed(123) (HF `_init_weights` semantics -- N(0, 0.02)
    # linears / embeddings, zero pad row and biases, LayerNorm (1, 0) -- drawn in the package's order, CPU generator),
    # loaded into the HF model.  bench.py can so rebuild the very same weights without touching oracle/ or transformers.
    b2.set_seed(123)
    init = {k: v.detach().clone() for k, v in b2.BertForSequenceClassification(cfg).named_parameters()}
    model = cpu_step.build_hf_model(cfg, seed=123)
    missing = model.load_state_dict(init, strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k or "token_type_ids" in k
                                               for k in missing.missing_keys), missing
    model.eval()          # dropout is 0 anyway; eval() keeps HF from drawing RNG
    init_checksum = float(sum(v.double().sum() for v in init.values()))
    t0 = time.time()
    # step 0 for ranks 0..max(world)-1 (identical across world sizes)
    step0 = [hf_loss_and_grads(model, make_batch(cfg, B, S, 0, r)) for r in range(max(worlds))]
    print("config %s: step 0 for %d ranks in %.0f s" % (name, max(worlds), time.time() - t0), flush=True)
    out = {"config": name, "batch": B, "seq": S, "steps": steps, "init_checksum": init_checksum,
           "input_ids_step0_rank0": make_batch(cfg, B, S, 0, 0)["input_ids"],
           "step0_rank0": {"grad_norms": {k: float(g.double().norm()) for k, g in step0[0][2].items()},
                           "grad_samples": {k: strided_sample(g) for k, g in step0[0][2].items()}},
           "worlds": {}}
    for w in worlds:
        params = {k: v.clone() for k, v in init.items()}
        opt = adamw_ref.HFAdamW(params, lr=3e-5, weight_decay=0.01)
        losses, logits = [], []
        for s in range(steps):
            if s == 0:
                res = step0[:w]
            else:
                with torch.no_grad():
                    for k, p in model.named_parameters():
                        p.copy_(params[k])
                res = [hf_loss_and_grads(model, make_batch(cfg, B, S, s, r)) for r in range(w)]
            losses.append(torch.stack([r[0] for r in res]))
            logits.append(torch.stack([r[1] for r in res]))
            opt.step(ddp_ref.mean_grads([r[2] for r in res]))
        out["worlds"][w] = {"loss": torch.stack(losses), "logits": torch.stack(logits),
                            "final_norms": {k: float(v.double().norm()) for k, v in params.items()},
                            "final_samples": {k: strided_sample(v) for k, v in params.items()}}
        print("config %s world %d: losses %s  (%.0f s)" % (name, w, [round(float(x), 5) for x in
                                                                      torch.stack(losses).mean(1)], time.time() - t0),
              flush=True)
    path = os.path.join(HERE, "config_%s_ddp.pt" % name.lower())
    torch.save(out, path)
    print(path, os.path.getsize(path), "bytes", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for nm in (sys.argv[1:] or ["A", "B", "C"]):
        generate(nm.upper())
