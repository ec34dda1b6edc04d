"""Generates the committed parity fixtures from the reference's actual arithmetic dependencies, run HERE on CPU:
HF transformers BertForSequenceClassification (eager attention, fp32) + torch DistributedDataParallel on gloo +
the restated HF AdamW (transformers 4.28.1's class is absent from the installed 5.5).

    python tests/golden/make_golden.py            # writes tests/golden/*.pt

Fixtures
  config_a_step0.pt   BASELINE config A (chinese-bert-wwm-ext, B=32, S=128, padded mask), dropout off:
                      loss, logits, per-tensor gradient norms, a 64-value sample of every gradient.
  tiny_ddp_w2.pt      tiny config, REAL torch DDP (gloo, world 2), 3 steps: per-rank loss/logits per step and the
                      post-training weights (norms + a few full tensors).
  tiny_w1.pt          tiny config, world 1, 3 steps (same content).
Inputs are regenerated from seeds by the tests (oracle.bert_ref.synthetic_batch); the fixtures carry the input ids
of step 0 and a checksum of the initial weights so a drifting RNG/initialiser is detected rather than mis-compared.
"""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import adamw_ref, bert_ref, cpu_step  # noqa: E402
from parity import full_config, tiny_config  # noqa: E402

SMALL = ["classifier.weight", "classifier.bias", "bert.pooler.dense.bias",
         "bert.encoder.layer.0.attention.output.LayerNorm.weight", "bert.encoder.layer.1.output.dense.bias",
         "bert.embeddings.LayerNorm.bias"]
STEPS = 3


def checksum(state):
    return float(sum(v.double().sum() for v in state.values()))


def summarize_weights(model):
    sd = {k: v.detach().clone() for k, v in model.named_parameters()}
    return {"norms": {k: float(v.double().norm()) for k, v in sd.items()},
            "small": {k: sd[k] for k in SMALL if k in sd}}


def tiny_batches(cfg, world):
    # step s, rank r -> seed 3000 + 10*s + r ; odd steps use the padded variant
    return [[bert_ref.synthetic_batch(cfg, 4, 128, 3000 + 10 * s + r, padded=(s % 2 == 1)) for r in range(world)]
            for s in range(STEPS)]


def _ddp_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = tiny_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = cpu_step.build_hf_model(cfg, seed=123 + rank)      # rank 0's weights must win (DDP broadcast)
    init = None
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    if rank == 0:
        init = checksum({k: v for k, v in model.named_parameters()})
    opt = cpu_step._HFOpt(model, 3e-5, 0.01)
    crit = torch.nn.CrossEntropyLoss()
    rec = {"loss": [], "logits": []}
    for s, per_rank in enumerate(tiny_batches(cfg, world)):
        b = per_rank[rank]
        out = ddp(input_ids=b["input_ids"], token_type_ids=b["token_type_ids"], attention_mask=b["attention_mask"],
                  labels=b["label"])
        loss = crit(out[1], b["label"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        rec["loss"].append(loss.detach().clone())
        rec["logits"].append(out[1].detach().clone())
    gathered = [None] * world
    dist.all_gather_object(gathered, rec)
    if rank == 0:
        torch.save({"world": world, "steps": STEPS, "init_checksum": init,
                    "loss": torch.stack([torch.stack([g["loss"][s] for g in gathered]) for s in range(STEPS)]),
                    "logits": torch.stack([torch.stack([g["logits"][s] for g in gathered]) for s in range(STEPS)]),
                    "final": summarize_weights(model),
                    "input_ids_step0_rank0": tiny_batches(cfg, world)[0][0]["input_ids"]}, out_path)
    dist.destroy_process_group()


def make_tiny_w1(out_path):
    cfg = tiny_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = cpu_step.build_hf_model(cfg, seed=123)
    init = checksum({k: v for k, v in model.named_parameters()})
    opt = cpu_step._HFOpt(model, 3e-5, 0.01)
    crit = torch.nn.CrossEntropyLoss()
    losses, logits = [], []
    for per_rank in tiny_batches(cfg, 1):
        b = per_rank[0]
        out = model(input_ids=b["input_ids"], token_type_ids=b["token_type_ids"],
                    attention_mask=b["attention_mask"], labels=b["label"])
        loss = crit(out[1], b["label"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.detach().clone())
        logits.append(out[1].detach().clone())
    torch.save({"world": 1, "steps": STEPS, "init_checksum": init, "loss": torch.stack(losses)[:, None],
                "logits": torch.stack(logits)[:, None], "final": summarize_weights(model),
                "input_ids_step0_rank0": tiny_batches(cfg, 1)[0][0]["input_ids"]}, out_path)


def make_config_a(out_path):
    cfg = full_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = cpu_step.build_hf_model(cfg, seed=123)
    state = {k: v for k, v in model.named_parameters()}
    b = bert_ref.synthetic_batch(cfg, 32, 128, 1000, padded=True)
    out = model(input_ids=b["input_ids"], token_type_ids=b["token_type_ids"], attention_mask=b["attention_mask"],
                labels=b["label"])
    loss = torch.nn.CrossEntropyLoss()(out[1], b["label"])
    loss.backward()
    torch.save({"init_checksum": checksum(state), "input_ids": b["input_ids"], "loss": float(loss),
                "hf_internal_loss": float(out[0]), "logits": out[1].detach().clone(),
                "grad_norms": {k: float(v.grad.double().norm()) for k, v in state.items()},
                "grad_samples": {k: v.grad.flatten()[:64].clone() for k, v in state.items()}}, out_path)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    make_config_a(os.path.join(HERE, "config_a_step0.pt"))
    make_tiny_w1(os.path.join(HERE, "tiny_w1.pt"))
    mp.spawn(_ddp_worker, args=(2, 29611, os.path.join(HERE, "tiny_ddp_w2.pt")), nprocs=2, join=True)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
