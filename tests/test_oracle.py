"""CPU: the oracle is pinned before it is trusted — against the HF model it restates, against the committed fixtures
(generated from HF + real torch DDP on gloo), and, for the AdamW whose upstream class is not installable here,
against a hand-worked case of the published formula."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parity import adamw_ref, bert_ref, full_config, tiny_config
from oracle import cpu_step, ddp_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _hf_state(cfg, seed=123):
    hf = cpu_step.build_hf_model(cfg, seed=seed)
    return hf, {k: v.detach().clone() for k, v in hf.named_parameters()}


# the oracle is pinned on the shapes every GPU parity test uses: the tiny config, BASELINE config B's sequence length
# (512, padded tail blocks) and config C's widths (bert-large: H 1024, 16 heads, I 4096), at reduced depth
_SHAPES = {
    "tiny": (dict(), 4, 128),
    "config-B-shape": (dict(vocab_size=2000, hidden_size=768, num_hidden_layers=1, num_attention_heads=12,
                            intermediate_size=3072, max_position_embeddings=512), 2, 512),
    "config-C-shape": (dict(vocab_size=2000, hidden_size=1024, num_hidden_layers=1, num_attention_heads=16,
                            intermediate_size=4096, max_position_embeddings=512), 2, 128),
}


@pytest.mark.parametrize("padded", [False, True])
@pytest.mark.parametrize("shape", sorted(_SHAPES))
def test_oracle_equals_hf_forward_backward(padded, shape):
    kw, batch, seq = _SHAPES[shape]
    cfg = tiny_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw)
    hf, state = _hf_state(cfg)
    b = bert_ref.synthetic_batch(cfg, batch, seq, 1000, padded=padded)
    out = hf(input_ids=b["input_ids"], token_type_ids=b["token_type_ids"], attention_mask=b["attention_mask"],
             labels=b["label"])
    out[0].backward()
    loss, logits, grads = bert_ref.loss_and_grads(state, cfg, b)
    assert abs(float(out[0].detach()) - float(loss)) < 2e-6
    assert float((out[1].detach() - logits).abs().max()) < 2e-6
    for k, p in hf.named_parameters():
        assert float((grads[k] - p.grad).abs().max()) < 2e-6 + 1e-5 * float(p.grad.abs().max()), k
    # reference comment at multi-gpu-distributed-cls.py:168: criterion(logits, label) == output[0]
    assert abs(float(torch.nn.functional.cross_entropy(logits, b["label"])) - float(loss)) < 1e-7


def test_oracle_dropout_masks_are_what_hf_dropout_does():
    """With an explicit keep-mask the oracle reproduces F.dropout's arithmetic (x * mask / (1-p))."""
    x = torch.randn(4, 8)
    mask = torch.rand(4, 8) > 0.1
    assert torch.equal(bert_ref._drop(x, 0.1, {"k": mask}, "k"), x * mask / 0.9)


@pytest.mark.parametrize("name", ["tiny_w1.pt", "tiny_ddp_w2.pt"])
def test_oracle_trajectory_matches_fixture(name):
    """bert_ref + adamw_ref + ddp_ref reproduce 3 steps of HF (+ real torch DDP, world 2) recorded in the fixture."""
    gold = torch.load(os.path.join(GOLD, name))
    world = gold["world"]
    cfg = tiny_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    _, state = _hf_state(cfg)
    chk = float(sum(v.double().sum() for v in state.values()))
    assert abs(chk - gold["init_checksum"]) < 1e-6 * max(1.0, abs(chk))
    batches = [[bert_ref.synthetic_batch(cfg, 4, 128, 3000 + 10 * s + r, padded=(s % 2 == 1)) for r in range(world)]
               for s in range(gold["steps"])]
    assert torch.equal(batches[0][0]["input_ids"], gold["input_ids_step0_rank0"])
    hist = ddp_ref.train(state, cfg, batches)
    for s, h in enumerate(hist):
        assert float((h["loss_per_rank"] - gold["loss"][s]).abs().max()) < 2e-6
        for r in range(world):
            assert float((h["logits_per_rank"][r] - gold["logits"][s][r]).abs().max()) < 5e-6
    for k, n in gold["final"]["norms"].items():
        assert abs(float(state[k].double().norm()) - n) < 1e-5 * max(1.0, n), k
    for k, v in gold["final"]["small"].items():
        assert float((state[k] - v).abs().max()) < 2e-6, k


def test_config_a_fixture_matches_oracle():
    gold = torch.load(os.path.join(GOLD, "config_a_step0.pt"))
    cfg = full_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    _, state = _hf_state(cfg)
    b = bert_ref.synthetic_batch(cfg, 32, 128, 1000, padded=True)
    assert torch.equal(b["input_ids"], gold["input_ids"])
    torch.set_num_threads(os.cpu_count() or 1)
    loss, logits, grads = bert_ref.loss_and_grads(state, cfg, b)
    assert abs(float(loss) - gold["loss"]) < 5e-6 and abs(gold["loss"] - gold["hf_internal_loss"]) < 1e-6
    assert float((logits - gold["logits"]).abs().max()) < 1e-5
    assert abs(gold["loss"] - math.log(6)) < 0.1          # fresh 6-class head: loss ~ ln 6
    scale = max(gold["grad_norms"].values())
    for k, n in gold["grad_norms"].items():   # key.bias gradients are analytically zero (softmax shift invariance)
        assert abs(float(grads[k].double().norm()) - n) < 1e-4 * max(n, 1e-4 * scale), k


def test_config_a_ddp_fixture_matches_oracle_world2():
    """The full-size multi-rank fixture bench.py's `parity` block is judged against (tests/golden/config_a_ddp.pt: HF
    fp32 per rank + DDP mean + HF AdamW, package initialiser under set_seed(123)) vs the oracle restatement: world 2,
    step 0 (both ranks) and, through ddp_ref.mean_grads + adamw_ref, the rank-0 loss of step 1."""
    import pytorch_distributed_nlp_b200 as b2
    path = os.path.join(GOLD, "config_a_ddp.pt")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    fx = torch.load(path)
    cfg = full_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    b2.set_seed(123)
    state = {k: v.detach().clone() for k, v in b2.BertForSequenceClassification(cfg).named_parameters()}
    chk = float(sum(v.double().sum() for v in state.values()))
    assert abs(chk - fx["init_checksum"]) <= 1e-6 * max(1.0, abs(fx["init_checksum"]))
    torch.set_num_threads(os.cpu_count() or 1)
    B, S = fx["batch"], fx["seq"]
    mk = lambda s, r: bert_ref.synthetic_batch(cfg, B, S, 5000 + 100 * s + r, padded=(s % 2 == 1))
    assert torch.equal(mk(0, 0)["input_ids"], fx["input_ids_step0_rank0"])
    w2 = fx["worlds"][2]
    res = [bert_ref.loss_and_grads(state, cfg, mk(0, r)) for r in range(2)]
    for r in range(2):
        assert abs(float(res[r][0]) - float(w2["loss"][0][r])) < 5e-6
        assert float((res[r][1] - w2["logits"][0][r]).abs().max()) < 1e-5
    scale = max(fx["step0_rank0"]["grad_norms"].values())
    for k, n in fx["step0_rank0"]["grad_norms"].items():
        assert abs(float(res[0][2][k].double().norm()) - n) < 1e-4 * max(n, 1e-4 * scale), k
    opt = adamw_ref.HFAdamW(state, lr=3e-5, weight_decay=0.01)
    opt.step(ddp_ref.mean_grads([res[0][2], res[1][2]]))
    l1, z1, _ = bert_ref.loss_and_grads(state, cfg, mk(1, 0))
    assert abs(float(l1) - float(w2["loss"][1][0])) < 2e-5
    assert float((z1 - w2["logits"][1][0]).abs().max()) < 5e-5


def test_hf_adamw_restatement_hand_case():
    """One scalar, two steps, worked by hand from transformers 4.28.1 optimization.py::AdamW.step."""
    p = {"w.weight": torch.tensor([1.0]), "w.bias": torch.tensor([1.0])}
    opt = adamw_ref.HFAdamW(p, lr=0.1, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01)
    g = {"w.weight": torch.tensor([0.5]), "w.bias": torch.tensor([0.5])}
    opt.step(g)
    m, v = 0.05, 0.00025
    step_size = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    w = 1.0 - step_size * m / (math.sqrt(v) + 1e-6)
    w_decay = w - 0.1 * 0.01 * w                       # decay AFTER the Adam update, on the updated weight
    assert abs(float(p["w.weight"]) - w_decay) < 1e-6
    assert abs(float(p["w.bias"]) - w) < 1e-6          # 'bias' is in the reference's no_decay list (:101)
    opt.step(g)
    m2, v2 = 0.9 * m + 0.05, 0.999 * v + 0.00025
    ss2 = 0.1 * math.sqrt(1 - 0.999 ** 2) / (1 - 0.9 ** 2)
    w2 = w - ss2 * m2 / (math.sqrt(v2) + 1e-6)
    assert abs(float(p["w.bias"]) - w2) < 1e-6
    assert adamw_ref.decays("bert.encoder.layer.0.output.dense.weight")
    assert not adamw_ref.decays("bert.encoder.layer.0.output.LayerNorm.weight")
    assert not adamw_ref.decays("classifier.bias")


def _gloo_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = tiny_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, num_hidden_layers=1)
    hf = cpu_step.build_hf_model(cfg, seed=123 + rank)
    ddp = torch.nn.parallel.DistributedDataParallel(hf)
    b = bert_ref.synthetic_batch(cfg, 2, 128, 50 + rank)
    out = ddp(input_ids=b["input_ids"], token_type_ids=b["token_type_ids"], attention_mask=b["attention_mask"],
              labels=b["label"])
    torch.nn.functional.cross_entropy(out[1], b["label"]).backward()
    if rank == 0:
        torch.save({k: v.grad.clone() for k, v in hf.named_parameters()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_hf_adamw_restatement_vs_executable_upstreams():
    """HF's AdamW class cannot be executed here (removed from the installed transformers), but two limits of its
    published algorithm coincide with optimizers that CAN: (1) with eps = 0 and weight_decay = 0 it is exactly
    torch.optim.Adam with eps = 0 (lr * sqrt(bc2) / bc1 * m / sqrt(v) either way): pins the moment recursions and the
    bias correction over several steps; (2) its weight decay is torch.optim.AdamW's decoupled decay applied AFTER the Adam
    update instead of before it, so with eps = 0 the two differ by exactly lr * wd * (the Adam update): pins the decay
    ordering.  What stays anchored on the formula + hand case alone is the placement of eps."""
    torch.manual_seed(0)
    p0 = torch.randn(64, dtype=torch.float64)
    grads = [torch.randn(64, dtype=torch.float64) * 0.1 for _ in range(6)]
    # (1) vs torch.optim.Adam
    tp = torch.nn.Parameter(p0.clone())
    to = torch.optim.Adam([tp], lr=3e-3, betas=(0.9, 0.999), eps=0.0)
    hp = {"w.weight": p0.clone()}
    ho = adamw_ref.HFAdamW(hp, lr=3e-3, eps=0.0, weight_decay=0.0)
    for g in grads:
        tp.grad = g.clone()
        to.step()
        ho.step({"w.weight": g})
        assert float((tp.data - hp["w.weight"]).abs().max()) < 1e-14
    # (2) vs torch.optim.AdamW: one step from identical state, eps = 0
    lr, wd = 3e-3, 0.1
    tp = torch.nn.Parameter(p0.clone())
    to = torch.optim.AdamW([tp], lr=lr, betas=(0.9, 0.999), eps=0.0, weight_decay=wd)
    hp = {"w.weight": p0.clone()}
    ho = adamw_ref.HFAdamW(hp, lr=lr, eps=0.0, weight_decay=wd)
    tp.grad = grads[0].clone()
    to.step()
    ho.step({"w.weight": grads[0]})
    u = lr * torch.sign(grads[0])            # first Adam update with eps = 0: lr * m_hat / sqrt(v_hat) = lr * sign(g)
    # torch: p0 (1 - lr wd) - u ;  HF: (p0 - u) (1 - lr wd)  ->  HF - torch = lr wd u
    assert float(((hp["w.weight"] - tp.data) - lr * wd * u).abs().max()) < 1e-14


def test_ddp_mean_semantics_on_gloo_world2(tmp_path):
    """ddp_ref.mean_grads == what real torch DDP leaves in .grad (gloo, world 2, rank-0 weights broadcast)."""
    ctx = mp.get_context("spawn")
    out_path = str(tmp_path / "ddp_grads.pt")
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, 29633, out_path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = torch.load(out_path)
    cfg = tiny_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, num_hidden_layers=1)
    _, state = _hf_state(cfg, seed=123)
    per_rank = [bert_ref.loss_and_grads(state, cfg, bert_ref.synthetic_batch(cfg, 2, 128, 50 + r))[2] for r in range(2)]
    avg = ddp_ref.mean_grads(per_rank)
    for k in avg:
        assert float((avg[k] - got[k]).abs().max()) < 2e-6, k
