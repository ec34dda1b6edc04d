"""Per-tensor gradient error of the full config-A step against the fp32 oracle (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from parity import bert_ref, full_config, make_model, state_from_hf_init, to_dev

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device("cuda", 0)
cfg = full_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, num_hidden_layers=layers)
state = state_from_hf_init(cfg)
model = make_model(cfg, state, dev).train()
batch = bert_ref.synthetic_batch(cfg, 32, 128, 1000, padded=True)
d = to_dev(batch, dev)
out = model(input_ids=d["input_ids"], token_type_ids=d["token_type_ids"], attention_mask=d["attention_mask"], labels=d["label"])
loss = F.cross_entropy(out[1], d["label"]); loss.backward(); torch.cuda.synchronize()
torch.set_num_threads(os.cpu_count() or 1)
rl, rz, rg = bert_ref.loss_and_grads(state, cfg, batch)
print("loss", float(loss), float(rl), "max dlogit", float((out[1].detach().cpu() - rz).abs().max()))
g = model.grad_dict()
for k, r in rg.items():
    a = g[k].cpu().double().flatten(); r = r.double().flatten()
    rn = float(r.norm())
    if rn < 1e-9:
        continue
    rel = float((a - r).norm() / rn)
    ratio = float(a.norm() / rn)
    cos = float((a @ r) / (a.norm() * r.norm() + 1e-30))
    proj = float((a @ r) / (r @ r))     # least-squares gain
    print("%-62s rel %.4f normratio %.4f gain %.4f cos %.5f" % (k, rel, ratio, proj, cos))
