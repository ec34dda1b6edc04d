"""ORACLE (test infrastructure): restatement of HF ``AdamW.step`` and of the reference's parameter grouping.

The reference imports ``AdamW`` from transformers (multi-gpu-distributed-cls.py:14) and builds it at :100-111.  That
class is third-party code pinned at transformers==4.28.1 (README.md:6) and is ABSENT from the installed transformers
5.5 (``from transformers import AdamW`` raises), so it cannot be executed here.  Its published algorithm
(transformers 4.28.1, src/transformers/optimization.py::AdamW.step) is restated below, statement for statement:

    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    denom = exp_avg_sq.sqrt().add_(eps)
    step_size = lr * sqrt(1 - beta2**t) / (1 - beta1**t)        # if correct_bias
    p.addcdiv_(exp_avg, denom, value=-step_size)
    if weight_decay > 0: p.add_(p, alpha=-lr * weight_decay)

Defaults: lr as passed (3e-5 in the reference Args :252), betas (0.9, 0.999), eps 1e-6, correct_bias True.
Parity for this piece is "unpinned" in the brief's sense (no executable upstream, no upstream vectors); it is anchored
on (a) the formula above, (b) a hand-computed scalar case in tests/test_oracle.py, (c) two limits in which the
algorithm coincides with optimizers that CAN be executed here (tests/test_oracle.py::
test_hf_adamw_restatement_vs_executable_upstreams): eps = 0, no decay == torch.optim.Adam(eps=0) over several steps
(moment recursions + bias correction), and eps = 0 with decay differs from torch.optim.AdamW by exactly lr * wd * update
(decay applied after the update); only the placement of eps rests on the formula and the hand case alone.
"""
import math

import torch

NO_DECAY = ("bias", "LayerNorm.weight")  # multi-gpu-distributed-cls.py:101


def decays(name):
    return not any(nd in name for nd in NO_DECAY)


class HFAdamW:
    def __init__(self, params, lr=3e-5, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, correct_bias=True):
        """params: dict name -> fp32 tensor (updated in place)."""
        self.params = params
        self.lr, self.betas, self.eps, self.wd, self.correct_bias = lr, betas, eps, weight_decay, correct_bias
        self.state = {k: {"step": 0, "exp_avg": torch.zeros_like(v), "exp_avg_sq": torch.zeros_like(v)}
                      for k, v in params.items()}

    @torch.no_grad()
    def step(self, grads):
        b1, b2 = self.betas
        for name, p in self.params.items():
            g = grads[name]
            st = self.state[name]
            st["step"] += 1
            st["exp_avg"].mul_(b1).add_(g, alpha=(1.0 - b1))
            st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
            denom = st["exp_avg_sq"].sqrt().add_(self.eps)
            step_size = self.lr
            if self.correct_bias:
                bc1 = 1.0 - b1 ** st["step"]
                bc2 = 1.0 - b2 ** st["step"]
                step_size = step_size * math.sqrt(bc2) / bc1
            p.addcdiv_(st["exp_avg"], denom, value=-step_size)
            wd = self.wd if decays(name) else 0.0
            if wd > 0.0:
                p.add_(p, alpha=(-self.lr * wd))
