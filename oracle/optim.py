"""Oracle (test infrastructure): the optimizers of the training step, restated on CPU tensors.

* Adan — the reference's stage-1 optimizer: ``optimizer: adanp`` (Config/multi_modal_stage1.yaml:89) is handed to timm's
  ``create_optimizer_v2`` (lhrs/optimizer/build_optimizer.py:76-86), which maps ``adanp`` to ``Adan(no_prox=False)`` and
  ``adanw`` to ``Adan(no_prox=True)``.  timm==0.9.12 (pyproject.toml:17) is NOT vendored under /root/reference and is not
  installed here, so the update rule below is restated from the published algorithm (Xie et al. 2022, "Adan: Adaptive
  Nesterov Momentum Algorithm", Alg. 1, in timm's parameterisation with betas = (0.98, 0.92, 0.99), eps = 1e-8):
  **parity unpinned** beyond that definition — there is no timm installation or reference fixture to check it against.
* AdamW — stages 2-3 (DeepSpeed "AdamW", main_pretrain_stage1.py:30-41): ``torch.optim.AdamW`` is the checker, nothing to restate.
* Gradient clipping — DeepSpeed ``gradient_clipping`` (main_pretrain_stage1.py:58,82): g *= max_norm / (norm + 1e-6) if norm > max_norm.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch


def clip_coef(grads: List[torch.Tensor], max_norm: float) -> float:
    norm = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    return max_norm / (norm + 1e-6) if (max_norm > 0 and norm > max_norm) else 1.0


class Adan:
    """Per-tensor restatement of timm.optim.Adan.step (fp32 state, in-place on ``params``)."""

    def __init__(self, params: List[torch.Tensor], lr=1e-3, betas=(0.98, 0.92, 0.99), eps=1e-8, weight_decay=0.0, no_prox=False):
        self.params, self.lr, self.betas, self.eps, self.wd, self.no_prox = params, lr, betas, eps, weight_decay, no_prox
        self.state: Dict[int, Dict[str, torch.Tensor]] = {}
        self.step_count = 0

    @torch.no_grad()
    def step(self, grads: List[torch.Tensor], lr=None, weight_decays=None):
        b1, b2, b3 = self.betas
        lr = self.lr if lr is None else lr
        self.step_count += 1
        bc1, bc2, bc3 = 1.0 - b1 ** self.step_count, 1.0 - b2 ** self.step_count, 1.0 - b3 ** self.step_count
        for i, (p, g) in enumerate(zip(self.params, grads)):
            st = self.state.setdefault(i, {})
            if not st:
                st["exp_avg"], st["exp_avg_diff"], st["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
                st["pre_grad"] = g.clone()
            diff = g - st["pre_grad"]
            st["exp_avg"].lerp_(g, 1.0 - b1)
            st["exp_avg_diff"].lerp_(diff, 1.0 - b2)
            update = g + b2 * diff
            st["exp_avg_sq"].mul_(b3).addcmul_(update, update, value=1.0 - b3)
            denom = (st["exp_avg_sq"].sqrt() / math.sqrt(bc3)).add_(self.eps)
            upd = (st["exp_avg"] / bc1 + b2 * st["exp_avg_diff"] / bc2).div_(denom)
            wd = self.wd if weight_decays is None else weight_decays[i]
            if self.no_prox:
                p.mul_(1.0 - lr * wd)
                p.add_(upd, alpha=-lr)
            else:
                p.add_(upd, alpha=-lr)
                p.div_(1.0 + lr * wd)
            st["pre_grad"].copy_(g)
