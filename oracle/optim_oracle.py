"""CPU restatement of the optimizers on the CPT few-shot path — TEST INFRASTRUCTURE ONLY (imported by tests/).

adamw_hf1   : `AdamW.step` of pytorch-transformers 1.x (huggingface/transformers @ 067923d3, the commit pinned by
              /root/reference/install.sh:29-32, file pytorch_transformers/optimization.py).  That dependency is NOT
              vendored under /root/reference, so its published algorithm is restated here:
                  exp_avg    = b1 * exp_avg    + (1 - b1) * grad
                  exp_avg_sq = b2 * exp_avg_sq + (1 - b2) * grad^2
                  denom      = sqrt(exp_avg_sq) + eps
                  step_size  = lr * sqrt(1 - b2^t) / (1 - b1^t)      (lr when correct_bias=False)
                  p         -= step_size * exp_avg / denom
                  p         -= lr * weight_decay * p                  (after the Adam update, if weight_decay > 0)
              Call sites: Oscar/oscar/fewshot/gqa_cpt.py:342, vcr_nsp_cpt.py.  PARITY UNPINNED for this function (no
              copy of the original in the container); tests/test_oracle_optim.py pins the torch-semantics twin below
              against torch.optim.AdamW and checks the two restatements agree where the algorithms coincide
              (weight_decay = 0, eps -> 0).
adamw_torch : torch.optim.AdamW's update (Oscar/oscar/fewshot/refcoco_cpt.py:342), pinned against torch itself.
"""
import math

import torch


def adamw_hf1(p, grad, state, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
    b1, b2 = betas
    state["step"] = state.get("step", 0) + 1
    m = state.setdefault("exp_avg", torch.zeros_like(p))
    v = state.setdefault("exp_avg_sq", torch.zeros_like(p))
    m.mul_(b1).add_(grad, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1.0 - b2)
    denom = v.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        t = state["step"]
        step_size = step_size * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
    return p


def adamw_torch(p, grad, state, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
    b1, b2 = betas
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    m = state.setdefault("exp_avg", torch.zeros_like(p))
    v = state.setdefault("exp_avg_sq", torch.zeros_like(p))
    p.mul_(1.0 - lr * weight_decay)
    m.mul_(b1).add_(grad, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1.0 - b2)
    denom = (v.sqrt() / math.sqrt(1.0 - b2 ** t)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / (1.0 - b1 ** t))
    return p
