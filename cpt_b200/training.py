"""Training step of the CPT few-shot loops behind autograd.

Reference call (Oscar/oscar/fewshot/refcoco_cpt.py:243-248, gqa_cpt.py:437-462):

    loss, output = model(input_ids, segment_ids, input_mask, img_feats=img, masked_lm_labels=mlm_labels)
    loss.backward(); optimizer.step()

`mlm_loss` is a torch.autograd.Function whose forward runs cpt_train_forward_mlm and whose backward runs
cpt_train_backward_mlm (include/cpt_b200.h): autograd sees one node with the model's parameters as inputs, so
optimizers, gradient accumulation, clip_grad_norm_ and DistributedDataParallel's gradient hooks work unchanged.
The parameters stay fp32 (master weights); the engine refreshes its 16-bit GEMM copies from them before every
forward, in stream order.
"""
import torch

from .engine import GLOBAL_KEYS, layer_keys

# parameters the MLM loss does not depend on (their .grad stays None, as in the reference)
_UNUSED = ("pooler_w", "pooler_b", "nsp_w", "nsp_b")


def trainable_keys(cfg, has_img=True):
    keys = [k for f, k in GLOBAL_KEYS.items() if f not in _UNUSED and (has_img or not f.startswith("img_"))]
    for i in range(cfg.num_hidden_layers):
        keys.extend(layer_keys(i).values())
    return keys


class _MlmLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, keys, inputs, *params):
        input_ids, token_type_ids, attention_mask, position_ids, img_feats, rows, targets = inputs
        loss, saved = engine.train_forward_mlm(input_ids, token_type_ids, attention_mask, position_ids, img_feats,
                                               rows, targets)
        ctx.engine, ctx.keys, ctx.saved = engine, keys, saved
        ctx.shapes = [tuple(p.shape) for p in params]
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        eng = ctx.engine
        grads = {k: torch.zeros(s, dtype=torch.float32, device=eng.device) for k, s in zip(ctx.keys, ctx.shapes)}
        eng.train_backward_mlm(ctx.saved, grad_loss.to(torch.float32), grads)
        ctx.saved = None  # drop the tape
        out = tuple(grads[k] if ctx.needs_input_grad[3 + i] else None for i, k in enumerate(ctx.keys))
        return (None, None, None) + out


def mlm_loss(engine, named_params, input_ids, token_type_ids, attention_mask, position_ids, img_feats,
             masked_lm_labels):
    """CrossEntropyLoss(ignore_index=-1) of the MLM scores against masked_lm_labels [B,S], differentiable with
    respect to `named_params` (dict keyed like the state_dict).  Returns (loss, rows): rows = flat indices of the
    labelled positions."""
    flat = masked_lm_labels.reshape(-1)
    rows = torch.nonzero(flat != -1, as_tuple=False).squeeze(1)  # device -> host sync on its size, once per step
    if rows.numel() == 0:
        raise RuntimeError("cpt_b200: masked_lm_labels has no labelled position (the reference returns NaN here)")
    targets = flat[rows].contiguous()
    has_img = img_feats is not None and img_feats.shape[1] > 0
    keys = [k for k in trainable_keys(engine.cfg, has_img) if k in named_params]
    params = [named_params[k] for k in keys]
    inputs = (input_ids, token_type_ids, attention_mask, position_ids, img_feats, rows, targets)
    return _MlmLoss.apply(engine, keys, inputs, *params), rows
