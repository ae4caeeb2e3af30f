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

# parameters each loss does not depend on (their .grad stays None, as in the reference)
_UNUSED = {"mlm": ("pooler_w", "pooler_b", "nsp_w", "nsp_b"),
           "nsp": ("mlm_dense_w", "mlm_dense_b", "mlm_ln_g", "mlm_ln_b", "mlm_bias")}


_HEAD_FIELDS = ("mlm_dense_w", "mlm_dense_b", "mlm_ln_g", "mlm_ln_b", "mlm_bias", "pooler_w", "pooler_b", "nsp_w",
                "nsp_b")


def trainable_groups(cfg, head, has_img=True):
    """Parameter keys in the order the backward COMPLETES their gradients (include/cpt_b200.h, progress callback):
    [loss head], [layer L-1], ..., [layer 0], [embeddings + region embedding].  The gradient slab is laid out in this
    order so that each group is one contiguous range a data-parallel run can all-reduce as soon as it is final."""
    ok = [f for f in GLOBAL_KEYS if f not in _UNUSED[head] and (has_img or not f.startswith("img_"))]
    groups = [[GLOBAL_KEYS[f] for f in ok if f in _HEAD_FIELDS]]
    for i in reversed(range(cfg.num_hidden_layers)):
        lk = layer_keys(i)
        # query / key / value weight gradients adjacent: the backward then writes them as ONE [3H, H] product
        first = [lk["q_w"], lk["k_w"], lk["v_w"]]
        groups.append(first + [k for k in lk.values() if k not in first])
    groups.append([GLOBAL_KEYS[f] for f in ok if f not in _HEAD_FIELDS])
    return groups


def trainable_keys(cfg, head, has_img=True):
    return [k for g in trainable_groups(cfg, head, has_img) for k in g]


class _Loss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, keys, inputs, *params):
        head, input_ids, token_type_ids, attention_mask, position_ids, img_feats, rows, targets, dropout = inputs
        loss, saved = engine.train_forward(head, input_ids, token_type_ids, attention_mask, position_ids, img_feats,
                                           rows, targets, dropout)
        ctx.engine, ctx.keys, ctx.saved = engine, keys, saved
        saved["groups"] = trainable_groups(engine.cfg, head, img_feats is not None and img_feats.shape[1] > 0)
        ctx.shapes = [tuple(p.shape) for p in params]
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        eng = ctx.engine
        if ctx.saved is None:
            raise RuntimeError("cpt_b200: backward() was called a second time on the same training step; the tape of a "
                               "step is released by its first backward (retain_graph=True is not supported)")
        slot = getattr(eng, "owner_slot", None)
        if slot is not None:  # an optimizer step may follow: see _EngineSlot.dirty_*
            slot.dirty_train = slot.dirty_infer = True
        # one zero-filled slab for every gradient (one memset instead of ~200 fill launches); 256-byte aligned views
        grads = eng.grad_buffers(ctx.saved, ctx.keys, ctx.shapes)
        grads = eng.train_backward(ctx.saved, grad_loss.to(torch.float32), grads)
        ctx.saved = None  # drop the tape
        out = tuple(grads[k] if ctx.needs_input_grad[3 + i] else None for i, k in enumerate(ctx.keys))
        return (None, None, None) + out


def _check_targets(targets, n_classes, what):
    """Labels other than ignore_index must be class indices: torch's CrossEntropyLoss raises for anything else (a
    device-side assert on CUDA); the native kernels index the logits row with them, so check before launching.  The
    stream was already synchronised by the nonzero() that produced `targets`."""
    lo, hi = torch.aminmax(targets)
    lo, hi = int(lo), int(hi)
    if lo < 0 or hi >= n_classes:
        raise RuntimeError("cpt_b200: %s holds a target outside [0, %d) other than ignore_index -1 (min %d, max %d)"
                           % (what, n_classes, lo, hi))


def draw_dropout(cfg, training):
    """(p_hidden, p_attn, seed) for one forward, or None.  The seed comes from torch's default CPU generator, so
    torch.manual_seed() makes a run reproducible (the masks themselves are this library's, include/cpt_b200.h)."""
    p_h, p_a = float(cfg.hidden_dropout_prob), float(cfg.attention_probs_dropout_prob)
    if not training or (p_h <= 0 and p_a <= 0):
        return None
    return (p_h, p_a, int(torch.randint(0, 2 ** 62, (1,)).item()))


def mlm_loss(engine, named_params, input_ids, token_type_ids, attention_mask, position_ids, img_feats,
             masked_lm_labels, dropout=None):
    """CrossEntropyLoss(ignore_index=-1) of the MLM scores against masked_lm_labels [B,S], differentiable with
    respect to `named_params` (dict keyed like the state_dict).  Returns (loss, rows): rows = flat indices of the
    labelled positions."""
    B, T = input_ids.shape
    S = T + (0 if img_feats is None else img_feats.shape[1])
    if tuple(masked_lm_labels.shape) != (B, S):
        # the reference's CrossEntropyLoss raises a batch-size mismatch here (modeling_rec.py:146-149); the native
        # kernels index rows of the [B*S, H] stream with these positions
        raise ValueError("cpt_b200: masked_lm_labels must have shape [B, T+R] = %s, got %s"
                         % ((B, S), tuple(masked_lm_labels.shape)))
    flat = masked_lm_labels.reshape(-1)
    rows = torch.nonzero(flat != -1, as_tuple=False).squeeze(1)  # device -> host sync on its size, once per step
    if rows.numel() == 0:
        raise RuntimeError("cpt_b200: masked_lm_labels has no labelled position (the reference returns NaN here)")
    targets = flat[rows].contiguous()
    _check_targets(targets, engine.cfg.vocab_size, "masked_lm_labels")
    has_img = img_feats is not None and img_feats.shape[1] > 0
    keys = [k for k in trainable_keys(engine.cfg, "mlm", has_img) if k in named_params]
    params = [named_params[k] for k in keys]
    inputs = ("mlm", input_ids, token_type_ids, attention_mask, position_ids, img_feats, rows, targets, dropout)
    return _Loss.apply(engine, keys, inputs, *params), rows


def nsp_loss(engine, named_params, input_ids, token_type_ids, attention_mask, position_ids, img_feats,
             next_sentence_label, dropout=None):
    """CrossEntropyLoss(ignore_index=-1) of cls.seq_relationship(pooled) against next_sentence_label [B]
    (modeling_vcr.py:120-127), differentiable with respect to `named_params`."""
    flat = next_sentence_label.reshape(-1)
    if flat.numel() != input_ids.shape[0]:
        raise ValueError("cpt_b200: next_sentence_label must hold one label per sample (%d), got %d"
                         % (input_ids.shape[0], flat.numel()))
    S = input_ids.shape[1] + (0 if img_feats is None else img_feats.shape[1])
    keep = torch.nonzero(flat != -1, as_tuple=False).squeeze(1)
    if keep.numel() == 0:
        raise RuntimeError("cpt_b200: next_sentence_label has no labelled sample (the reference returns NaN here)")
    targets = flat[keep].contiguous()
    _check_targets(targets, int(getattr(engine.cfg, "num_contrast_classes", 2)), "next_sentence_label")
    rows = (keep * S).contiguous()  # the [CLS] row of every labelled sample
    has_img = img_feats is not None and img_feats.shape[1] > 0
    keys = [k for k in trainable_keys(engine.cfg, "nsp", has_img) if k in named_params]
    params = [named_params[k] for k in keys]
    inputs = ("nsp", input_ids, token_type_ids, attention_mask, position_ids, img_feats, rows, targets, dropout)
    return _Loss.apply(engine, keys, inputs, *params), rows
