"""CPU fp32 oracle for the CPT cross-modal BERT hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file; the product (`cpt_b200/`) never does and fails loudly when the
CUDA library is missing.

It is a functional, state-dict-driven restatement (plain torch fp32 ops, no nn.Modules) of

  * the reference's own code:
      BertImgModel.forward            /root/reference/Oscar/oscar/modeling/modeling_bert.py:199-279
      CaptionBertSelfAttention.forward                                   modeling_bert.py:30-70
      CaptionBertAttention / Layer / Encoder wiring                      modeling_bert.py:82-87,100-126,139-147
      REC_MLM_CPT.forward             /root/reference/Oscar/oscar/modeling/modeling_rec.py:137-152
      NSPCPT.forward                  /root/reference/Oscar/oscar/modeling/modeling_vcr.py:115-129
      caller-side gathers             Oscar/oscar/zeroshot/refcoco_cpt.py:217-219,234-242
                                      Oscar/oscar/fewshot/refcoco_cpt.py:283-291
                                      Oscar/oscar/fewshot/gqa_cpt.py:597-601
                                      Oscar/oscar/fewshot/vcr_nsp_cpt.py:597-604
  * the un-vendored dependency the reference calls into: huggingface/transformers (then
    "pytorch-transformers" 1.x) at commit 067923d3267325f525f4e46f357360c191ba562e, pinned by
    /root/reference/install.sh:29-32.  Its published algorithm for BertEmbeddings,
    BertSelfOutput, BertIntermediate, BertOutput, BertPooler, BertLayerNorm,
    BertPredictionHeadTransform and BertLMPredictionHead is restated below; the reference
    call sites are modeling_bert.py:85,144-145,244-245,263,275 and modeling_rec.py:105,143.

Pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF RUN IN THE BUILD CONTAINER:
tests/golden/make_golden.py imports the reference's unmodified oscar/modeling files
(via oracle/ref_shim.py) and stores their outputs under tests/golden/*.pt;
tests/test_oracle_golden.py checks this file against them (and, where transformers 5.x is
importable, against HF's eager BERT blocks as an independent second opinion).
"""
import math

import torch
import torch.nn.functional as F


def _ln(x, w, b, eps):
    # BertLayerNorm: biased variance, eps inside the sqrt (== torch.nn.LayerNorm)
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return (x - u) / torch.sqrt(s + eps) * w + b


def _gelu(x):
    # hidden_act == "gelu": exact erf form
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _linear(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


# Optional mask source for training-mode parity tests: callable(site, x) -> {0,1} tensor shaped like x, with
# site = ("emb_text"|"emb_img", None) or ("attn"|"ao"|"down", layer).  None -> torch's own dropout (the reference's
# behaviour).  tests/test_gpu_train.py sets it to a restatement of the CUDA path's counter-based mask so that both
# sides drop the same elements.
MASK_PROVIDER = None


def _dropout(x, p, training, site=None):
    if not (training and p > 0):
        return x
    if MASK_PROVIDER is not None:
        return x * MASK_PROVIDER(site, x).to(x.dtype) / (1.0 - p)
    return F.dropout(x, p, training)


def extended_attention_mask(attention_mask, dtype=torch.float32):
    """modeling_bert.py:213-226 : additive mask, -10000 (not -inf) on masked keys."""
    if attention_mask.dim() == 2:
        ext = attention_mask[:, None, None, :]
    elif attention_mask.dim() == 3:
        ext = attention_mask[:, None, :, :]
    else:
        raise NotImplementedError
    return (1.0 - ext.to(dtype)) * -10000.0


def text_embeddings(sd, cfg, input_ids, token_type_ids=None, position_ids=None, training=False):
    """BertEmbeddings.forward (pytorch-transformers 1.x); call site modeling_bert.py:244-245."""
    T = input_ids.size(1)
    if position_ids is None:
        position_ids = torch.arange(T, dtype=torch.long, device=input_ids.device)[None].expand_as(input_ids)
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    p = "bert.embeddings."
    e = (F.embedding(input_ids, sd[p + "word_embeddings.weight"], padding_idx=0)
         + F.embedding(position_ids, sd[p + "position_embeddings.weight"])
         + F.embedding(token_type_ids, sd[p + "token_type_embeddings.weight"]))
    e = _ln(e, sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"], cfg.layer_norm_eps)
    return _dropout(e, cfg.hidden_dropout_prob, training, ("emb_text", None))


def region_embeddings(sd, cfg, img_feats, training=False):
    """modeling_bert.py:261-266 : Linear(F->H) [+ LayerNorm iff use_img_layernorm] + dropout."""
    x = _linear(img_feats, sd, "bert.img_embedding")
    if getattr(cfg, "use_img_layernorm", 0):
        x = _ln(x, sd["bert.LayerNorm.weight"], sd["bert.LayerNorm.bias"], cfg.img_layer_norm_eps)
    return _dropout(x, cfg.hidden_dropout_prob, training, ("emb_img", None))


def self_attention(sd, cfg, prefix, h, ext_mask, training=False, layer=None):
    """CaptionBertSelfAttention.forward, modeling_bert.py:38-67 (history_state is None on CPT)."""
    B, S, H = h.shape
    nH = cfg.num_attention_heads
    dH = H // nH

    def split(x):  # transpose_for_scores
        return x.view(B, S, nH, dH).permute(0, 2, 1, 3)

    q = split(_linear(h, sd, prefix + ".query"))
    k = split(_linear(h, sd, prefix + ".key"))
    v = split(_linear(h, sd, prefix + ".value"))
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dH)
    scores = scores + ext_mask
    probs = torch.softmax(scores, dim=-1)
    probs = _dropout(probs, cfg.attention_probs_dropout_prob, training, ("attn", layer))
    ctx = torch.matmul(probs, v)
    return ctx.permute(0, 2, 1, 3).contiguous().view(B, S, H), probs


def encoder_layer(sd, cfg, i, h, ext_mask, training=False):
    """CaptionBertLayer.forward, modeling_bert.py:139-147 (+ BertSelfOutput/Intermediate/Output)."""
    p = "bert.encoder.layer.%d." % i
    ctx, probs = self_attention(sd, cfg, p + "attention.self", h, ext_mask, training, i)
    a = _dropout(_linear(ctx, sd, p + "attention.output.dense"), cfg.hidden_dropout_prob, training, ("ao", i))
    a = _ln(a + h, sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"],
            cfg.layer_norm_eps)
    inter = _gelu(_linear(a, sd, p + "intermediate.dense"))
    o = _dropout(_linear(inter, sd, p + "output.dense"), cfg.hidden_dropout_prob, training, ("down", i))
    o = _ln(o + a, sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], cfg.layer_norm_eps)
    return o, probs


def bert_img_model(sd, cfg, input_ids, token_type_ids=None, attention_mask=None, position_ids=None,
                   img_feats=None, training=False, collect_hidden=False):
    """BertImgModel.forward, modeling_bert.py:199-279.  Returns (seq_out, pooled, [hidden states])."""
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    ext = extended_attention_mask(attention_mask)
    h = text_embeddings(sd, cfg, input_ids, token_type_ids, position_ids, training)
    if img_feats is not None:
        h = torch.cat((h, region_embeddings(sd, cfg, img_feats, training)), 1)  # text first, regions last
    hidden = [h] if collect_hidden else None
    for i in range(cfg.num_hidden_layers):
        h, _ = encoder_layer(sd, cfg, i, h, ext, training)
        if collect_hidden:
            hidden.append(h)
    pooled = torch.tanh(_linear(h[:, 0], sd, "bert.pooler.dense"))  # BertPooler
    return h, pooled, hidden


def lm_head(sd, cfg, x, vocab_ids=None, prefix="cls.predictions"):
    """BertLMPredictionHead.forward: decoder(LN(gelu(dense(x)))) + bias; decoder weight is tied to
    the word embeddings (modeling_rec.py:130-135).  vocab_ids restricts the decoder columns
    (== computing all V columns and indexing them afterwards)."""
    t = _gelu(_linear(x, sd, prefix + ".transform.dense"))
    t = _ln(t, sd[prefix + ".transform.LayerNorm.weight"], sd[prefix + ".transform.LayerNorm.bias"],
            cfg.layer_norm_eps)
    W = sd["bert.embeddings.word_embeddings.weight"]
    b = sd[prefix + ".bias"]
    if vocab_ids is not None:
        W, b = W[vocab_ids], b[vocab_ids]
    return F.linear(t, W) + b


def rec_mlm_cpt(sd, cfg, input_ids, token_type_ids=None, attention_mask=None, masked_lm_labels=None,
                position_ids=None, img_feats=None, training=False):
    """REC_MLM_CPT.forward, modeling_rec.py:137-152 -> ((loss,) scores[B,S,V])."""
    seq, _, _ = bert_img_model(sd, cfg, input_ids, token_type_ids, attention_mask, position_ids, img_feats,
                               training)
    scores = lm_head(sd, cfg, seq)
    out = (scores,)
    if masked_lm_labels is not None:
        loss = F.cross_entropy(scores.view(-1, cfg.vocab_size), masked_lm_labels.view(-1), ignore_index=-1)
        out = (loss,) + out
    return out


def cpt_mlm_logits(sd, cfg, input_ids, token_type_ids, attention_mask, img_feats, mask_pos, vocab_ids,
                   position_ids=None):
    """What the CPT callers consume: scores[arange(B), mask_pos][:, vocab_ids]
    (zeroshot/refcoco_cpt.py:217-219,234-235; gqa_cpt.py:597-600).  Computed gather-first, which
    is algebraically identical because the head is row-wise."""
    seq, _, _ = bert_img_model(sd, cfg, input_ids, token_type_ids, attention_mask, position_ids, img_feats)
    rows = seq[torch.arange(seq.size(0)), mask_pos]
    return lm_head(sd, cfg, rows, vocab_ids)


def nsp_cpt(sd, cfg, input_ids, token_type_ids=None, attention_mask=None, next_sentence_label=None,
            position_ids=None, img_feats=None, training=False):
    """NSPCPT.forward, modeling_vcr.py:115-129 (cls == pretraining seq_relationship Linear)."""
    _, pooled, _ = bert_img_model(sd, cfg, input_ids, token_type_ids, attention_mask, position_ids, img_feats,
                                  training)
    score = _linear(pooled, sd, "cls.seq_relationship")
    out = (score,)
    if next_sentence_label is not None:
        loss = F.cross_entropy(score.view(-1, score.size(-1)), next_sentence_label.view(-1), ignore_index=-1)
        out = (loss,) + out
    return out


def vcr_choice_scores(nsp_logits):
    """vcr_nsp_cpt.py:600 : 1 - softmax(out)[:, 1]."""
    return 1.0 - torch.softmax(nsp_logits, -1)[:, 1]


def refcoco_zsl_pick(color_logits):
    """zeroshot/refcoco_cpt.py:242-245 for one query: rows = proposals, columns = colours + ["none"];
    keep [:-1] of every row, concatenate, argmax."""
    return int(color_logits[:, :-1].reshape(-1).argmax())


def refcoco_fsl_pick(color_logits):
    """fewshot/refcoco_cpt.py:291-294: score = colour / none, argmax."""
    return int((color_logits[:, :-1] / color_logits[:, -1:]).reshape(-1).argmax())


def box_iou_xywh(a, b):
    """Oscar/oscar/utils/iou.py:1-12 restated: boxes are [x, y, w, h] in inclusive pixel coordinates (right edge
    x + w - 1); the overlap only counts when it is strictly more than one pixel wide and high."""
    left, top = max(a[0], b[0]), max(a[1], b[1])
    right = min(a[0] + a[2] - 1, b[0] + b[2] - 1)
    bottom = min(a[1] + a[3] - 1, b[1] + b[3] - 1)
    overlap = (right - left + 1) * (bottom - top + 1) if (left < right and top < bottom) else 0
    return float(overlap) / (a[2] * a[3] + b[2] * b[3] - overlap)


def refcoco_decide(scores_rows, colour_sets, rect_sets, few_shot=False):
    """One image of the reference's val() loop (zeroshot/refcoco_cpt.py:224-246, fewshot/refcoco_cpt.py:273-294).
    scores_rows: one [K] tensor per proposal set, ALREADY gathered at the palette ids + "none" (last entry);
    colour_sets[j]: how many palette colours proposal set j uses (its rectangles' count); rect_sets[j]: its rectangles.
    Each row contributes its own colour entries (score, or score / none for few-shot); the entries of all rows are
    concatenated and the argmax position selects the rectangle.  Returns (max_idx, rect)."""
    rects, pieces = [], []
    for row, n, rs in zip(scores_rows, colour_sets, rect_sets):
        assert len(rs) == n
        mine = torch.cat((row[:n], row[-1:]))          # scores[ptr][color2id(cur_color_set + ["none"])]
        pieces.append(mine[:-1] / mine[-1] if few_shot else mine[:-1])
        rects += list(rs)
    idx = int(torch.cat(pieces, -1).argmax())
    return idx, rects[idx]


def refcoco_hit(rect, gt_xywh):
    """zeroshot/refcoco_cpt.py:268-276: the picked [x1, y1, x2, y2] becomes [x, y, w + 1, h + 1] and counts as correct
    when its IoU with the ground-truth box exceeds 0.5."""
    assert rect[2] > rect[0] and rect[3] > rect[1]
    box = [rect[0], rect[1], rect[2] - rect[0] + 1, rect[3] - rect[1] + 1]
    v = box_iou_xywh(box, gt_xywh)
    return v, v > 0.5
