"""REC_MLM_CPT — the CPT masked-colour-token model for RefCOCO / GQA / VG, backed by the sm_100a engine.

Reference: /root/reference/Oscar/oscar/modeling/modeling_rec.py:100-152.  Same constructor,
`copy_from_pretraining_model`, `tie_weights` and forward signature / return tuple.  Two extra keyword-only
arguments expose what every CPT caller does right after the call
(zeroshot/refcoco_cpt.py:219,234-235, fewshot/gqa_cpt.py:597-600):

    scores = model(ids, seg, mask, img_feats=f)[0][arange(B), mask_pos][:, vocab_ids]     # reference
    logits = model(ids, seg, mask, img_feats=f, mask_pos=mask_pos, vocab_ids=vocab_ids)[0]  # same values

so the [B,S,V] score tensor (937 MB at B=64) is never formed.  Without them the full tensor is returned.

The Visual-Genome CPT caller reads SEVERAL [MASK] positions per row and needs the whole vocabulary at each of them
(fewshot/vg_cpt.py:270-271: `[out[mask_pos].softmax(-1) for out, mask_pos in zip(output, mask_token_pos)]`).  For it:

    rows = model(ids, seg, mask, img_feats=f, mask_rows=flat)[0]     # [n, V]; flat[i] = b * S + s of the i-th position

(== `output.view(-1, V)[flat]`; add `vocab_ids=` to restrict the columns).
"""
import torch
from torch import nn

from .modeling_bert import BertImgModel, BertLMPredictionHead, BertPreTrainedModel


class REC_MLM_CPT(BertPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.bert = BertImgModel(config)
        self.cls = BertLMPredictionHead(config)
        self.num_seq_relations = getattr(config, "num_contrast_classes", 2)
        self.apply(self.init_weights)
        self.tie_weights()

    def copy_from_pretraining_model(self, model, possible_colors=[]):
        self.bert = model.bert
        self.cls = model.cls.predictions
        self.tie_weights()

    def tie_weights(self):
        self._tie_or_clone_weights(self.cls.decoder, self.bert.embeddings.word_embeddings)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, masked_lm_labels=None,
                position_ids=None, head_mask=None, img_feats=None, *, mask_pos=None, vocab_ids=None, mask_rows=None, gather=None):
        if self.cls.decoder.weight is not self.bert.embeddings.word_embeddings.weight:
            raise RuntimeError("cpt_b200: cls.decoder.weight must stay tied to the word embeddings "
                               "(modeling_rec.py:130-135); call tie_weights()")
        self.bert.register_head_tensors(self.cls.head_tensors())
        # labelled call: the native training step when a gradient can be asked for — or when dropout is active (train()
        # mode under no_grad, e.g. a loss probe): only that path applies dropout
        if masked_lm_labels is not None and ((torch.is_grad_enabled() and self._any_requires_grad())
                                             or self.bert._dropout_active()):
            return self._train_step(input_ids, token_type_ids, attention_mask, masked_lm_labels, position_ids,
                                    head_mask, img_feats, mask_pos)
        if (mask_pos is not None and masked_lm_labels is None and head_mask is None
                and not getattr(self.config, "output_hidden_states", False)):
            # the CPT inference call: one fused (and CUDA-graph-cached) encoder + gathered-head launch sequence
            return (self.bert._cpt_logits(input_ids, token_type_ids, attention_mask, position_ids, img_feats, mask_pos,
                                          vocab_ids, gather=gather),)
        if gather is not None:
            raise ValueError("cpt_b200: gather= (comm.LogitsExchange) goes with the gather-first call (mask_pos=...)")
        outputs = self.bert._encode(input_ids, position_ids=position_ids, token_type_ids=token_type_ids,
                                    attention_mask=attention_mask, head_mask=head_mask, img_feats=img_feats,
                                    want_pooled=False)
        eng = self.bert.engine()
        if mask_rows is not None:
            if masked_lm_labels is not None or mask_pos is not None:
                raise ValueError("mask_rows excludes mask_pos and masked_lm_labels")
            seq = outputs[0]
            n_rows = seq.shape[0] * seq.shape[1]
            flat = mask_rows.reshape(-1)
            if flat.dtype != torch.int64 or flat.device != seq.device:
                raise ValueError("cpt_b200: mask_rows must be an int64 tensor on the model's device")
            if flat.numel():
                lo, hi = torch.aminmax(flat)     # one host read: an out-of-range row would be a device-side assert
                if int(lo) < 0 or int(hi) >= n_rows:
                    raise ValueError("cpt_b200: mask_rows holds a row outside [0, B*S = %d)" % n_rows)
            picked = seq.reshape(n_rows, -1).index_select(0, flat)
            scores = eng.mlm_scores(picked)                          # [n, V]: transform + tied decoder + bias
            if vocab_ids is not None:
                scores = scores.index_select(1, vocab_ids)
            return (scores,) + outputs[2:]
        if mask_pos is not None:
            if masked_lm_labels is not None:
                raise ValueError("mask_pos/vocab_ids (gathered logits) and masked_lm_labels are exclusive")
            return (eng.mlm_gather(outputs[0], mask_pos, vocab_ids),) + outputs[2:]
        scores = eng.mlm_scores(outputs[0])
        out = (scores,) + outputs[2:]
        if masked_lm_labels is not None:
            loss = nn.functional.cross_entropy(scores.view(-1, self.config.vocab_size), masked_lm_labels.view(-1),
                                               ignore_index=-1)
            out = (loss,) + out
        return out

    # -- training step (fewshot/refcoco_cpt.py:243-248, gqa_cpt.py:437-462) --------------------------------------
    def _any_requires_grad(self):
        return any(p.requires_grad for p in self.parameters())

    def _train_step(self, input_ids, token_type_ids, attention_mask, masked_lm_labels, position_ids, head_mask,
                    img_feats, mask_pos):
        """(loss, None): the loss is differentiable with respect to every parameter (native forward + backward,
        cpt_b200/training.py).  The reference also returns the [B,S,V] prediction_scores next to the loss; no
        caller reads them on this path (`loss, output = model(...)`), so they are not materialised."""
        from .training import draw_dropout, mlm_loss
        if head_mask is not None or mask_pos is not None:
            raise NotImplementedError("cpt_b200: head_mask / mask_pos are not supported together with masked_lm_labels")
        if getattr(self.config, "output_hidden_states", False) or getattr(self.config, "output_attentions", False):
            raise NotImplementedError("cpt_b200: output_hidden_states / output_attentions in the training step")
        if attention_mask is not None and attention_mask.dim() != 2:
            raise NotImplementedError("cpt_b200: only 2-D attention masks are supported")
        if attention_mask is not None and attention_mask.dtype != torch.int64:
            attention_mask = attention_mask.to(torch.int64)
        eng, named = self.bert.train_engine()
        self.last_dropout = draw_dropout(self.config, self.training)
        loss, _ = mlm_loss(eng, named, input_ids, token_type_ids, attention_mask, position_ids, img_feats,
                           masked_lm_labels, self.last_dropout)
        return (loss, None)
