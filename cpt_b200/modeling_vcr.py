"""NSPCPT — the CPT next-sentence-style scorer for VCR, backed by the sm_100a engine.

Reference: /root/reference/Oscar/oscar/modeling/modeling_vcr.py:79-129: the pooled [CLS] vector goes through the
pre-training `seq_relationship` Linear(H, num_contrast_classes); the caller scores a choice as
1 - softmax(out)[:, 1] (fewshot/vcr_nsp_cpt.py:600).
"""
import copy

import torch
from torch import nn

from .modeling_bert import BertImgModel, BertLMPredictionHead, BertPreTrainedModel, nsp_head_tensors


class NSPCPT(BertPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.bert = BertImgModel(config)
        self.cls = BertLMPredictionHead(config)  # replaced by seq_relationship in copy_from_pretraining_model
        self.num_seq_relations = getattr(config, "num_contrast_classes", 2)
        self.apply(self.init_weights)
        self.tie_weights()

    def copy_from_pretraining_model(self, model, possible_colors=[]):
        self.bert = model.bert
        self.cls = model.cls.seq_relationship

    def tie_weights(self):
        if isinstance(self.cls, BertLMPredictionHead):
            self._tie_or_clone_weights(self.cls.decoder, self.bert.embeddings.word_embeddings)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, next_sentence_label=None,
                position_ids=None, head_mask=None, img_feats=None):
        if not isinstance(self.cls, nn.Linear):
            raise RuntimeError("cpt_b200: NSPCPT scores with the pre-training seq_relationship head; call "
                               "copy_from_pretraining_model(BertImgForPreTraining) first (modeling_vcr.py:90-92)")
        self.bert.register_head_tensors(nsp_head_tensors(self.cls))
        if next_sentence_label is not None and ((torch.is_grad_enabled()
                                                 and any(p.requires_grad for p in self.parameters()))
                                                or self.bert._dropout_active()):
            return self._train_step(input_ids, token_type_ids, attention_mask, next_sentence_label, position_ids,
                                    head_mask, img_feats)
        if (head_mask is None and not getattr(self.config, "output_hidden_states", False)
                and not getattr(self.config, "output_attentions", False)
                and (attention_mask is None or attention_mask.dim() == 2)):
            # the VCR inference call: one fused (and CUDA-graph-cached) encoder + pooler + head launch sequence
            self.bert._check_mode()
            if attention_mask is not None and attention_mask.dtype != torch.int64:
                attention_mask = attention_mask.to(torch.int64)
            score = self.bert.engine().nsp_scores(input_ids, token_type_ids, attention_mask, position_ids, img_feats)
            out = (score,)
        else:
            outputs = self.bert(input_ids, position_ids=position_ids, token_type_ids=token_type_ids,
                                attention_mask=attention_mask, head_mask=head_mask, img_feats=img_feats)
            score = self.bert.engine().nsp(outputs[1])
            out = (score,) + outputs[2:]
        if next_sentence_label is not None:
            loss = nn.functional.cross_entropy(score.view(-1, self.num_seq_relations), next_sentence_label.view(-1),
                                               ignore_index=-1)
            out = (loss,) + out
        return out

    def _train_step(self, input_ids, token_type_ids, attention_mask, next_sentence_label, position_ids, head_mask,
                    img_feats):
        """(loss, None) — vcr_nsp_cpt.py:445-461 reads `loss, logits = outputs[:2]` and only uses the loss.  Native
        forward + backward behind autograd (cpt_b200/training.py)."""
        from .training import draw_dropout, nsp_loss
        if head_mask is not None:
            raise NotImplementedError("cpt_b200: head_mask is never used on the CPT path")
        if getattr(self.config, "output_hidden_states", False) or getattr(self.config, "output_attentions", False):
            raise NotImplementedError("cpt_b200: output_hidden_states / output_attentions in the training step")
        if attention_mask is not None and attention_mask.dim() != 2:
            raise NotImplementedError("cpt_b200: only 2-D attention masks are supported")
        if attention_mask is not None and attention_mask.dtype != torch.int64:
            attention_mask = attention_mask.to(torch.int64)
        eng, named = self.bert.train_engine()
        self.last_dropout = draw_dropout(self.config, self.training)
        loss, _ = nsp_loss(eng, named, input_ids, token_type_ids, attention_mask, position_ids, img_feats,
                           next_sentence_label, self.last_dropout)
        return (loss, None)


class VCRQAR_NSPCPT(BertPreTrainedModel):
    """The two-head VCR variant (question -> answer and question + answer -> rationale share one encoder;
    reference: Oscar/oscar/modeling/modeling_vcr.py:194-252): `cls_ans` IS the pre-training seq_relationship head,
    `cls_rat` a deep copy of it; `head="ans" | "rat"` picks one per call.  Same constructor,
    copy_from_pretraining_model and forward signature / return tuple as the reference."""

    def __init__(self, config):
        super().__init__(config)
        self.bert = BertImgModel(config)
        self.cls_ans = None
        self.cls_rat = None
        self.num_seq_relations = getattr(config, "num_contrast_classes", 2)
        self.apply(self.init_weights)

    def copy_from_pretraining_model(self, model, possible_colors=[]):
        self.bert = model.bert
        self.cls_ans = model.cls.seq_relationship
        self.cls_rat = copy.deepcopy(model.cls.seq_relationship)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, next_sentence_label=None,
                position_ids=None, head_mask=None, img_feats=None, head=""):
        cls = {"ans": self.cls_ans, "rat": self.cls_rat}.get(head)
        if cls is None:
            # the reference would call None(...) / view() on None here: an error either way
            raise RuntimeError("cpt_b200: VCRQAR_NSPCPT needs head='ans' or head='rat' and "
                               "copy_from_pretraining_model(BertImgForPreTraining) first (modeling_vcr.py:208-211,236-240)")
        if next_sentence_label is not None and ((torch.is_grad_enabled()
                                                 and any(p.requires_grad for p in self.parameters()))
                                                or self.bert._dropout_active()):
            return self._train_step(cls, input_ids, token_type_ids, attention_mask, next_sentence_label, position_ids,
                                    head_mask, img_feats)
        outputs = self.bert(input_ids, position_ids=position_ids, token_type_ids=token_type_ids,
                            attention_mask=attention_mask, head_mask=head_mask, img_feats=img_feats)
        score = self.bert.engine().head_linear(outputs[1], cls.weight, cls.bias)
        out = (score,) + outputs[2:]
        if next_sentence_label is not None:
            loss = nn.functional.cross_entropy(score.view(-1, self.num_seq_relations), next_sentence_label.view(-1),
                                               ignore_index=-1)
            out = (loss,) + out
        return out

    def _train_step(self, cls, input_ids, token_type_ids, attention_mask, next_sentence_label, position_ids, head_mask,
                    img_feats):
        """(loss, None): the native NSP training step with the CHOSEN head in the seq_relationship slot of the training
        handle (its 16-bit copies are refreshed before every training forward anyway, so switching heads costs no extra
        pass); gradients land on `cls.weight` / `cls.bias` of that head and on the shared encoder."""
        from .training import draw_dropout, nsp_loss
        if head_mask is not None:
            raise NotImplementedError("cpt_b200: head_mask is never used on the CPT path")
        if getattr(self.config, "output_hidden_states", False) or getattr(self.config, "output_attentions", False):
            raise NotImplementedError("cpt_b200: output_hidden_states / output_attentions in the training step")
        if attention_mask is not None and attention_mask.dim() != 2:
            raise NotImplementedError("cpt_b200: only 2-D attention masks are supported")
        if attention_mask is not None and attention_mask.dtype != torch.int64:
            attention_mask = attention_mask.to(torch.int64)
        self.bert.register_head_tensors(nsp_head_tensors(cls))
        eng, named = self.bert.train_engine()
        self.last_dropout = draw_dropout(self.config, self.training)
        loss, _ = nsp_loss(eng, named, input_ids, token_type_ids, attention_mask, position_ids, img_feats,
                           next_sentence_label, self.last_dropout)
        return (loss, None)
