"""Drop-in counterparts of the reference's cross-modal BERT modules, backed by the sm_100a engine.

Mirrors the public surface of /root/reference/Oscar/oscar/modeling/modeling_bert.py that the CPT callers use
(SURVEY.md 8b): `BertImgModel` (:150-279), `BertPreTrainingHeads` (:914-924), `BertImgForPreTraining`
(:927-1021) and the `from_pretrained / save_pretrained / tie_weights` subset of
oscar/modeling/modeling_utils.py:680-875 — same constructor and forward signatures, same positional order,
same return tuples, same state_dict keys.  The nn.Modules below only HOLD parameters (so optimizers, DDP,
state_dict, .to() see exactly the reference's tensors); the arithmetic of forward() runs in cpt_b200/csrc
through the C ABI.  Features the CPT path never uses raise instead of silently differing: head_mask,
output_attentions, encoder_history_states, 3-D attention masks, dis_code* feature types, and a label-free forward in
train() mode with active dropout (dropout belongs to the native training step, cpt_b200/training.py, reached through
the task wrappers' calls with labels).
"""
import logging
import os
import threading

import torch
from torch import nn

from .config import BertConfig  # noqa: F401  (re-exported like the reference module does)
from .engine import CptError, Engine

logger = logging.getLogger(__name__)
WEIGHTS_NAME = "pytorch_model.bin"

# "Did a parameter change since the handle converted its 16-bit copies?" answered without walking the module tree: every
# event that can MOVE or replace a parameter bumps this epoch — any torch.optim.Optimizer step (global post-step hook: covers torch's optimizers,
# pytorch-transformers' AdamW, apex, cpt_b200.optimization.AdamW), load_state_dict, Module._apply (.to / .cuda / .half),
# weight tying — and a native backward sets the slot's dirty flags; in-place edits (`p.mul_()`, `p.copy_()`) show up in
# the sum of the version counters of the cached tensor list.  Only when one of these moved do engine() / train_engine()
# walk the modules and rebuild the ~200-entry (data_ptr, version) signature.  Code that writes parameters behind all of
# them (`p.data.copy_()` in an inference loop, no optimizer) calls `model.bert.mark_weights_changed()`.
_WEIGHT_EPOCH = [0]
# Optimizer steps are counted separately: pytorch-transformers 1.x AdamW and apex update through `p.data`, which leaves
# the autograd version counters untouched, so after ANY optimizer step the 16-bit copies are refreshed even when the
# (data_ptr, version) signature did not move.
_OPT_EPOCH = [0]


def _bump_weight_epoch(*_a, **_k):
    _WEIGHT_EPOCH[0] += 1


def _on_optimizer_step(*_a, **_k):
    _WEIGHT_EPOCH[0] += 1
    _OPT_EPOCH[0] += 1


try:  # torch >= 2.0
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_opt_hook
    _reg_opt_hook(_on_optimizer_step)
except Exception:  # pragma: no cover - older torch: fall back to scanning on every call
    _WEIGHT_EPOCH = None
BertLayerNorm = nn.LayerNorm


# ----------------------------------------------------------------------------------------------------------------
# Parameter containers.  Attribute names are dictated by the checkpoint's state_dict keys.
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise CptError("cpt_b200: sub-modules only hold parameters; call the model's forward()")


class BertEmbeddings(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size)
        self.token_type_embeddings = nn.Embedding(cfg.type_vocab_size, cfg.hidden_size)
        self.LayerNorm = BertLayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)


class _SelfAttention(_Holder):
    def __init__(self, cfg):
        super().__init__()
        H = cfg.hidden_size
        self.query, self.key, self.value = nn.Linear(H, H), nn.Linear(H, H), nn.Linear(H, H)
        self.dropout = nn.Dropout(cfg.attention_probs_dropout_prob)


class _DenseLN(_Holder):
    def __init__(self, n_in, cfg):
        super().__init__()
        self.dense = nn.Linear(n_in, cfg.hidden_size)
        self.LayerNorm = BertLayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)


class _Attention(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.self = _SelfAttention(cfg)
        self.output = _DenseLN(cfg.hidden_size, cfg)


class _Dense(_Holder):
    def __init__(self, n_in, n_out):
        super().__init__()
        self.dense = nn.Linear(n_in, n_out)


class CaptionBertLayer(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.attention = _Attention(cfg)
        self.intermediate = _Dense(cfg.hidden_size, cfg.intermediate_size)
        self.output = _DenseLN(cfg.intermediate_size, cfg)


class CaptionBertEncoder(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([CaptionBertLayer(cfg) for _ in range(cfg.num_hidden_layers)])


class BertPooler(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.dense = nn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.activation = nn.Tanh()


class BertPredictionHeadTransform(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.dense = nn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.LayerNorm = BertLayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


class BertLMPredictionHead(_Holder):
    """decoder(LN(gelu(dense(x)))) + bias; decoder.weight is tied to the word embeddings."""

    def __init__(self, cfg):
        super().__init__()
        self.transform = BertPredictionHeadTransform(cfg)
        self.decoder = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(cfg.vocab_size))

    def head_tensors(self):
        t = self.transform
        return {"cls.predictions.transform.dense.weight": t.dense.weight,
                "cls.predictions.transform.dense.bias": t.dense.bias,
                "cls.predictions.transform.LayerNorm.weight": t.LayerNorm.weight,
                "cls.predictions.transform.LayerNorm.bias": t.LayerNorm.bias,
                "cls.predictions.bias": self.bias}


class BertPreTrainingHeads(_Holder):
    """modeling_bert.py:914-924: predictions (MLM) + seq_relationship Linear(H, num_contrast_classes)."""

    def __init__(self, cfg):
        super().__init__()
        self.predictions = BertLMPredictionHead(cfg)
        self.seq_relationship = nn.Linear(cfg.hidden_size, getattr(cfg, "num_contrast_classes", 2))


def nsp_head_tensors(linear):
    return {"cls.seq_relationship.weight": linear.weight, "cls.seq_relationship.bias": linear.bias}


class _EngineSlot(object):
    """Holds the (non-copyable, non-picklable) native handle; deepcopy / pickle yield an empty slot.

    nn.DataParallel (the non-distributed multi-GPU branch of the reference scripts, gqa_cpt.py:358-359) replicates a
    module by copying its __dict__, so every replica shares this object while its parameters live on another device
    and its forward runs on another thread: `for_device` hands each device its own slot (own handle, own head
    tensors), created under a lock."""

    def __init__(self):
        self.engine, self.sig, self.heads, self.frozen = None, None, {}, False
        self.train_engine, self.train_sig = None, None
        self.epoch, self.train_epoch = -1, -1  # value of _WEIGHT_EPOCH the signatures were taken at
        self.train_named = None
        self.opt_epoch, self.train_opt_epoch = 0, 0
        self.tensors, self.version_sum, self.train_version_sum = [], -1, -1
        # Set by every native backward (training._Loss.backward): an optimizer step may follow, and optimizers that
        # update through `p.data` (pytorch-transformers 1.x AdamW, apex) do NOT bump the autograd version counters the
        # signatures below key on — so after a backward both handles refresh their 16-bit copies unconditionally.
        self.dirty_train, self.dirty_infer = False, False
        self.grad_sync_group, self.grad_sync_dtype = None, "auto"  # comm.enable_overlapped_grad_sync
        self.home, self._per_device, self._lock = None, {}, threading.Lock()

    def for_device(self, dev):
        if dev.type != "cuda":
            return self
        with self._lock:
            if self.home is None:
                self.home = dev
            if self.home == dev:
                return self
            s = self._per_device.get(dev)
            if s is None:
                s = _EngineSlot()
                # replicas get freshly broadcast parameter tensors on every forward: never skip the weight check
                s.home, s.frozen, s.grad_sync_group = dev, False, self.grad_sync_group
                s.grad_sync_dtype = self.grad_sync_dtype
                self._per_device[dev] = s
            return s

    def apply_grad_sync(self, eng):
        eng.grad_sync_group = self.grad_sync_group
        want = self.grad_sync_dtype
        if want == "auto":
            want = "bf16" if eng.dtype == "bf16" else "fp32"
        eng.grad_sync_dtype = torch.bfloat16 if (want == "bf16" and self.grad_sync_group is not None) else None

    def __deepcopy__(self, memo):
        return _EngineSlot()

    def __getstate__(self):
        return {}

    def __setstate__(self, st):
        self.__init__()


# ----------------------------------------------------------------------------------------------------------------
class BertPreTrainedModel(nn.Module):
    """The subset of pytorch-transformers' PreTrainedModel / Oscar's ImgPreTrainedModel the CPT scripts call."""
    config_class = BertConfig
    base_model_prefix = "bert"

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        self.config = config

    def init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, BertLayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def _apply(self, fn, *a, **k):  # .to() / .cuda() / .float(): storages move
        _bump_weight_epoch()
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):  # reached from load_state_dict of this module or of any ancestor
        _bump_weight_epoch()
        return super()._load_from_state_dict(*a, **k)

    def _tie_or_clone_weights(self, first_module, second_module):
        _bump_weight_epoch()
        if getattr(self.config, "torchscript", False):
            first_module.weight = nn.Parameter(second_module.weight.clone())
        else:
            first_module.weight = second_module.weight

    def tie_weights(self):
        pass

    def save_pretrained(self, save_directory):
        assert os.path.isdir(save_directory), "save_pretrained needs an existing directory"
        model = self.module if hasattr(self, "module") else self
        model.config.save_pretrained(save_directory)
        torch.save(model.state_dict(), os.path.join(save_directory, WEIGHTS_NAME))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs):
        """Loads `<dir>/pytorch_model.bin` the way modeling_utils.py:803-851 does: legacy gamma/beta renamed,
        `bert.` prefix added or stripped to fit the receiving class, weights tied, model returned in eval()."""
        config = kwargs.pop("config", None)
        state_dict = kwargs.pop("state_dict", None)
        for k in ("cache_dir", "from_tf", "output_loading_info"):
            kwargs.pop(k, None)
        path = pretrained_model_name_or_path
        if config is None:
            config = cls.config_class.from_pretrained(path)
        model = cls(config, *model_args, **kwargs)
        if state_dict is None:
            f = os.path.join(path, WEIGHTS_NAME) if os.path.isdir(path) else path
            state_dict = torch.load(f, map_location="cpu")
        fixed = {}
        for k, v in state_dict.items():
            nk = k.replace("gamma", "weight") if "gamma" in k else k
            nk = nk.replace("beta", "bias") if "beta" in nk else nk
            fixed[nk] = v
        pfx = cls.base_model_prefix + "."
        has_pfx = any(k.startswith(pfx) for k in fixed)
        if hasattr(model, cls.base_model_prefix) and not has_pfx:
            fixed = {pfx + k: v for k, v in fixed.items()}
        elif not hasattr(model, cls.base_model_prefix) and has_pfx:
            fixed = {k[len(pfx):]: v for k, v in fixed.items() if k.startswith(pfx)}
        missing, unexpected = model.load_state_dict(fixed, strict=False)
        if missing:
            logger.info("Weights of %s not initialized from pretrained model: %s", cls.__name__, missing)
        if unexpected:
            logger.info("Weights from pretrained model not used in %s: %s", cls.__name__, unexpected)
        if hasattr(model, "tie_weights"):
            model.tie_weights()
        model.eval()
        return model


class BertImgModel(BertPreTrainedModel):
    """Text + region-feature BERT encoder (reference: modeling_bert.py:150-279)."""

    def __init__(self, config):
        super().__init__(config)
        self.embeddings = BertEmbeddings(config)
        self.encoder = CaptionBertEncoder(config)
        self.pooler = BertPooler(config)
        self.img_dim = config.img_feature_dim
        self.img_feature_type = config.img_feature_type
        self.use_img_layernorm = getattr(config, "use_img_layernorm", None)
        if str(self.img_feature_type).startswith("dis_code"):
            raise NotImplementedError("cpt_b200: img_feature_type '%s' is not on the CPT path (reference "
                                      "modeling_bert.py:167-176); only region-feature input is built"
                                      % self.img_feature_type)
        self.img_embedding = nn.Linear(self.img_dim, config.hidden_size, bias=True)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        if self.use_img_layernorm:
            self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.img_layer_norm_eps)
        self.apply(self.init_weights)
        self._slot = _EngineSlot()

    # -- engine plumbing ----------------------------------------------------------------------------------------
    def _named_tensors(self):
        # like state_dict(keep_vars=True), but also valid on nn.DataParallel replicas, whose (non-leaf) parameter
        # copies are plain attributes recorded in `_former_parameters` instead of `_parameters`
        sd = {}
        for mname, sub in self.named_modules():
            pre = "bert." + (mname + "." if mname else "")
            for src in (sub._parameters, getattr(sub, "_former_parameters", {}), sub._buffers):
                for k, v in src.items():
                    if v is not None:
                        sd[pre + k] = v
        sd.update(self._dev_slot().heads)
        return sd

    def _dev_slot(self):
        return self._slot.for_device(self.embeddings.word_embeddings.weight.device)

    def register_head_tensors(self, tensors):
        """Called by the task wrappers (REC_MLM_CPT, NSPCPT, BertImgForPreTraining) so their head weights ride in
        the same native handle as the encoder."""
        slot = self._dev_slot()
        cur = slot.heads
        if any(cur.get(k) is not v for k, v in tensors.items()):
            cur.update(tensors)
            slot.sig = None
            slot.epoch = slot.train_epoch = -1

    def freeze_engine_weights(self, frozen=True):
        """Skip the 'did any parameter change?' check altogether (kept for callers of round 1; the check is O(1) now)."""
        self._slot.frozen = frozen

    def mark_weights_changed(self):
        """For code that writes parameters without an optimizer, load_state_dict or .to(): forces the next forward to
        re-scan the parameters and refresh the handle's 16-bit copies."""
        _bump_weight_epoch()


    def engine(self):
        slot = self._dev_slot()
        if slot.engine is not None and slot.sig is not None and not slot.dirty_infer and (
                slot.frozen or (_WEIGHT_EPOCH is not None and slot.epoch == _WEIGHT_EPOCH[0]
                                and slot.version_sum == sum(t._version for t in slot.tensors))):
            # nothing that can move or rewrite a parameter happened (epoch), and no in-place edit either (the sum of
            # the autograd version counters over the cached tensor list: ~10 us, no module walk, no tuple building)
            return slot.engine
        epoch = _WEIGHT_EPOCH[0] if _WEIGHT_EPOCH is not None else -1
        sd = self._named_tensors()
        dev = self.embeddings.word_embeddings.weight.device
        if dev.type != "cuda":
            raise CptError("cpt_b200: the model is on '%s'; this implementation runs on a CUDA sm_100a device only "
                           "(no CPU path) — call model.to('cuda') first" % dev)
        sig = tuple((k, t.data_ptr(), t._version) for k, t in sd.items())
        if slot.engine is None or slot.engine.device != dev:
            if slot.engine is not None:
                slot.engine.close()
            dtype = getattr(self.config, "cpt_b200_dtype", None) or os.environ.get("CPT_B200_DTYPE", "fp16")
            slot.engine = Engine(self.config, dev, dtype=dtype)
            slot.sig = None
        if sig != slot.sig or slot.dirty_infer or slot.opt_epoch != _OPT_EPOCH[0]:
            slot.engine.load_state_dict(sd)
            slot.sig, slot.dirty_infer, slot.opt_epoch = sig, False, _OPT_EPOCH[0]
        slot.epoch = epoch if slot is self._slot else -1  # DataParallel replicas get fresh tensors every forward
        slot.tensors = list(sd.values())
        slot.version_sum = sum(t._version for t in slot.tensors)
        return slot.engine

    def train_engine(self):
        """The handle of the training step: 16-bit GEMM copies in bf16 by default (gradients span more binades than
        fp16 holds without loss scaling; config.cpt_b200_train_dtype / CPT_B200_TRAIN_DTYPE override), transposed
        copies for the dgrad GEMMs, refreshed in place from the fp32 parameters whenever one of them changed."""
        slot = self._dev_slot()
        if (slot.train_engine is not None and slot.train_sig is not None and not slot.dirty_train
                and _WEIGHT_EPOCH is not None and slot.train_epoch == _WEIGHT_EPOCH[0] and slot.train_named is not None
                and slot.train_version_sum == sum(t._version for t in slot.train_named.values())):
            return slot.train_engine, slot.train_named
        epoch = _WEIGHT_EPOCH[0] if _WEIGHT_EPOCH is not None else -1
        sd = self._named_tensors()
        dev = self.embeddings.word_embeddings.weight.device
        if dev.type != "cuda":
            raise CptError("cpt_b200: the model is on '%s'; this implementation runs on a CUDA sm_100a device only "
                           "(no CPU path) — call model.to('cuda') first" % dev)
        sig = tuple((k, t.data_ptr(), t._version) for k, t in sd.items())
        if slot.train_engine is None or slot.train_engine.device != dev:
            if slot.train_engine is not None:
                slot.train_engine.close()
            dtype = (getattr(self.config, "cpt_b200_train_dtype", None)
                     or os.environ.get("CPT_B200_TRAIN_DTYPE", "bf16"))
            slot.train_engine = Engine(self.config, dev, dtype=dtype, train=True)
            slot.apply_grad_sync(slot.train_engine)
            slot.train_sig = None
        slot.train_engine.owner_slot = slot
        if sig != slot.train_sig or slot.dirty_train or slot.train_opt_epoch != _OPT_EPOCH[0]:
            slot.train_engine.load_state_dict(sd)
            slot.train_sig, slot.dirty_train, slot.train_opt_epoch = sig, False, _OPT_EPOCH[0]
        slot.train_epoch = epoch if slot is self._slot else -1
        slot.train_named = sd
        slot.train_version_sum = sum(t._version for t in sd.values())
        return slot.train_engine, sd

    def _dropout_active(self):
        return self.training and (self.config.hidden_dropout_prob > 0 or self.config.attention_probs_dropout_prob > 0)

    def _check_mode(self):
        if self._dropout_active():
            raise NotImplementedError("cpt_b200: a label-free forward in train() mode with active dropout has no "
                                      "native path (dropout is applied by the training step, i.e. the call with "
                                      "masked_lm_labels / next_sentence_label) — call model.eval() for inference")

    # -- forward ------------------------------------------------------------------------------------------------
    def forward(self, input_ids, token_type_ids=None, attention_mask=None, position_ids=None, head_mask=None,
                img_feats=None, encoder_history_states=None):
        return self._encode(input_ids, token_type_ids, attention_mask, position_ids, head_mask, img_feats,
                            encoder_history_states, want_pooled=True)

    def _cpt_logits(self, input_ids, token_type_ids, attention_mask, position_ids, img_feats, mask_pos, vocab_ids,
                    gather=None):
        if getattr(self.config, "output_attentions", False):
            raise NotImplementedError("cpt_b200: attention probabilities never leave the SM (output_attentions)")
        if attention_mask is not None and attention_mask.dim() != 2:
            raise NotImplementedError("cpt_b200: only 2-D attention masks are supported")
        self._check_mode()
        if attention_mask is not None and attention_mask.dtype != torch.int64:
            attention_mask = attention_mask.to(torch.int64)
        return self.engine().cpt_logits(input_ids, token_type_ids, attention_mask, position_ids, img_feats, mask_pos,
                                        vocab_ids, exchange=gather)

    def _encode(self, input_ids, token_type_ids=None, attention_mask=None, position_ids=None, head_mask=None,
                img_feats=None, encoder_history_states=None, want_pooled=True):
        """forward() with the BertPooler made optional: the MLM wrappers never read pooled_output (the reference
        computes it anyway, modeling_bert.py:275), so they pass want_pooled=False and get None in its place."""
        if head_mask is not None:
            raise NotImplementedError("cpt_b200: head_mask is never used on the CPT path (modeling_bert.py:60-61)")
        if encoder_history_states is not None:
            raise NotImplementedError("cpt_b200: encoder_history_states is a captioning-only feature")
        if getattr(self.config, "output_attentions", False):
            raise NotImplementedError("cpt_b200: attention probabilities never leave the SM (output_attentions)")
        if attention_mask is not None and attention_mask.dim() != 2:
            # the reference accepts 3-D masks (captioning) and raises NotImplementedError for anything else
            raise NotImplementedError("cpt_b200: only 2-D attention masks are supported")
        self._check_mode()
        eng = self.engine()
        if attention_mask is not None and attention_mask.dtype != torch.int64:
            attention_mask = attention_mask.to(torch.int64)
        want_hidden = bool(getattr(self.config, "output_hidden_states", False))
        seq, pooled, hidden = eng.encoder_forward(input_ids, token_type_ids, attention_mask, position_ids, img_feats,
                                                  want_pooled=want_pooled, want_hidden=want_hidden)
        out = (seq, pooled)
        if want_hidden:
            out = out + (tuple(hidden.unbind(0)),)
        return out


class BertImgForPreTraining(BertPreTrainedModel):
    """The container Oscar/VinVL checkpoints are loaded into (reference: modeling_bert.py:927-1021).  Its forward
    returns (prediction_scores[B,S,V], seq_relationship_score[B,C]) like the reference's inference branch."""

    def __init__(self, config):
        super().__init__(config)
        self.bert = BertImgModel(config)
        self.cls = BertPreTrainingHeads(config)
        self.num_seq_relations = getattr(config, "num_contrast_classes", 2)
        self.apply(self.init_weights)
        self.tie_weights()

    def tie_weights(self):
        self._tie_or_clone_weights(self.cls.predictions.decoder, self.bert.embeddings.word_embeddings)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, masked_lm_labels=None,
                next_sentence_label=None, position_ids=None, head_mask=None, img_feats=None):
        if masked_lm_labels is not None or next_sentence_label is not None:
            raise NotImplementedError("cpt_b200: the pre-training loss is outside the CPT path")
        heads = self.cls.predictions.head_tensors()
        heads.update(nsp_head_tensors(self.cls.seq_relationship))
        self.bert.register_head_tensors(heads)
        outputs = self.bert(input_ids, position_ids=position_ids, token_type_ids=token_type_ids,
                            attention_mask=attention_mask, head_mask=head_mask, img_feats=img_feats)
        eng = self.bert.engine()
        return (eng.mlm_scores(outputs[0]), eng.nsp(outputs[1])) + outputs[2:]
