"""Shim that lets the reference's OWN, unmodified Oscar files run in this container.

Test infrastructure only (used by tests/golden/make_golden.py and by bench.py's CPU reference legs, which run the
reference's own modules from baseline/_ref when that offline install is present).  Nothing in the product
imports this.

The reference imports its BERT building blocks from an un-vendored git clone
(`transformers.pytorch_transformers`, pinned at huggingface/transformers commit
067923d3267325f525f4e46f357360c191ba562e by /root/reference/install.sh:29-32).  That
package does not exist here, so this module registers stand-ins in `sys.modules`
that restate the published pytorch-transformers 1.x behaviour of exactly the symbols
listed at /root/reference/Oscar/oscar/modeling/modeling_bert.py:10-16 and
/root/reference/Oscar/oscar/modeling/modeling_utils.py:10-15, plus an empty `anytree`
(imported by Oscar/oscar/utils/cbs.py:5, never used on the CPT path).

With the shim installed, `oscar.modeling.modeling_bert` (attention math, wiring, region
embedding, mask build), `modeling_rec` and `modeling_vcr` execute their own code.
"""
import math
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = "/root/reference/Oscar"


class BertConfig(object):
    """Attribute bag with the fields BertConfig carried in pytorch-transformers 1.x."""

    def __init__(self, vocab_size_or_config_json_file=30522, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 max_position_embeddings=512, type_vocab_size=2, initializer_range=0.02,
                 layer_norm_eps=1e-12, **kwargs):
        self.vocab_size = vocab_size_or_config_json_file
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_act = hidden_act
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range
        self.layer_norm_eps = layer_norm_eps
        # PretrainedConfig defaults
        self.finetuning_task = kwargs.pop("finetuning_task", None)
        self.num_labels = kwargs.pop("num_labels", 2)
        self.output_attentions = kwargs.pop("output_attentions", False)
        self.output_hidden_states = kwargs.pop("output_hidden_states", False)
        self.torchscript = kwargs.pop("torchscript", False)
        self.pruned_heads = kwargs.pop("pruned_heads", {})
        for k, v in kwargs.items():
            setattr(self, k, v)


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


ACT2FN = {"gelu": gelu, "relu": torch.nn.functional.relu}
BertLayerNorm = nn.LayerNorm


class PreTrainedModel(nn.Module):
    config_class = None
    pretrained_model_archive_map = {}
    load_tf_weights = lambda model, config, path: None
    base_model_prefix = ""

    def __init__(self, config, *inputs, **kwargs):
        super(PreTrainedModel, self).__init__()
        self.config = config

    def _tie_or_clone_weights(self, first_module, second_module):
        if self.config.torchscript:
            first_module.weight = nn.Parameter(second_module.weight.clone())
        else:
            first_module.weight = second_module.weight

    def tie_weights(self):
        pass


class BertPreTrainedModel(PreTrainedModel):
    config_class = BertConfig
    base_model_prefix = "bert"

    def init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, BertLayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()


class BertEmbeddings(nn.Module):
    def __init__(self, config):
        super(BertEmbeddings, self).__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, input_ids, token_type_ids=None, position_ids=None):
        seq_length = input_ids.size(1)
        if position_ids is None:
            position_ids = torch.arange(seq_length, dtype=torch.long, device=input_ids.device)
            position_ids = position_ids.unsqueeze(0).expand_as(input_ids)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        e = (self.word_embeddings(input_ids) + self.position_embeddings(position_ids)
             + self.token_type_embeddings(token_type_ids))
        return self.dropout(self.LayerNorm(e))


class BertSelfAttention(nn.Module):
    def __init__(self, config):
        super(BertSelfAttention, self).__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError("hidden size not a multiple of the number of attention heads")
        self.output_attentions = config.output_attentions
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)

    def transpose_for_scores(self, x):
        new_x_shape = x.size()[:-1] + (self.num_attention_heads, self.attention_head_size)
        x = x.view(*new_x_shape)
        return x.permute(0, 2, 1, 3)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super(BertSelfOutput, self).__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        hidden_states = self.dropout(self.dense(hidden_states))
        return self.LayerNorm(hidden_states + input_tensor)


class BertAttention(nn.Module):
    def __init__(self, config):
        super(BertAttention, self).__init__()
        self.self = BertSelfAttention(config)
        self.output = BertSelfOutput(config)


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super(BertIntermediate, self).__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        self.intermediate_act_fn = ACT2FN[config.hidden_act] if isinstance(config.hidden_act, str) \
            else config.hidden_act

    def forward(self, hidden_states):
        return self.intermediate_act_fn(self.dense(hidden_states))


class BertOutput(nn.Module):
    def __init__(self, config):
        super(BertOutput, self).__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        hidden_states = self.dropout(self.dense(hidden_states))
        return self.LayerNorm(hidden_states + input_tensor)


class BertLayer(nn.Module):
    def __init__(self, config):
        super(BertLayer, self).__init__()
        self.attention = BertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)


class BertEncoder(nn.Module):
    def __init__(self, config):
        super(BertEncoder, self).__init__()
        self.output_attentions = config.output_attentions
        self.output_hidden_states = config.output_hidden_states
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(config.num_hidden_layers)])


class BertPooler(nn.Module):
    def __init__(self, config):
        super(BertPooler, self).__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(hidden_states[:, 0]))


class BertPredictionHeadTransform(nn.Module):
    def __init__(self, config):
        super(BertPredictionHeadTransform, self).__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.transform_act_fn = ACT2FN[config.hidden_act] if isinstance(config.hidden_act, str) \
            else config.hidden_act
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def forward(self, hidden_states):
        return self.LayerNorm(self.transform_act_fn(self.dense(hidden_states)))


class BertLMPredictionHead(nn.Module):
    def __init__(self, config):
        super(BertLMPredictionHead, self).__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))

    def forward(self, hidden_states):
        return self.decoder(self.transform(hidden_states)) + self.bias


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config):
        super(BertOnlyMLMHead, self).__init__()
        self.predictions = BertLMPredictionHead(config)

    def forward(self, sequence_output):
        return self.predictions(sequence_output)


def _import_only(name):
    def __init__(self, *a, **k):
        raise NotImplementedError("%s is un-vendored pytorch-transformers 1.x code the shim does not restate" % name)
    return type(name, (object,), {"__init__": __init__})


def install(reference_root=REFERENCE_ROOT):
    """Register the stand-in modules and put the reference's Oscar/ on sys.path."""
    import transformers  # the real (5.x) package stays importable
    if "transformers.pytorch_transformers" in sys.modules:
        return
    pkg = types.ModuleType("transformers.pytorch_transformers")
    pkg.__path__ = []
    mb = types.ModuleType("transformers.pytorch_transformers.modeling_bert")
    for name in ("BertConfig BertLayerNorm BertEmbeddings BertSelfAttention BertSelfOutput BertAttention "
                 "BertIntermediate BertOutput BertLayer BertEncoder BertPooler BertPredictionHeadTransform "
                 "BertLMPredictionHead BertOnlyMLMHead BertPreTrainedModel gelu ACT2FN").split():
        setattr(mb, name, globals()[name])
    mb.BERT_PRETRAINED_MODEL_ARCHIVE_MAP = {}
    mb.load_tf_weights_in_bert = None
    mu = types.ModuleType("transformers.pytorch_transformers.modeling_utils")
    mu.PreTrainedModel = PreTrainedModel
    mu.WEIGHTS_NAME = "pytorch_model.bin"
    mu.TF_WEIGHTS_NAME = "model.ckpt"
    fu = types.ModuleType("transformers.pytorch_transformers.file_utils")
    fu.cached_path = lambda p, **kw: p
    pkg.modeling_bert, pkg.modeling_utils, pkg.file_utils = mb, mu, fu
    # names the task scripts import at module scope (zeroshot/refcoco_cpt.py:11,24, fewshot/gqa_cpt.py:23-24); the loops
    # the drop-in test runs (val(), evaluate()) take model and tokenizer as arguments and construct none of these
    pkg.BertConfig, pkg.WEIGHTS_NAME = BertConfig, mu.WEIGHTS_NAME
    for name in ("BertTokenizer", "AdamW", "WarmupLinearSchedule", "WarmupConstantSchedule"):
        setattr(pkg, name, _import_only(name))
    for m in (pkg, mb, mu, fu):
        sys.modules[m.__name__] = m
    transformers.pytorch_transformers = pkg
    if "anytree" not in sys.modules:
        sys.modules["anytree"] = types.ModuleType("anytree")
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)


def reference_available(reference_root=REFERENCE_ROOT):
    import os
    return os.path.isfile(os.path.join(reference_root, "oscar/modeling/modeling_bert.py"))
