"""Model configuration for the CPT cross-modal BERT path.

Mirrors the attribute names the reference reads from `BertConfig`
(/root/reference/Oscar/oscar/modeling/modeling_bert.py:96-97,159-181 and
/root/reference/Oscar/oscar/run_oscarplus_pretrain.py:238-249), so a config loaded from an
Oscar/VinVL checkpoint's config.json can be passed through unchanged.
"""
import copy
import json
import os


class BertConfig(object):
    def __init__(self, vocab_size_or_config_json_file=30522, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 max_position_embeddings=512, type_vocab_size=2, initializer_range=0.02,
                 layer_norm_eps=1e-12, img_feature_dim=2054, img_feature_type="faster_r-cnn",
                 use_img_layernorm=1, img_layer_norm_eps=1e-12, num_contrast_classes=3, **kwargs):
        if isinstance(vocab_size_or_config_json_file, str):
            with open(vocab_size_or_config_json_file, "r", encoding="utf-8") as f:
                d = json.load(f)
            self.__init__(**{("vocab_size_or_config_json_file" if k == "vocab_size" else k): v
                             for k, v in d.items()})
            return
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
        self.img_feature_dim = img_feature_dim
        self.img_feature_type = img_feature_type
        self.use_img_layernorm = use_img_layernorm
        self.img_layer_norm_eps = img_layer_norm_eps
        self.num_contrast_classes = num_contrast_classes
        self.finetuning_task = kwargs.pop("finetuning_task", None)
        self.num_labels = kwargs.pop("num_labels", 2)
        self.output_attentions = kwargs.pop("output_attentions", False)
        self.output_hidden_states = kwargs.pop("output_hidden_states", False)
        self.torchscript = kwargs.pop("torchscript", False)
        self.pruned_heads = kwargs.pop("pruned_heads", {})
        for k, v in kwargs.items():
            setattr(self, k, v)

    # -- the subset of PretrainedConfig the reference scripts use --------------------------
    @classmethod
    def from_pretrained(cls, path, **kwargs):
        f = os.path.join(path, "config.json") if os.path.isdir(path) else path
        cfg = cls(f)
        for k, v in kwargs.items():
            setattr(cfg, k, v)
        return cfg

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def to_json_string(self):
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def save_pretrained(self, save_directory):
        with open(os.path.join(save_directory, "config.json"), "w", encoding="utf-8") as f:
            f.write(self.to_json_string())


def oscar_base(**kw):
    """Oscar/VinVL base (BERT-base geometry)."""
    return BertConfig(**kw)


def oscar_large(**kw):
    d = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)
    d.update(kw)
    return BertConfig(**d)


def oscar_tiny(**kw):
    """Small geometry for fast parity tests (same head size 64 as base/large)."""
    d = dict(vocab_size_or_config_json_file=1024, hidden_size=128, num_hidden_layers=2,
             num_attention_heads=2, intermediate_size=512, max_position_embeddings=256)
    d.update(kw)
    return BertConfig(**d)
