"""CPU: host-side logic of the training step that needs no GPU — the gradient-slab ordering contract
(include/cpt_b200.h progress callback: loss head, layer L-1 .. 0, embeddings), dropout seed drawing."""
import torch

from cpt_b200 import config as C
from cpt_b200.engine import GLOBAL_KEYS, layer_keys
from cpt_b200.training import draw_dropout, trainable_groups, trainable_keys


def test_groups_cover_every_trainable_tensor_once_in_completion_order():
    cfg = C.oscar_tiny(num_hidden_layers=3)
    for head, unused in (("mlm", {"pooler_w", "pooler_b", "nsp_w", "nsp_b"}),
                         ("nsp", {"mlm_dense_w", "mlm_dense_b", "mlm_ln_g", "mlm_ln_b", "mlm_bias"})):
        groups = trainable_groups(cfg, head, has_img=True)
        assert len(groups) == cfg.num_hidden_layers + 2
        flat = [k for g in groups for k in g]
        want = {k for f, k in GLOBAL_KEYS.items() if f not in unused}
        for i in range(cfg.num_hidden_layers):
            want |= set(layer_keys(i).values())
        assert len(flat) == len(set(flat)) and set(flat) == want
        assert flat == trainable_keys(cfg, head, True)
        # stage s (1..L) holds layer L - s; the tied word embeddings complete last
        for s in range(1, cfg.num_hidden_layers + 1):
            assert all((".layer.%d." % (cfg.num_hidden_layers - s)) in k for k in groups[s])
        assert GLOBAL_KEYS["word_emb"] in groups[-1] and GLOBAL_KEYS["word_emb"] not in groups[0]
        # query / key / value weights adjacent (one [3H,H] weight-gradient product)
        for g in groups[1:-1]:
            assert [k.split("attention.self.")[1] for k in g[:3]] == ["query.weight", "key.weight", "value.weight"]
    no_img = trainable_keys(cfg, "mlm", has_img=False)
    assert not any("img_embedding" in k or k.startswith("bert.LayerNorm") for k in no_img)


def test_draw_dropout_follows_the_torch_generator():
    cfg = C.oscar_tiny()
    cfg.hidden_dropout_prob, cfg.attention_probs_dropout_prob = 0.1, 0.2
    assert draw_dropout(cfg, training=False) is None
    torch.manual_seed(3)
    a = draw_dropout(cfg, True)
    b = draw_dropout(cfg, True)
    torch.manual_seed(3)
    assert draw_dropout(cfg, True) == a and a != b and a[:2] == (0.1, 0.2) and 0 <= a[2] < 2 ** 62
    cfg.hidden_dropout_prob = cfg.attention_probs_dropout_prob = 0.0
    assert draw_dropout(cfg, True) is None
