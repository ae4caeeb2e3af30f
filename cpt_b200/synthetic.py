"""Seeded synthetic weights and CPT input batches (no checkpoints or datasets exist offline).

Distributions follow SURVEY.md section 8(d):
  * Linear / Embedding weights ~ N(0, 0.02); LayerNorm gamma ~ U(0.5, 1.5); LayerNorm beta and all
    biases ~ N(0, 0.02)  (non-trivial so every epilogue term is exercised);
  * input_ids = [CLS] a.. [SEP] b.. [SEP] pad..  with one [MASK] (id 103) inside text_a, the layout
    built at /root/reference/Oscar/oscar/datasets/refcoco_zsl_cpt_dataset.py:251-296;
  * img_feats[:, :n_r, :2048] = relu(N(0,1)), last 6 columns = box geometry in [0,1]
    (/root/reference/prompt_feat/maskrcnn_benchmark/engine/inference_ref.py:263-274), padded rows zero.
"""
import torch

CLS, SEP, MASK, PAD = 101, 102, 103, 0


def state_dict_keys(cfg, with_heads=True):
    """(key, shape, kind) in a fixed order; kind in {w, b, g(amma), e(mbedding)}."""
    H, I, V, F = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.img_feature_dim
    ks = [("bert.embeddings.word_embeddings.weight", (V, H), "e"),
          ("bert.embeddings.position_embeddings.weight", (cfg.max_position_embeddings, H), "e"),
          ("bert.embeddings.token_type_embeddings.weight", (cfg.type_vocab_size, H), "e"),
          ("bert.embeddings.LayerNorm.weight", (H,), "g"), ("bert.embeddings.LayerNorm.bias", (H,), "b"),
          ("bert.img_embedding.weight", (H, F), "w"), ("bert.img_embedding.bias", (H,), "b")]
    if getattr(cfg, "use_img_layernorm", 0):
        ks += [("bert.LayerNorm.weight", (H,), "g"), ("bert.LayerNorm.bias", (H,), "b")]
    for i in range(cfg.num_hidden_layers):
        p = "bert.encoder.layer.%d." % i
        for n in ("query", "key", "value"):
            ks += [(p + "attention.self.%s.weight" % n, (H, H), "w"), (p + "attention.self.%s.bias" % n, (H,), "b")]
        ks += [(p + "attention.output.dense.weight", (H, H), "w"), (p + "attention.output.dense.bias", (H,), "b"),
               (p + "attention.output.LayerNorm.weight", (H,), "g"), (p + "attention.output.LayerNorm.bias", (H,), "b"),
               (p + "intermediate.dense.weight", (I, H), "w"), (p + "intermediate.dense.bias", (I,), "b"),
               (p + "output.dense.weight", (H, I), "w"), (p + "output.dense.bias", (H,), "b"),
               (p + "output.LayerNorm.weight", (H,), "g"), (p + "output.LayerNorm.bias", (H,), "b")]
    ks += [("bert.pooler.dense.weight", (H, H), "w"), ("bert.pooler.dense.bias", (H,), "b")]
    if with_heads:
        ks += [("cls.predictions.transform.dense.weight", (H, H), "w"),
               ("cls.predictions.transform.dense.bias", (H,), "b"),
               ("cls.predictions.transform.LayerNorm.weight", (H,), "g"),
               ("cls.predictions.transform.LayerNorm.bias", (H,), "b"),
               ("cls.predictions.bias", (V,), "b"),
               ("cls.seq_relationship.weight", (getattr(cfg, "num_contrast_classes", 2), H), "w"),
               ("cls.seq_relationship.bias", (getattr(cfg, "num_contrast_classes", 2),), "b")]
    return ks


def synth_state_dict(cfg, seed=88):
    """Pre-training-layout state dict (BertImgForPreTraining keys; decoder.weight is the tied
    word-embedding tensor and is therefore not listed separately)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape, kind in state_dict_keys(cfg):
        if kind == "g":
            sd[k] = torch.rand(shape, generator=g) + 0.5
        else:
            sd[k] = torch.randn(shape, generator=g) * 0.02
    sd["cls.predictions.decoder.weight"] = sd["bert.embeddings.word_embeddings.weight"]
    return sd


def synth_batch(cfg, B, T=70, R=50, seed=88, dense=False):
    """Returns dict(input_ids[B,T] i64, token_type_ids[B,T] i64, attention_mask[B,T+R] i64,
    img_feats[B,R,F] f32, mask_pos[B] i64)."""
    g = torch.Generator().manual_seed(seed + 1000003)
    V, F = cfg.vocab_size, cfg.img_feature_dim
    lo = min(1000, V // 2)
    ids = torch.zeros(B, T, dtype=torch.long)
    seg = torch.zeros(B, T, dtype=torch.long)
    mask = torch.zeros(B, T + R, dtype=torch.long)
    mask_pos = torch.zeros(B, dtype=torch.long)
    feats = torch.zeros(B, R, F)
    for b in range(B):
        if dense:
            n_a = max(2, (T - 3) // 3)
            n_b = T - 3 - n_a
            n_r = R
        else:
            n_a = int(torch.randint(6, 21, (1,), generator=g))
            n_b = int(torch.randint(10, 41, (1,), generator=g))
            n_a = min(n_a, max(2, T - 4))
            n_b = max(1, min(n_b, T - 3 - n_a))
            n_r = int(torch.randint(min(10, R), R + 1, (1,), generator=g))
        toks = torch.randint(lo, V, (n_a + n_b,), generator=g)
        row = [CLS] + toks[:n_a].tolist() + [SEP] + toks[n_a:].tolist() + [SEP]
        mp = 1 + int(torch.randint(0, n_a, (1,), generator=g))
        row[mp] = MASK
        n = len(row)
        ids[b, :n] = torch.tensor(row)
        seg[b, n_a + 2:n] = 1
        mask[b, :n] = 1
        mask[b, T:T + n_r] = 1
        mask_pos[b] = mp
        f = torch.relu(torch.randn(n_r, F, generator=g))
        if F >= 6:
            xy = torch.rand(n_r, 4, generator=g).sort(dim=1).values  # x1<=y.. not needed exactly; keep in [0,1]
            x1, x2 = xy[:, 0], xy[:, 2]
            y1, y2 = xy[:, 1], xy[:, 3]
            f[:, F - 6:] = torch.stack([x1, y1, x2, y2, x2 - x1, y2 - y1], 1)
        feats[b, :n_r] = f
    return dict(input_ids=ids, token_type_ids=seg, attention_mask=mask, img_feats=feats, mask_pos=mask_pos)


def synth_vocab_ids(cfg, K, seed=88):
    """K distinct vocabulary ids standing in for colour words / GQA answer first-pieces."""
    g = torch.Generator().manual_seed(seed + 7)
    lo = min(1000, cfg.vocab_size // 2)
    return (torch.randperm(cfg.vocab_size - lo, generator=g)[:K] + lo).sort().values
