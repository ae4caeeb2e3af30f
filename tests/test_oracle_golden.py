"""CPU: the oracle restatement (oracle/cpt_oracle.py) against outputs of the reference's own
files (tests/golden/*.pt, made by tests/golden/make_golden.py) and against HF transformers 5.x
eager BERT blocks as an independent second opinion."""
import os

import pytest
import torch

from cpt_b200 import config as C
from cpt_b200.synthetic import synth_state_dict, synth_batch, synth_vocab_ids
from oracle import cpt_oracle as O

TOL = 2e-5  # fp32 vs fp32, different op order (absolute, on O(1..10) values)


def load_case(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name + ".pt"))
    d = dict(g["cfg"])
    v = d.pop("vocab_size")
    cfg = C.BertConfig(v, **d)
    sd = synth_state_dict(cfg, seed=g["seed"])
    batch = synth_batch(cfg, g["B"], g["T"], g["R"], seed=g["seed"])
    vids = synth_vocab_ids(cfg, g["K"], seed=g["seed"])
    return g, cfg, sd, batch, vids


@pytest.mark.parametrize("name", ["tiny_s120", "tiny_noimgln_s40", "base_s120", "base_s210"])
def test_oracle_matches_reference_outputs(golden_dir, name):
    g, cfg, sd, b, vids = load_case(golden_dir, name)
    with torch.no_grad():
        seq, pooled, _ = O.bert_img_model(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                                          img_feats=b["img_feats"])
        logits = O.cpt_mlm_logits(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                                  b["img_feats"], b["mask_pos"], vids)
        nsp = O.nsp_cpt(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                        img_feats=b["img_feats"])[0]
    assert (seq[:, ::7, ::16] - g["seq_sub"]).abs().max() < TOL
    assert (pooled - g["pooled"]).abs().max() < TOL
    assert (logits - g["logits"]).abs().max() < TOL
    assert (nsp - g["nsp"]).abs().max() < TOL
    assert (seq.double().sum(-1).float() - g["seq_sum"]).abs().max() < 1e-3
    if "seq" in g:
        assert (seq - g["seq"]).abs().max() < TOL
        with torch.no_grad():
            scores = O.rec_mlm_cpt(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                                   img_feats=b["img_feats"])[0]
        rows = scores[torch.arange(g["B"]), b["mask_pos"]]
        assert (rows - g["rows"]).abs().max() < TOL
        assert (scores[:, ::13, ::509] - g["scores_sub"]).abs().max() < TOL


@pytest.mark.parametrize("name", ["tiny_s120", "tiny_noimgln_s40"])
def test_oracle_training_loss_and_grads(golden_dir, name):
    g, cfg, sd, b, vids = load_case(golden_dir, name)
    cfg.hidden_dropout_prob = 0.0
    cfg.attention_probs_dropout_prob = 0.0
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "cls.predictions.decoder.weight"}
    B, K = g["B"], g["K"]
    labels = torch.full((B, g["T"] + g["R"]), -1, dtype=torch.long)
    labels[torch.arange(B), b["mask_pos"]] = vids[torch.arange(B) % K]
    loss = O.rec_mlm_cpt(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"], labels,
                         img_feats=b["img_feats"], training=True)[0]
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    ren = {"cls.transform.dense.weight": "cls.predictions.transform.dense.weight"}
    for k, ref in g.items():
        if not k.startswith("grad:"):
            continue
        name_ = ren.get(k[5:], k[5:])
        gr = sd[name_].grad
        gr = gr if gr.numel() <= 70000 else gr.flatten()[::17]
        assert (gr - ref).abs().max() < 1e-6 + 1e-4 * ref.abs().max(), k
    # only the pooler receives no gradient on the MLM path (why DDP needs find_unused_parameters)
    assert g["grad_none"] == ["bert.pooler.dense.bias", "bert.pooler.dense.weight"]
    assert sd["bert.pooler.dense.weight"].grad is None
    wsum = sd["bert.embeddings.word_embeddings.weight"].grad.double().sum(1).float()
    assert (wsum - g["grad_word_rowsum"]).abs().max() < 1e-5


@pytest.mark.parametrize("name", ["tiny_s120", "tiny_noimgln_s40"])
def test_oracle_nsp_training_loss_and_grads(golden_dir, name):
    """NSPCPT with next_sentence_label (the VCR few-shot step) — fixtures from the reference's own modeling_vcr.py."""
    g, cfg, sd, b, vids = load_case(golden_dir, name)
    cfg.hidden_dropout_prob = 0.0
    cfg.attention_probs_dropout_prob = 0.0
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "cls.predictions.decoder.weight"}
    loss = O.nsp_cpt(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                     next_sentence_label=g["nsp_labels"], img_feats=b["img_feats"], training=True)[0]
    loss.backward()
    assert abs(float(loss) - float(g["nsp_loss"])) < 1e-5
    ren = {"cls.weight": "cls.seq_relationship.weight", "cls.bias": "cls.seq_relationship.bias"}
    n = 0
    for k, ref in g.items():
        if not k.startswith("nsp_grad:"):
            continue
        gr = sd[ren.get(k[9:], k[9:])].grad
        assert (gr - ref).abs().max() < 1e-6 + 1e-4 * ref.abs().max(), k
        n += 1
    assert n == 6


def test_state_dict_keys_match_reference(golden_dir):
    g, cfg, sd, _, _ = load_case(golden_dir, "tiny_s120")
    assert sorted(sd.keys()) == g["state_dict_keys"]


def test_oracle_against_hf_blocks():
    """Independent cross-check: HF transformers 5.x eager BertEncoder fed the same concatenated
    embeddings + additive mask must agree with the oracle's encoder."""
    tr = pytest.importorskip("transformers")
    from transformers.models.bert.modeling_bert import BertEncoder
    cfg = C.oscar_tiny()
    sd = synth_state_dict(cfg, 88)
    b = synth_batch(cfg, 2, 24, 16, 88)
    hf_cfg = tr.BertConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size,
                           num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                           intermediate_size=cfg.intermediate_size, layer_norm_eps=cfg.layer_norm_eps,
                           hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf_cfg._attn_implementation = "eager"
    enc = BertEncoder(hf_cfg).eval()
    esd = {k[len("bert.encoder."):]: v for k, v in sd.items() if k.startswith("bert.encoder.")}
    enc.load_state_dict(esd, strict=True)
    with torch.no_grad():
        h0 = torch.cat((O.text_embeddings(sd, cfg, b["input_ids"], b["token_type_ids"]),
                        O.region_embeddings(sd, cfg, b["img_feats"])), 1)
        ext = O.extended_attention_mask(b["attention_mask"])
        out = enc(h0, attention_mask=ext)
        hf_seq = out.last_hidden_state if hasattr(out, "last_hidden_state") else out[0]
        seq, _, _ = O.bert_img_model(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                                     img_feats=b["img_feats"])
    assert (hf_seq - seq).abs().max() < TOL
