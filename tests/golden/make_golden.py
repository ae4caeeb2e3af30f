"""Generate golden vectors by running the REFERENCE'S OWN files (unmodified, from
/root/reference/Oscar) on CPU fp32, through oracle/ref_shim.py.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/<case>.pt (small tensors; weights/inputs are regenerated from seeds by
cpt_b200.synthetic, so only outputs are stored).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import ref_shim  # noqa: E402
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.synthetic import synth_state_dict, synth_batch, synth_vocab_ids  # noqa: E402

CASES = {
    # name: (config factory, kwargs, B, T, R, K)
    "tiny_s120": (C.oscar_tiny, {}, 3, 70, 50, 5),
    "tiny_noimgln_s40": (C.oscar_tiny, {"use_img_layernorm": 0, "num_contrast_classes": 2}, 2, 24, 16, 3),
    "base_s120": (C.oscar_base, {}, 2, 70, 50, 8),
    "base_s210": (C.oscar_base, {}, 2, 165, 45, 8),
}


def ref_config(cfg):
    d = cfg.to_dict()
    v = d.pop("vocab_size")
    return ref_shim.BertConfig(v, **d)


def build_reference_models(cfg, sd):
    from oscar.modeling.modeling_bert import BertImgForPreTraining
    from oscar.modeling.modeling_rec import REC_MLM_CPT
    from oscar.modeling.modeling_vcr import NSPCPT
    rcfg = ref_config(cfg)
    pre = BertImgForPreTraining(rcfg)
    missing, unexpected = pre.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    pre.tie_weights()
    assert pre.cls.predictions.decoder.weight.data_ptr() == pre.bert.embeddings.word_embeddings.weight.data_ptr()
    rec = REC_MLM_CPT(rcfg)
    rec.copy_from_pretraining_model(pre)
    nsp = NSPCPT(rcfg)
    nsp.copy_from_pretraining_model(pre)
    return pre.eval(), rec.eval(), nsp.eval()


def run_case(name):
    fac, kw, B, T, R, K = CASES[name]
    cfg = fac(**kw)
    sd = synth_state_dict(cfg, seed=88)
    batch = synth_batch(cfg, B, T, R, seed=88)
    vocab_ids = synth_vocab_ids(cfg, K, seed=88)
    pre, rec, nsp = build_reference_models(cfg, sd)
    ids, seg, mask, feats, mp = (batch[k] for k in ("input_ids", "token_type_ids", "attention_mask",
                                                    "img_feats", "mask_pos"))
    out = {"case": name, "B": B, "T": T, "R": R, "K": K, "seed": 88, "cfg": cfg.to_dict(),
           "state_dict_keys": sorted(pre.state_dict().keys())}
    with torch.no_grad():
        seq, pooled = rec.bert(ids, seg, mask, img_feats=feats)[:2]   # positional (ids, segment, mask)
        scores = rec(ids, seg, mask, img_feats=feats)[0]              # [B,S,V]
        rows = scores[torch.arange(B), mp]                            # zeroshot/refcoco_cpt.py:219
        out["logits"] = rows[:, vocab_ids].clone()                    # [B,K]
        out["max_abs_logit_row"] = rows.abs().max(dim=1).values
        out["nsp"] = nsp(ids, seg, mask, img_feats=feats)[0].clone()
        out["pooled"] = pooled.clone()
        out["seq_sub"] = seq[:, ::7, ::16].clone()
        out["seq_sum"] = seq.double().sum(dim=-1).float()
        out["seq_abs_max"] = seq.abs().max()
        out["scores_sub"] = scores[:, ::13, ::509].clone()
        if name.startswith("tiny"):
            out["seq"] = seq.clone()
            out["rows"] = rows.clone()
    if name.startswith("tiny"):
        # training-mode parity case: dropout forced to 0, loss + a few grads (few-shot path,
        # fewshot/refcoco_cpt.py:231-250 / gqa_cpt.py:428-453)
        for m in (rec,):
            m.train()
            for mod in m.modules():
                if isinstance(mod, torch.nn.Dropout):
                    mod.p = 0.0
        labels = torch.full((B, T + R), -1, dtype=torch.long)
        labels[torch.arange(B), mp] = vocab_ids[torch.arange(B) % K]
        loss = rec(ids, seg, mask, masked_lm_labels=labels, img_feats=feats)[0]
        loss.backward()
        out["loss"] = loss.detach().clone()
        named = dict(rec.named_parameters())
        for k in ("bert.encoder.layer.0.attention.self.query.weight", "bert.encoder.layer.1.output.dense.bias",
                  "bert.img_embedding.weight", "cls.transform.dense.weight",
                  "bert.embeddings.LayerNorm.weight", "bert.encoder.layer.0.intermediate.dense.weight"):
            g = named[k].grad
            out["grad:" + k] = (g if g.numel() <= 70000 else g.flatten()[::17]).clone()
        out["grad_none"] = sorted(k for k, p in named.items() if p.grad is None)
        wg = named["bert.embeddings.word_embeddings.weight"].grad
        out["grad_word_rowsum"] = wg.double().sum(1).float()
        # the VCR few-shot step (vcr_nsp_cpt.py:434-461): NSPCPT with next_sentence_label, dropout forced to 0.
        # rec and nsp share `bert`: clear the MLM gradients first
        rec.zero_grad()
        nsp.train()
        for mod in nsp.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        nsp_labels = torch.arange(B) % cfg.num_contrast_classes
        if B > 2:
            nsp_labels[1] = -1  # ignore_index
        nloss = nsp(ids, seg, mask, next_sentence_label=nsp_labels, img_feats=feats)[0]
        nloss.backward()
        out["nsp_labels"] = nsp_labels.clone()
        out["nsp_loss"] = nloss.detach().clone()
        nnamed = dict(nsp.named_parameters())
        for k in ("cls.weight", "cls.bias", "bert.pooler.dense.weight", "bert.encoder.layer.1.attention.self.value.weight",
                  "bert.encoder.layer.0.output.LayerNorm.bias", "bert.embeddings.position_embeddings.weight"):
            out["nsp_grad:" + k] = nnamed[k].grad.clone()
    path = os.path.join(HERE, name + ".pt")
    torch.save(out, path)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    assert ref_shim.reference_available(), "needs /root/reference (build container only)"
    ref_shim.install()
    torch.manual_seed(88)
    for n in (sys.argv[1:] or CASES):
        run_case(n)
