#!/usr/bin/env python
"""Max error of every forward variant against the fp32 oracle on Oscar-base (12 layers): hidden states relative to
max|seq_out| and colour logits relative to the row maximum.  One process per variant (the handle reads its switches
from the environment when it is created).
    python tools/path_precision.py [--batch 8] [--T 70] [--R 50]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "unfused (round 1)": {"CPT_B200_CHAIN": "0"},
    "unfused + LayerNorm folded (round 1, CPT_B200_FOLD_LN=1)": {"CPT_B200_CHAIN": "0", "CPT_B200_FOLD_LN": "1"},
    "chain, LayerNorm row tasks": {"CPT_B200_CHAIN": "1", "CPT_B200_CHAIN_FUSE_LN": "0", "CPT_B200_CHAIN_MIN_ROWS": "1"},
    "chain, LayerNorm in the dense epilogue": {"CPT_B200_CHAIN": "1", "CPT_B200_CHAIN_FUSE_LN": "1",
                                               "CPT_B200_CHAIN_MIN_ROWS": "1"},
    "chain, LayerNorm deferred to the consumers": {"CPT_B200_CHAIN": "1", "CPT_B200_CHAIN_FUSE_LN": "2",
                                                   "CPT_B200_CHAIN_MIN_ROWS": "1"},
}


def worker(a):
    import torch
    from cpt_b200 import config as C
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    from cpt_b200.synthetic import synth_batch, synth_state_dict, synth_vocab_ids
    from oracle import cpt_oracle as O
    cfg = C.oscar_large() if a.large else C.oscar_base()
    sd = synth_state_dict(cfg, seed=88)
    b = synth_batch(cfg, a.batch, a.T, a.R, seed=5)
    vids = synth_vocab_ids(cfg, 7, seed=88)
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(sd, strict=False)
    pre.tie_weights()
    rec = REC_MLM_CPT(cfg)
    rec.copy_from_pretraining_model(pre.cuda().eval())
    rec.eval()
    d = {k: v.cuda() for k, v in b.items()}
    with torch.no_grad():
        seq = rec.bert(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"])[0].cpu()
        lg = rec(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                 mask_pos=d["mask_pos"], vocab_ids=vids.cuda())[0].cpu()
        cache = a.cache
        if cache and os.path.exists(cache):
            oseq, rows = torch.load(cache)
        else:
            oseq = O.bert_img_model(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                                    img_feats=b["img_feats"])[0]
            rows = O.lm_head(sd, cfg, oseq[torch.arange(a.batch), b["mask_pos"]])
            if cache:
                torch.save((oseq, rows), cache)
    e_seq = ((seq - oseq).abs().max() / oseq.abs().max()).item()
    e_lg = ((lg - rows[:, vids]).abs() / rows.abs().max(dim=1, keepdim=True).values).max().item()
    print(json.dumps({"seq_out_rel_err": e_seq, "logit_rel_to_row_max_err": e_lg}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--T", type=int, default=70)
    ap.add_argument("--R", type=int, default=50)
    ap.add_argument("--large", action="store_true")
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--cache", default="")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if a.worker:
        return worker(a)
    cache = "/tmp/path_precision_oracle_%d_%d_%d_%d.pt" % (a.batch, a.T, a.R, int(a.large))
    for name, env in VARIANTS.items():
        if a.only and a.only not in name:
            continue
        e = dict(os.environ)
        e.update(env)
        cmd = [sys.executable, os.path.abspath(__file__), "--worker", "--batch", str(a.batch), "--T", str(a.T), "--R",
               str(a.R), "--cache", cache] + (["--large"] if a.large else [])
        r = subprocess.run(cmd, env=e, capture_output=True, text=True)
        out = r.stdout.strip().splitlines()
        print("%-58s %s" % (name, out[-1] if out and r.returncode == 0 else "FAILED: " + (r.stderr.strip().splitlines() or ["?"])[-1]))


if __name__ == "__main__":
    main()
