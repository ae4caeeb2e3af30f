"""Prints how much of the 1e-3 parity budget each golden case uses on this GPU (tests assert; this reports)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.modeling_bert import BertImgForPreTraining  # noqa: E402
from cpt_b200.modeling_rec import REC_MLM_CPT  # noqa: E402
from cpt_b200.modeling_vcr import NSPCPT  # noqa: E402
from cpt_b200.synthetic import synth_batch, synth_state_dict, synth_vocab_ids  # noqa: E402

for name in ("tiny_s120", "base_s120", "base_s210"):
    g = torch.load(os.path.join("tests", "golden", name + ".pt"))
    d = dict(g["cfg"])
    v = d.pop("vocab_size")
    cfg = C.BertConfig(v, **d)
    sd = synth_state_dict(cfg, seed=g["seed"])
    b = {k: t.cuda() for k, t in synth_batch(cfg, g["B"], g["T"], g["R"], seed=g["seed"]).items()}
    vids = synth_vocab_ids(cfg, g["K"], seed=g["seed"]).cuda()
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(sd, strict=False)
    pre.tie_weights()
    pre = pre.cuda().eval()
    rec, nsp = REC_MLM_CPT(cfg), NSPCPT(cfg)
    rec.copy_from_pretraining_model(pre)
    nsp.copy_from_pretraining_model(pre)
    with torch.no_grad():
        seq, pooled = rec.eval().bert(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"])[:2]
        lg = rec(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"], mask_pos=b["mask_pos"],
                 vocab_ids=vids)[0]
        ns = nsp.eval()(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"])[0]
    e_seq = (seq.cpu()[:, ::7, ::16] - g["seq_sub"]).abs().max().item() / float(g["seq_abs_max"])
    e_lg = ((lg.cpu() - g["logits"]).abs() / g["max_abs_logit_row"][:, None]).max().item()
    e_po = (pooled.cpu() - g["pooled"]).abs().max().item()
    e_ns = (ns.cpu() - g["nsp"]).abs().max().item() / max(1.0, g["nsp"].abs().max().item())
    print("%-10s seq %.2e  logits(rel row max) %.2e  pooled(abs) %.2e  nsp %.2e" % (name, e_seq, e_lg, e_po, e_ns), flush=True)
