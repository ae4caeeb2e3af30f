"""Diagnostic: NSP training gradients, bf16 vs fp16 (loss-scaled) operands, against the fp32 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cpt_b200 import config as C
from cpt_b200.synthetic import synth_batch, synth_state_dict
from oracle import cpt_oracle as O
from cpt_b200.modeling_bert import BertImgForPreTraining
from cpt_b200.modeling_vcr import NSPCPT

B, T, R = 8, 60, 40
for dtype, scale in (("bf16", 1.0), ("fp16", 1024.0)):
    cfg = C.oscar_tiny(num_hidden_layers=2)
    cfg.hidden_dropout_prob = cfg.attention_probs_dropout_prob = 0.0
    cfg.cpt_b200_train_dtype = dtype
    sd = synth_state_dict(cfg, seed=6)
    b = synth_batch(cfg, B, T, R, seed=40 + B)
    labels = torch.arange(B) % cfg.num_contrast_classes
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref_loss = O.nsp_cpt(leaf, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                         next_sentence_label=labels, img_feats=b["img_feats"])[0]
    ref_loss.backward()
    pre = BertImgForPreTraining(cfg); pre.load_state_dict(sd, strict=False); pre.tie_weights(); pre = pre.to("cuda")
    nsp = NSPCPT(cfg); nsp.copy_from_pretraining_model(pre); nsp.train()
    d = {k: v.cuda() for k, v in b.items()}
    poison = torch.full((1 << 28,), float("nan"), device="cuda")  # 1 GiB of NaN handed back to the caching allocator
    del poison
    loss = nsp(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
               next_sentence_label=labels.cuda())[0]
    (loss * scale).backward()
    print(dtype, "loss", loss.item(), ref_loss.item())
    for k, p in nsp.named_parameters():
        key = k if k.startswith("bert.") else "cls.seq_relationship." + k[len("cls."):]
        r = leaf[key].grad
        if p.grad is None:
            continue
        g = p.grad.cpu() / scale
        e = (g - r).abs().max().item() / max(r.abs().max().item(), 1e-30)
        cos = torch.nn.functional.cosine_similarity(g.flatten().double(), r.flatten().double(), dim=0).item()
        if e > 5e-3:
            print("  %-60s err %.4f cos %.6f refmax %.3e" % (key, e, cos, r.abs().max().item()))
