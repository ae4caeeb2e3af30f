"""Small batches: the one-kernel-per-op path vs the dataflow chain kernel (CPT_B200_CHAIN_MIN_ROWS), Oscar-base, S=120.
    python tools/small_batch.py"""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ".")
    from cpt_b200 import config as C
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    from cpt_b200.synthetic import synth_batch, synth_state_dict, synth_vocab_ids
    cfg = C.oscar_base()
    pre = BertImgForPreTraining(cfg)
    pre.load_state_dict(synth_state_dict(cfg, seed=88), strict=False)
    pre.tie_weights()
    rec = REC_MLM_CPT(cfg)
    rec.copy_from_pretraining_model(pre.cuda().eval())
    rec.eval()
    vids = synth_vocab_ids(cfg, 2, seed=3).cuda()
    out = []
    for B in (1, 2, 4, 8, 16):
        b = {k: v.cuda() for k, v in synth_batch(cfg, B, 70, 50, seed=3).items()}
        with torch.no_grad():
            f = lambda: rec(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"],  # noqa: E731
                            mask_pos=b["mask_pos"], vocab_ids=vids)[0]
            for _ in range(5):
                r = f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                f()
            e1.record()
            torch.cuda.synchronize()
        out.append("B=%d %.3f ms (%.4f)" % (B, e0.elapsed_time(e1) / 50, float(r[0, 0])))
    print("CHAIN_MIN_ROWS=%s: %s" % (os.environ.get("CPT_B200_CHAIN_MIN_ROWS", "1024 (default)"), "; ".join(out)))
else:
    for mr in ("", "1"):
        env = dict(os.environ)
        if mr:
            env["CPT_B200_CHAIN_MIN_ROWS"] = mr
        subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
