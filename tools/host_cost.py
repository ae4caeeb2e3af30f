"""Host (CPU) cost of enqueuing one forward vs its GPU time; and the same forward replayed from a CUDA graph."""
import sys
import time

import torch

sys.path.insert(0, ".")
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.modeling_bert import BertImgForPreTraining  # noqa: E402
from cpt_b200.modeling_rec import REC_MLM_CPT  # noqa: E402
from cpt_b200.synthetic import synth_batch, synth_state_dict, synth_vocab_ids  # noqa: E402

cfg = C.oscar_base()
pre = BertImgForPreTraining(cfg)
pre.load_state_dict(synth_state_dict(cfg, 88), strict=False)
pre.tie_weights()
pre = pre.cuda().eval()
m = REC_MLM_CPT(cfg)
m.copy_from_pretraining_model(pre)
m.eval()
vids = synth_vocab_ids(cfg, 2, 88).cuda()
b = {k: v.cuda() for k, v in synth_batch(cfg, 64, 70, 50, 1).items()}


def step():
    return m(b["input_ids"], b["token_type_ids"], b["attention_mask"], img_feats=b["img_feats"], mask_pos=b["mask_pos"],
             vocab_ids=vids)[0]


with torch.no_grad():
    for _ in range(5):
        step()
    m.bert.freeze_engine_weights(True)
    torch.cuda.synchronize()
    N = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(N):
        step()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("eager: host enqueue %.3f ms/forward, GPU %.3f ms/forward" % ((t1 - t0) * 1e3 / N, e0.elapsed_time(e1) / N))
    # CUDA graph of the same forward
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
        with torch.cuda.graph(g, stream=s):
            out = step()
    torch.cuda.synchronize()
    ref = step().clone()
    g.replay()
    torch.cuda.synchronize()
    print("graph replay matches eager:", torch.equal(out, ref))
    e0.record()
    t0 = time.perf_counter()
    for _ in range(N):
        g.replay()
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    print("graph: host %.3f ms/forward, GPU %.3f ms/forward" % ((t1 - t0) * 1e3 / N, e0.elapsed_time(e1) / N))
