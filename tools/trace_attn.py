"""Per-phase cycle breakdown of the ping-pong attention kernel's softmax group 0 (CPT_B200_TRACE=1)."""
import os
import sys

os.environ["CPT_B200_TRACE"] = "1"
import torch  # noqa: E402

sys.path.insert(0, ".")
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.engine import Engine  # noqa: E402

eng = Engine(C.oscar_base(), "cuda:0")
names = ["mask+bar", "wait S", "tmem ld", "scale+max", "exp+P->smem", "fence+arrive", "wait O", "readout+store"]
for B, S in ((64, 120),):
    qkv = (torch.randn(B * S, 2304, device="cuda") * 1.5).half()
    ext = torch.zeros(B, S, device="cuda")
    for _ in range(3):
        eng.attention(qkv, ext, B, S, 0)
    torch.cuda.synchronize()
    eng.attention(qkv, ext, B, S, 0)
    torch.cuda.synchronize()
    tr = [t for t in eng.gemm_trace() if t[8] > 0]
    items = sum(t[8] for t in tr) / len(tr)
    print("B=%d S=%d: group-0 items per CTA %.1f; cycles per item:" % (B, S, items))
    for i, n in enumerate(names):
        print("   %-14s %7.0f" % (n, sum(t[i] for t in tr) / len(tr) / items))
    print("   total          %7.0f" % (sum(sum(t[:8]) for t in tr) / len(tr) / items))
