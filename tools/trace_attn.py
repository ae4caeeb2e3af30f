"""Per-role wait breakdown of the pipelined attention kernel (CPT_B200_TRACE=1)."""
import os
import sys

os.environ["CPT_B200_TRACE"] = "1"
import torch  # noqa: E402

sys.path.insert(0, ".")
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.engine import Engine  # noqa: E402

eng = Engine(C.oscar_base(), "cuda:0")
for B, S in ((64, 120), (32, 210)):
    qkv = (torch.randn(B * S, 2304, device="cuda") * 1.5).half()
    ext = torch.zeros(B, S, device="cuda")
    for _ in range(3):
        eng.attention(qkv, ext, B, S, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.attention(qkv, ext, B, S, 0)
    e1.record()
    torch.cuda.synchronize()
    tr = [t for t in eng.gemm_trace() if t[7] > 0]
    av = lambda i: sum(t[i] for t in tr) / len(tr)  # noqa: E731
    print("B=%d S=%d: %.1f us/launch | items/cta %.1f | producer wait_slot %.0f | mma wait_qk %.0f wait_softmax %.0f | "
          "softmax wait_S %.0f wait_O %.0f total %.0f (avg cycles per CTA)"
          % (B, S, e0.elapsed_time(e1) * 100, av(7), av(0), av(1), av(2), av(3), av(4), av(5)), flush=True)
