"""How much of a GEMM launch is fixed cost?  One tile per CTA (148 tiles) with K = 64 .. 3072."""
import sys

import torch

sys.path.insert(0, ".")
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.engine import Engine  # noqa: E402

eng = Engine(C.oscar_base(), "cuda:0")
for cfg, m, n in ((1256, 128 * 37, 1024), (2256, 128 * 74, 512), (1128, 128 * 37, 512)):
    for k in (64, 256, 768, 1536, 3072):
        A = torch.randn(m, k, device="cuda").half()
        W = (torch.randn(n, k, device="cuda") * 0.05).half()
        bias = torch.randn(n, device="cuda")
        for _ in range(5):
            eng.gemm(A, W, bias, None, 0, False, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            eng.gemm(A, W, bias, None, 0, False, cfg)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        print("cfg %d M=%d N=%d K=%4d (1 tile/CTA): %.2f us/launch, %.0f TF" % (cfg, m, n, k, us, 2.0 * m * n * k / us / 1e6),
              flush=True)
# an empty-ish kernel for reference: layernorm of 8 rows
x = torch.randn(8, 768, device="cuda")
g = torch.ones(768, device="cuda")
for _ in range(5):
    eng.layernorm(x, g, g, 1e-12)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    eng.layernorm(x, g, g, 1e-12)
e1.record()
torch.cuda.synchronize()
print("tiny layernorm launch (includes 2 torch.empty per call): %.2f us" % (e0.elapsed_time(e1) * 1e3 / 200))
