"""Large-K GEMM: does the CTA-pair (cta_group::2) variant beat the single-CTA one when fixed per-tile costs vanish?"""
import sys

import torch

sys.path.insert(0, ".")
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.engine import Engine  # noqa: E402

eng = Engine(C.oscar_base(), "cuda:0")
for (m, n, k) in ((8192, 8192, 8192), (7680, 3072, 3072), (7680, 2304, 768)):
    A = torch.randn(m, k, device="cuda").half()
    W = (torch.randn(n, k, device="cuda") * 0.05).half()
    bias = torch.randn(n, device="cuda")
    for cfg in (1128, 1256, 2128, 2256):
        for _ in range(3):
            eng.gemm(A, W, bias, None, 0, False, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.gemm(A, W, bias, None, 0, False, cfg)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 10
        print("M=%d N=%d K=%d cfg %d: %.1f us  %.0f TF" % (m, n, k, cfg, us, 2.0 * m * n * k / us / 1e6), flush=True)
    ref = torch.matmul(A, W.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        torch.matmul(A, W.t())
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 10
    print("M=%d N=%d K=%d cuBLAS (torch.matmul fp16): %.1f us  %.0f TF" % (m, n, k, us, 2.0 * m * n * k / us / 1e6), flush=True)
