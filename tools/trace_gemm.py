"""Per-role cycle breakdown + launch timeline of one GEMM launch (CPT_B200_TRACE=1)."""
import os
import sys

os.environ["CPT_B200_TRACE"] = "1"
import torch  # noqa: E402

sys.path.insert(0, ".")
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.engine import Engine  # noqa: E402

eng = Engine(C.oscar_base(), "cuda:0")
M = 7680
cases = [("qkv", M, 2304, 768, 0, False, [1256, 2256]), ("ffn_up", M, 3072, 768, 1, False, [2256])]
for name, m, n, k, epi, f32, cfgs in cases:
    A = torch.randn(m, k, device="cuda").half()
    W = (torch.randn(n, k, device="cuda") * 0.05).half()
    bias = torch.randn(n, device="cuda")
    for cfg in cfgs:
        for _ in range(3):
            eng.gemm(A, W, bias, None, epi, f32, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.gemm(A, W, bias, None, epi, f32, cfg)
        e1.record()
        torch.cuda.synchronize()
        tr = [t for t in eng.gemm_trace() if t[10] > 0]
        av = lambda i: sum(t[i] for t in tr) / len(tr)  # noqa: E731
        g0, g1 = min(t[11] for t in tr), max(t[12] for t in tr)
        last_entry = max(t[11] for t in tr)
        first_exit = min(t[12] for t in tr)
        life_ns = sum(t[12] - t[11] for t in tr) / len(tr)
        print("%-8s cfg %d: event %.1f us | first CTA entry -> last CTA exit %.1f us | entry skew %.1f us, exit skew %.1f us | "
              "CTA life %.1f us = %.0f cyc (%.2f GHz) | prologue %.0f cyc, +pdl wait %.0f | mma role %.0f cyc | epi role %.0f cyc "
              "(wait mma %.0f; per-warp phases: tmem wait %.0f, math+sts %.0f, readback+stg %.0f)"
              % (name, cfg, e0.elapsed_time(e1) * 1e3, (g1 - g0) / 1e3, (last_entry - g0) / 1e3, (g1 - first_exit) / 1e3,
                 life_ns / 1e3, av(10), av(10) / life_ns, av(8), av(9), av(2), av(5), av(6), av(13), av(14), av(15)), flush=True)
