"""Per-role cycle breakdown of one GEMM launch (CPT_B200_TRACE=1): where do the producer / MMA issuer / epilogue wait?"""
import os
import sys

os.environ["CPT_B200_TRACE"] = "1"
import torch  # noqa: E402

sys.path.insert(0, ".")
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.engine import Engine  # noqa: E402

eng = Engine(C.oscar_base(), "cuda:0")
M = 7680
cases = [("qkv", M, 2304, 768, 0, False, [1128, 1256, 2256]), ("ffn_up", M, 3072, 768, 1, False, [1256, 2256]),
         ("ffn_down", M, 768, 3072, 2, True, [1128, 2128, 2192]), ("attn_out", M, 768, 768, 2, True, [1128, 2192])]
for name, m, n, k, epi, f32, cfgs in cases:
    A = torch.randn(m, k, device="cuda").half()
    W = (torch.randn(n, k, device="cuda") * 0.05).half()
    bias = torch.randn(n, device="cuda")
    res = torch.randn(m, n, device="cuda") if epi == 2 else None
    for cfg in cfgs:
        for _ in range(3):
            eng.gemm(A, W, bias, res, epi, f32, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.gemm(A, W, bias, res, epi, f32, cfg)
        e1.record()
        torch.cuda.synchronize()
        tr = [t for t in eng.gemm_trace() if t[7] > 0 or t[0] > 0]
        mx = lambda i: max(t[i] for t in tr)  # noqa: E731
        av = lambda i: sum(t[i] for t in tr) / len(tr)  # noqa: E731
        print("%-8s cfg %d: %.1f us | ctas %d tiles/cta max %d | producer total %.0f wait_free %.0f | mma total %.0f "
              "wait_operands %.0f wait_acc_drained %.0f | epi total %.0f wait_mma %.0f  (avg cycles; max mma total %.0f)"
              % (name, cfg, e0.elapsed_time(e1) * 1e3, len(tr), mx(7), av(0), av(1), av(2), av(3), av(4), av(5), av(6),
                 mx(2)), flush=True)
