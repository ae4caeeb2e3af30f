"""Times the layer GEMM shapes of the CPT path for a list of (block_n, cluster) choices on the current GPU.
    python tools/tune_gemm.py [B] [S]
Each timing: CUDA events around 20 back-to-back launches over 4 rotating operand sets, after 3 warm-ups."""
import sys

import torch

sys.path.insert(0, ".")
from cpt_b200 import config as C  # noqa: E402
from cpt_b200.engine import Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 120
M = B * S
eng = Engine(C.oscar_base(), "cuda:0")
shapes = [("qkv", M, 2304, 768, 0, False), ("attn_out", M, 768, 768, 2, True), ("ffn_up", M, 3072, 768, 1, False),
          ("ffn_down", M, 768, 3072, 2, True), ("img", B * 50, 768, 2056, 0, True)]
cfgs = [1064, 1128, 1192, 1256, 2064, 2128, 2192, 2256]
NSET = 4
for name, m, n, k, epi, f32 in shapes:
    A = [torch.randn(m, k, device="cuda").half() for _ in range(NSET)]
    W = [(torch.randn(n, k, device="cuda") * 0.05).half() for _ in range(NSET)]
    bias = torch.randn(n, device="cuda")
    res = torch.randn(m, n, device="cuda") if epi == 2 else None
    line = []
    for cfg in cfgs:
        try:
            for i in range(3):
                eng.gemm(A[i % NSET], W[i % NSET], bias, res, epi, f32, cfg)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(20):
                eng.gemm(A[i % NSET], W[i % NSET], bias, res, epi, f32, cfg)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 20
            line.append("%d:%.1fus/%.0fTF" % (cfg, us, 2.0 * m * n * k / us / 1e6))
        except Exception as ex:  # noqa: BLE001
            line.append("%d:ERR(%s)" % (cfg, str(ex)[:40]))
    print("%-9s M=%d N=%d K=%d | %s" % (name, m, n, k, "  ".join(line)), flush=True)
