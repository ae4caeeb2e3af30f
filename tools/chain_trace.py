#!/usr/bin/env python
"""Event log of ONE launch of the dataflow chain kernel at the bench shape (Oscar-base, M = B*120 rows): where do the
producer / MMA issuer / epilogue warps of every CTA pair spend their time, stage by stage.
    CPT_B200_CHAIN_TRACE=1 python tools/chain_trace.py [--batch 64] [--ksplit 1] [--json gpurun_out/chain_trace.json]
"""
import argparse
import json
import os
import sys

os.environ["CPT_B200_CHAIN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cpt_b200 import config as C  # noqa: E402
from cpt_b200.engine import Engine  # noqa: E402

GHZ = 1.965


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seq", type=int, default=120)
    ap.add_argument("--ksplit", type=int, default=1)
    ap.add_argument("--json", default="")
    ap.add_argument("--stages", default="ao,ln1,up,down,ln2,qkv")
    a = ap.parse_args()
    eng = Engine(C.oscar_base(), "cuda:0")
    M, H, I = a.batch * a.seq, 768, 3072
    g = torch.Generator(device="cuda").manual_seed(0)
    r16 = lambda *s, sc=1.0: (torch.randn(*s, device="cuda", generator=g) * sc).half()  # noqa: E731
    ctx, h32 = r16(M, H), torch.randn(M, H, device="cuda", generator=g)
    Wao, Wi, Wo, Wq = r16(H, H, sc=0.03), r16(I, H, sc=0.03), r16(H, I, sc=0.02), r16(3 * H, H, sc=0.03)
    bao, bi, bo, bq = (torch.randn(n, device="cuda", generator=g) * 0.1 for n in (H, I, H, 3 * H))
    g1, b1, g2, b2 = (torch.rand(H, device="cuda", generator=g) + 0.5 for _ in range(4))
    a32, o32 = torch.empty_like(h32), torch.empty_like(h32)
    a16 = torch.empty(M, H, device="cuda", dtype=torch.float16)
    o16 = torch.empty_like(a16)
    inter = torch.empty(M, I, device="cuda", dtype=torch.float16)
    qkv = torch.empty(M, 3 * H, device="cuda", dtype=torch.float16)
    want = a.stages.split(",")
    f32 = h32.clone()
    allst = [("ao", dict(kind="gemm", A=ctx, W=Wao, bias=bao, out=h32)),
             ("ln1", dict(kind="ln", x=h32, gamma=g1, beta=b1, eps=1e-12, out32=a32, out16=a16)),
             ("up", dict(kind="gemm", A=a16, W=Wi, bias=bi, out=inter, gelu=1)),
             ("down", dict(kind="gemm", A=inter, W=Wo, bias=bo, out=a32, ksplit=a.ksplit)),
             ("ln2", dict(kind="ln", x=a32, gamma=g2, beta=b2, eps=1e-12, out32=o32, out16=o16)),
             ("qkv", dict(kind="gemm", A=o16, W=Wq, bias=bq, out=qkv)),
             # dense + residual + LayerNorm in the epilogue
             ("aoln", dict(kind="gemm", A=ctx, W=Wao, bias=bao, resid=f32, gamma=g1, beta=b1, eps=1e-12, out32=a32,
                           out16=a16)),
             ("downln", dict(kind="gemm", A=inter, W=Wo, bias=bo, resid=a32, gamma=g2, beta=b2, eps=1e-12, out32=o32,
                             out16=o16))]
    # LayerNorm deferred to the consumers (production)
    def fold(W, gamma, beta, bias):
        Wf = (W.float() * gamma[None, :]).half()
        return Wf, Wf.float().sum(1).contiguous(), (W.float() @ beta + bias).contiguous()
    m_pad = ((M + 127) // 128 + 1) // 2 * 2 * 128
    P1, P2 = torch.zeros(6, m_pad, 2, device="cuda"), torch.zeros(6, m_pad, 2, device="cuda")
    Wi_f, gi, ci = fold(Wi, g1, b1, bi)
    Wq_f, gq, cq = fold(Wq, g2, b2, bq)
    allst += [("aod", dict(kind="gemm", A=ctx, W=Wao, bias=bao, resid=f32, out32=a32, out16=a16, part=P1)),
              ("upd", dict(kind="gemm", A=a16, W=Wi_f, bias=ci, gvec=gi, apart=P1, eps=1e-12, out=inter, gelu=1)),
              ("downd", dict(kind="gemm", A=inter, W=Wo, bias=bo, resid=a32, rpart=P1, gamma=g1, beta=b1, eps=1e-12,
                             out32=o32, out16=o16, part=P2)),
              ("qkvd", dict(kind="gemm", A=o16, W=Wq_f, bias=cq, gvec=gq, apart=P2, eps=1e-12, out=qkv))]
    stages, names = [], []
    table = dict(allst)
    for n in want:   # stages run in the order given on the command line
        if n in table:
            s = dict(table[n])
            s["dep"] = len(stages) - 1 if stages else None
            stages.append(s)
            names.append(n)
    for _ in range(5):
        eng.chain(stages)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.chain(stages)
    e1.record()
    torch.cuda.synchronize()
    print("chain %s: %.1f us per launch (20 back-to-back launches, CUDA events)" % (names, e0.elapsed_time(e1) * 50))
    hdr, ev = eng.chain_trace()
    t0 = min(h[0] for h in hdr)
    us = lambda p, c: (hdr[p][0] - t0) / 1e3 + c / GHZ / 1e3  # noqa: E731
    per = {}
    end_all = 0.0
    for p, lst in enumerate(ev):
        prev_mma_end = None
        for rec in lst:
            code = rec[8]
            if code == 0 and rec[5] == 0:
                continue
            st = code >> 24
            d = per.setdefault(st, dict(n=0, first=1e9, last=0.0, dep=0.0, issue=0.0, mma=0.0, mma_gap=0.0, ewait=0.0,
                                        ework=0.0, publish=0.0))
            d["n"] += 1
            start = us(p, rec[5])
            end = us(p, rec[7]) if rec[7] else start
            d["first"] = min(d["first"], us(p, rec[0]) if rec[0] else start)
            d["last"] = max(d["last"], end)
            end_all = max(end_all, end)
            if not names[st].startswith("ln"):
                d["dep"] += (rec[1] - rec[0]) / GHZ / 1e3
                d["issue"] += (rec[2] - rec[1]) / GHZ / 1e3
                d["mma"] += (rec[4] - rec[3]) / GHZ / 1e3
                if prev_mma_end is not None:
                    d["mma_gap"] += (rec[3] - prev_mma_end) / GHZ / 1e3
                prev_mma_end = rec[4]
                d["ewait"] += (rec[6] - rec[5]) / GHZ / 1e3
                if rec[10]:
                    d["efetch"] = d.get("efetch", 0.0) + (rec[10] - rec[5]) / GHZ / 1e3
                if rec[13]:
                    d["f_total"] = d.get("f_total", 0.0) + (rec[13] - rec[11]) / GHZ / 1e3
                    d["f_ready"] = d.get("f_ready", 0.0) + ((rec[12] - rec[11]) / GHZ / 1e3 if rec[12] else 0.0)
                    d["f_lead"] = d.get("f_lead", 0.0) + (rec[5] - rec[13]) / GHZ / 1e3
                    d["f_planes"] = d.get("f_planes", 0.0) + (rec[14] - (rec[12] or rec[11])) / GHZ / 1e3
                    d["f_vfree"] = d.get("f_vfree", 0.0) + (rec[15] - rec[14]) / GHZ / 1e3
                    d["f_vec"] = d.get("f_vec", 0.0) + (rec[13] - rec[15]) / GHZ / 1e3
                d["ework"] += (rec[9] - rec[6]) / GHZ / 1e3
                d["publish"] += (rec[7] - rec[9]) / GHZ / 1e3
            else:  # LayerNorm task: 6 rows ready | 1 rows normalised and stored | 2 fenced | 7 published
                d["ewait"] += (rec[6] - rec[5]) / GHZ / 1e3 if rec[6] else 0.0
                d["ework"] += (rec[1] - rec[6]) / GHZ / 1e3 if rec[6] else 0.0
                d["publish"] += (rec[7] - rec[1]) / GHZ / 1e3 if rec[6] else 0.0
                d["mma"] += (rec[2] - rec[1]) / GHZ / 1e3 if rec[6] else 0.0   # = the __threadfence alone
                d["dep"] += rec[0] / GHZ / 1e3      # LayerNorm: row 0 loaded (since the call)
                d["issue"] += rec[3] / GHZ / 1e3    # ... its statistics done
                d["mma_gap"] += rec[4] / GHZ / 1e3  # ... its outputs stored
    print("kernel span by the log: %.1f us; %d pairs" % (end_all, len(ev)))
    print("%-5s %5s %8s %8s | per task (us): %7s %7s %7s %8s %7s %7s %7s" %
          ("stage", "tasks", "first", "last", "depwait", "issue", "mma", "mma_gap", "e.wait", "e.work", "publish"))
    for st in sorted(per):
        d = per[st]
        n = d["n"]
        print("%-5s %5d %8.1f %8.1f |                %7.2f %7.2f %7.2f %8.2f %7.2f %7.2f %7.2f | of e.wait: vectors+statistics %.2f | fetch warp: whole %.2f (rows known published %.2f, planes %.2f, vectors' buffer free %.2f, vectors %.2f), done %.2f before the epilogue asks" %
              (names[st], n, d["first"], d["last"], d["dep"] / n, d["issue"] / n, d["mma"] / n, d["mma_gap"] / n,
               d["ewait"] / n, d["ework"] / n, d["publish"] / n, d.get("efetch", 0.0) / n, d.get("f_total", 0.0) / n, d.get("f_ready", 0.0) / n,
               d.get("f_planes", 0.0) / n, d.get("f_vfree", 0.0) / n, d.get("f_vec", 0.0) / n, d.get("f_lead", 0.0) / n))
    # timeline of a few pairs
    for p in (0, len(ev) // 2, len(ev) - 1):
        row = []
        for rec in ev[p]:
            if rec[8] == 0 and rec[5] == 0:
                continue
            row.append("%s%d[%.0f-%.0f]" % (names[rec[8] >> 24], rec[8] & 0xFFFFFF, us(p, rec[5]), us(p, rec[7] or rec[5])))
        print("pair %d epilogue timeline: %s" % (p, " ".join(row)))
    shown = 0
    for p, lst in enumerate(ev):
        for rec in lst:
            if (rec[8] or rec[5]) and names[rec[8] >> 24].startswith("ln") and shown < 6:
                shown += 1
                print("LN task raw (us): row0 loaded %.2f | stats %.2f | stored %.2f | whole call %.2f | fence %.2f | "
                      "publish %.2f" % (rec[0] / GHZ / 1e3, rec[3] / GHZ / 1e3, rec[4] / GHZ / 1e3,
                                        (rec[1] - rec[6]) / GHZ / 1e3, (rec[2] - rec[1]) / GHZ / 1e3,
                                        (rec[7] - rec[2]) / GHZ / 1e3))
    shown = 0
    for p, lst in enumerate(ev):
        for rec in lst:
            if (rec[8] or rec[5]) and rec[13] and shown < 8:
                shown += 1
                print("fused tile %s raw (us): waited for the accumulator %.2f | pass 1 %.2f | row statistics %.2f | "
                      "pass 2 %.2f | released+published %.2f" %
                      (names[rec[8] >> 24], (rec[10] - rec[5]) / GHZ / 1e3, (rec[11] - rec[10]) / GHZ / 1e3,
                       (rec[12] - rec[11]) / GHZ / 1e3, (rec[13] - rec[12]) / GHZ / 1e3, (rec[7] - rec[13]) / GHZ / 1e3))
    if a.json:
        json.dump(dict(names=names, hdr=hdr, ev=ev), open(a.json, "w"))


if __name__ == "__main__":
    main()
