"""GPU: the dataflow chain kernel (cpt_b200/csrc/chain_sm100.cuh) — dependent GEMM / LayerNorm stages in ONE persistent
launch, rows handed from stage to stage through readiness counters — against plain PyTorch statements of the same
ops, and the chained encoder forward against the one-kernel-per-op launch sequence and the oracle."""
import math
import os

import pytest
import torch

from cpt_b200 import config as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from cpt_b200.engine import Engine
    e = Engine(C.oscar_base(), "cuda:0")
    yield e
    e.close()


def _gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _rand16(g, *shape, scale=1.0):
    return (torch.randn(*shape, device="cuda", generator=g) * scale).half()


@pytest.mark.parametrize("M,N,K,gelu", [(256, 256, 64, 0), (128, 256, 128, 0), (300, 200, 136, 0), (1000, 2304, 768, 0),
                                        (777, 3072, 768, 1), (7680, 2304, 768, 0), (7680, 3072, 768, 1), (100, 64, 768, 1)])
def test_single_gemm_stage_16bit_out(eng, M, N, K, gelu):
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 7 * K)
    Kp, Np = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    A = torch.zeros(M, Kp, device="cuda", dtype=torch.float16)
    W = torch.zeros(N, Kp, device="cuda", dtype=torch.float16)
    A[:, :K] = _rand16(g, M, K)
    W[:, :K] = _rand16(g, N, K, scale=0.05)
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, Np), 7.0, device="cuda", dtype=torch.float16)
    eng.chain([dict(kind="gemm", A=A[:, :K], W=W[:, :K], bias=bias, out=out[:, :N], gelu=gelu)])
    torch.cuda.synchronize()
    ref = A[:, :K].double() @ W[:, :K].double().t() + bias.double()
    if gelu:
        ref = _gelu(ref)
    scale = ref.abs().max().item()
    err = (out[:, :N].double() - ref).abs().max().item()
    assert err <= 1.2e-3 * scale, "max err %.3e (scale %.3e)" % (err, scale)
    if Np > N:
        assert (out[:, N:] == 7.0).all()  # stores are clipped at N


@pytest.mark.parametrize("M,N,K,ksplit", [(256, 768, 768, 1), (777, 768, 3072, 1), (777, 768, 3072, 4), (7680, 768, 3072, 2),
                                          (7680, 768, 768, 1), (130, 128, 512, 1), (3000, 1024, 4096, 1)])
def test_gemm_accumulate_then_layernorm(eng, M, N, K, ksplit):
    """out += A W^T + b (TMA reduce-add into the residual), then LayerNorm of the sum by the same launch."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K + ksplit)
    A, W = _rand16(g, M, K), _rand16(g, N, K, scale=0.05)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5
    beta = torch.randn(N, device="cuda", generator=g) * 0.1
    x = resid.clone()
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    eng.chain([dict(kind="gemm", A=A, W=W, bias=bias, out=x, ksplit=ksplit),
               dict(kind="ln", x=x, gamma=gamma, beta=beta, eps=1e-12, out32=o32, out16=o16, dep=0)])
    torch.cuda.synchronize()
    pre = A.double() @ W.double().t() + bias.double() + resid.double()
    assert (x.double() - pre).abs().max().item() <= 3e-5 * pre.abs().max().item()
    ref = torch.nn.functional.layer_norm(pre, (N,), gamma.double(), beta.double(), 1e-12)
    assert (o32.double() - ref).abs().max().item() <= 2e-4
    assert (o16.double() - ref).abs().max().item() <= 1.5e-3 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(256, 768, 768), (777, 768, 3072), (7680, 768, 768), (7680, 768, 3072), (130, 128, 512),
                                   (3000, 1024, 1024), (100, 256, 64), (513, 512, 200)])
def test_dense_residual_layernorm_in_the_epilogue(eng, M, N, K):
    """LayerNorm(A W^T + b + resid) computed inside the GEMM tile epilogues: row statistics are exchanged between the N
    tiles of a row (different CTA pairs) through L2."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    Kp = (K + 7) // 8 * 8
    A = torch.zeros(M, Kp, device="cuda", dtype=torch.float16)
    W = torch.zeros(N, Kp, device="cuda", dtype=torch.float16)
    A[:, :K] = _rand16(g, M, K)
    W[:, :K] = _rand16(g, N, K, scale=0.05)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) * 2 + 0.5   # a non-zero row mean
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5
    beta = torch.randn(N, device="cuda", generator=g) * 0.1
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    keep = resid.clone()
    eng.chain([dict(kind="gemm", A=A[:, :K], W=W[:, :K], bias=bias, resid=resid, gamma=gamma, beta=beta, eps=1e-12,
                    out32=o32, out16=o16)])
    torch.cuda.synchronize()
    assert torch.equal(resid, keep)   # the residual is read, never written
    pre = A[:, :K].double() @ W[:, :K].double().t() + bias.double() + resid.double()
    ref = torch.nn.functional.layer_norm(pre, (N,), gamma.double(), beta.double(), 1e-12)
    assert (o32.double() - ref).abs().max().item() <= 2e-4
    assert (o16.double() - ref).abs().max().item() <= 1.5e-3 * ref.abs().max().item()
    # only one of the two outputs
    o32b = torch.empty_like(o32)
    eng.chain([dict(kind="gemm", A=A[:, :K], W=W[:, :K], bias=bias, resid=resid, gamma=gamma, beta=beta, eps=1e-12,
                    out32=o32b)])
    torch.cuda.synchronize()
    assert torch.equal(o32b, o32)


def test_fused_layernorm_statistics_are_robust_to_a_large_row_mean(eng):
    """mean >> spread: the (mean, M2) partials are merged pairwise, not formed as E[x^2] - E[x]^2."""
    M, N, K = 300, 768, 64
    g = torch.Generator(device="cuda").manual_seed(1)
    A, W = _rand16(g, M, K, scale=0.01), _rand16(g, N, K, scale=0.01)
    resid = torch.randn(M, N, device="cuda", generator=g) * 0.05 + 300.0
    gamma, beta = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    o32 = torch.empty(M, N, device="cuda")
    eng.chain([dict(kind="gemm", A=A, W=W, resid=resid, gamma=gamma, beta=beta, eps=1e-12, out32=o32)])
    torch.cuda.synchronize()
    pre = A.double() @ W.double().t() + resid.double()
    ref = torch.nn.functional.layer_norm(pre, (N,), gamma.double(), beta.double(), 1e-12)
    assert (o32.double() - ref).abs().max().item() <= 5e-3   # fp32 rounding of x itself at |x| = 300 is 3e-5 / 0.05


@pytest.mark.parametrize("M", [256, 1000, 7680])
def test_full_layer_chain_matches_unfused_kernels(eng, M):
    """AO -> LN -> UP(GELU) -> DOWN -> LN -> QKV' as one launch == the same ops as six launches."""
    H, I = 768, 3072
    g = torch.Generator(device="cuda").manual_seed(M)
    ctx = _rand16(g, M, H)
    h32 = torch.randn(M, H, device="cuda", generator=g)
    Wao, Wi, Wo, Wq = (_rand16(g, H, H, scale=0.03), _rand16(g, I, H, scale=0.03), _rand16(g, H, I, scale=0.02),
                       _rand16(g, 3 * H, H, scale=0.03))
    bao, bi, bo, bq = (torch.randn(n, device="cuda", generator=g) * 0.1 for n in (H, I, H, 3 * H))
    g1, b1, g2, b2 = (torch.rand(H, device="cuda", generator=g) + 0.5 for _ in range(4))
    # reference: the round-1 kernels, one launch each
    x1 = eng.gemm(ctx, Wao, bias=bao, resid=h32, epi=2, out_fp32=True)
    a32, a16 = eng.layernorm(x1, g1, b1, 1e-12)
    inter = eng.gemm(a16, Wi, bias=bi, epi=1)
    x2 = eng.gemm(inter, Wo, bias=bo, resid=a32, epi=2, out_fp32=True)
    o32, o16 = eng.layernorm(x2, g2, b2, 1e-12)
    qkv = eng.gemm(o16, Wq, bias=bq)
    # chain
    c_h32 = h32.clone()
    c_a32, c_o32 = torch.empty_like(h32), torch.empty_like(h32)
    c_a16 = torch.empty(M, H, device="cuda", dtype=torch.float16)
    c_o16 = torch.empty_like(c_a16)
    c_inter = torch.empty(M, I, device="cuda", dtype=torch.float16)
    c_qkv = torch.empty(M, 3 * H, device="cuda", dtype=torch.float16)
    eng.chain([dict(kind="gemm", A=ctx, W=Wao, bias=bao, out=c_h32),
               dict(kind="ln", x=c_h32, gamma=g1, beta=b1, eps=1e-12, out32=c_a32, out16=c_a16, dep=0),
               dict(kind="gemm", A=c_a16, W=Wi, bias=bi, out=c_inter, gelu=1, dep=1),
               dict(kind="gemm", A=c_inter, W=Wo, bias=bo, out=c_a32, dep=2),
               dict(kind="ln", x=c_a32, gamma=g2, beta=b2, eps=1e-12, out32=c_o32, out16=c_o16, dep=3),
               dict(kind="gemm", A=c_o16, W=Wq, bias=bq, out=c_qkv, dep=4)])
    torch.cuda.synchronize()
    assert (c_h32 - x1).abs().max().item() <= 1e-4 * x1.abs().max().item()
    assert (c_inter.float() - inter.float()).abs().max().item() <= 2e-3 * inter.float().abs().max().item()
    assert (c_o32 - o32).abs().max().item() <= 5e-3  # 16-bit rounding flips of the intermediate feed through
    assert (c_qkv.float() - qkv.float()).abs().max().item() <= 1e-2 * qkv.float().abs().max().item()
    # and run it again on the same buffers a few times: counters are re-zeroed per call, results identical
    first = c_qkv.clone()
    for _ in range(3):
        c_h32.copy_(h32)
        eng.chain([dict(kind="gemm", A=ctx, W=Wao, bias=bao, out=c_h32),
                   dict(kind="ln", x=c_h32, gamma=g1, beta=b1, eps=1e-12, out32=c_a32, out16=c_a16, dep=0),
                   dict(kind="gemm", A=c_a16, W=Wi, bias=bi, out=c_inter, gelu=1, dep=1),
                   dict(kind="gemm", A=c_inter, W=Wo, bias=bo, out=c_a32, dep=2),
                   dict(kind="ln", x=c_a32, gamma=g2, beta=b2, eps=1e-12, out32=c_o32, out16=c_o16, dep=3),
                   dict(kind="gemm", A=c_o16, W=Wq, bias=bq, out=c_qkv, dep=4)])
    torch.cuda.synchronize()
    assert torch.equal(first, c_qkv)
    # the same layer with the LayerNorms inside the dense epilogues: 4 stages, the residual buffers are only read
    f_a32, f_o32 = torch.empty_like(h32), torch.empty_like(h32)
    f_a16, f_o16 = torch.empty_like(c_a16), torch.empty_like(c_a16)
    f_qkv = torch.empty_like(c_qkv)
    eng.chain([dict(kind="gemm", A=ctx, W=Wao, bias=bao, resid=h32, gamma=g1, beta=b1, eps=1e-12, out32=f_a32, out16=f_a16),
               dict(kind="gemm", A=f_a16, W=Wi, bias=bi, out=c_inter, gelu=1, dep=0),
               dict(kind="gemm", A=c_inter, W=Wo, bias=bo, resid=f_a32, gamma=g2, beta=b2, eps=1e-12, out32=f_o32,
                    out16=f_o16, dep=1),
               dict(kind="gemm", A=f_o16, W=Wq, bias=bq, out=f_qkv, dep=2)])
    torch.cuda.synchronize()
    assert (f_a32 - a32).abs().max().item() <= 2e-4
    assert (f_o32 - o32).abs().max().item() <= 5e-3
    assert (f_qkv.float() - qkv.float()).abs().max().item() <= 1e-2 * qkv.float().abs().max().item()


def _fold(W, gamma, beta, bias):
    """Host statement of fold_weight_kernel (cpt_b200/csrc/rowwise.cuh): LN(x) W^T + b = rstd (x (gamma .* W)^T - mu g) + c."""
    Wf = (W.float() * gamma[None, :]).half()
    return Wf, Wf.float().sum(1).contiguous(), (W.float() @ beta + bias).contiguous()


def _part(M, N):
    m_pad = ((M + 127) // 128 + 1) // 2 * 2 * 128
    return torch.zeros(2 * ((N + 255) // 256), m_pad, 2, device="cuda")


@pytest.mark.parametrize("M", [256, 1000, 7680])
def test_layer_chain_with_deferred_layernorm(eng, M):
    """The production form: no LayerNorm pass at all.  Dense outputs are written pre-LayerNorm with per-row statistics;
    the consumer GEMMs read the raw rows with gamma folded into their weights and finish the normalisation in their
    epilogues; the residual of the second dense is normalised on the fly."""
    H, I = 768, 3072
    g = torch.Generator(device="cuda").manual_seed(M + 1)
    ctx = _rand16(g, M, H)
    h32 = torch.randn(M, H, device="cuda", generator=g)
    Wao, Wi, Wo, Wq = (_rand16(g, H, H, scale=0.03), _rand16(g, I, H, scale=0.03), _rand16(g, H, I, scale=0.02),
                       _rand16(g, 3 * H, H, scale=0.03))
    bao, bi, bo, bq = (torch.randn(n, device="cuda", generator=g) * 0.1 for n in (H, I, H, 3 * H))
    g1, g2 = (torch.rand(H, device="cuda", generator=g) + 0.5 for _ in range(2))
    b1, b2 = (torch.randn(H, device="cuda", generator=g) * 0.2 for _ in range(2))
    # reference in double precision on the same 16-bit operands
    x1 = ctx.double() @ Wao.double().t() + bao.double() + h32.double()
    a = torch.nn.functional.layer_norm(x1, (H,), g1.double(), b1.double(), 1e-12)
    inter = _gelu(a @ Wi.double().t() + bi.double())
    x2 = inter @ Wo.double().t() + bo.double() + a
    o = torch.nn.functional.layer_norm(x2, (H,), g2.double(), b2.double(), 1e-12)
    qkv = o @ Wq.double().t() + bq.double()
    Wi_f, gi, ci = _fold(Wi, g1, b1, bi)
    Wq_f, gq, cq = _fold(Wq, g2, b2, bq)
    P1, P2 = _part(M, H), _part(M, H)
    a32, x2_32 = torch.empty_like(h32), torch.empty_like(h32)
    a16 = torch.empty(M, H, device="cuda", dtype=torch.float16)
    x2_16 = torch.empty_like(a16)
    c_inter = torch.empty(M, I, device="cuda", dtype=torch.float16)
    c_qkv = torch.empty(M, 3 * H, device="cuda", dtype=torch.float16)
    for _ in range(2):
        eng.chain([dict(kind="gemm", A=ctx, W=Wao, bias=bao, resid=h32, out32=a32, out16=a16, part=P1),
                   dict(kind="gemm", A=a16, W=Wi_f, bias=ci, gvec=gi, apart=P1, eps=1e-12, out=c_inter, gelu=1, dep=0),
                   dict(kind="gemm", A=c_inter, W=Wo, bias=bo, resid=a32, rpart=P1, gamma=g1, beta=b1, eps=1e-12,
                        out32=x2_32, out16=x2_16, part=P2, dep=1),
                   dict(kind="gemm", A=x2_16, W=Wq_f, bias=cq, gvec=gq, apart=P2, eps=1e-12, out=c_qkv, dep=2)])
    torch.cuda.synchronize()
    assert (a32.double() - x1).abs().max().item() <= 3e-5 * x1.abs().max().item()
    assert (c_inter.double() - inter).abs().max().item() <= 4e-3 * inter.abs().max().item()
    got_o = torch.nn.functional.layer_norm(x2_32.double(), (H,), g2.double(), b2.double(), 1e-12)
    assert (got_o - o).abs().max().item() <= 1e-2           # fp16 rounding of inter / a16 feeds through K = 3072
    assert (c_qkv.double() - qkv).abs().max().item() <= 1.2e-2 * qkv.abs().max().item()
    # the partials describe the rows: merged mean / variance against torch
    # (the production kernel writes one plane per 256-column tile, the general one per 128-column half)
    pw = 256 if os.environ.get("CPT_B200_CHAIN_LEAN", "1") != "0" else 128
    n_sl = H // pw
    means, m2s = P2[:n_sl, :M, 0].double(), P2[:n_sl, :M, 1].double()
    mean = means.mean(0)
    var = (m2s.sum(0) + pw * ((means - mean[None]) ** 2).sum(0)) / H
    assert (mean - x2_32.double().mean(1)).abs().max().item() <= 1e-4
    assert (var - x2_32.double().var(1, unbiased=False)).abs().max().item() <= 1e-3 * var.max().item()


def _models(cfg, sd, chain):
    from cpt_b200.modeling_bert import BertImgForPreTraining
    from cpt_b200.modeling_rec import REC_MLM_CPT
    os.environ["CPT_B200_CHAIN"] = "1" if chain else "0"
    os.environ["CPT_B200_CHAIN_FUSE_LN"] = {"tasks": "0", "epilogue": "1"}.get(chain, "2")
    os.environ["CPT_B200_CHAIN_MIN_ROWS"] = "1"
    try:
        pre = BertImgForPreTraining(cfg)
        pre.load_state_dict(sd, strict=False)
        pre.tie_weights()
        pre = pre.to("cuda:0").eval()
        rec = REC_MLM_CPT(cfg)
        rec.copy_from_pretraining_model(pre)
        rec.eval()
        rec.bert.engine()  # the handle reads the environment when it is created
    finally:
        os.environ.pop("CPT_B200_CHAIN", None)
        os.environ.pop("CPT_B200_CHAIN_FUSE_LN", None)
        os.environ.pop("CPT_B200_CHAIN_MIN_ROWS", None)
    return rec


@pytest.mark.parametrize("geom,B,T,R", [("tiny", 4, 70, 50), ("tiny", 3, 30, 10), ("base", 8, 70, 50), ("base", 3, 165, 45)])
def test_chained_encoder_matches_unfused_and_oracle(geom, B, T, R):
    from cpt_b200.synthetic import synth_batch, synth_state_dict, synth_vocab_ids
    from oracle import cpt_oracle as O
    cfg = C.oscar_tiny() if geom == "tiny" else C.oscar_base()
    sd = synth_state_dict(cfg, seed=88)
    b = synth_batch(cfg, B, T, R, seed=5)
    vids = synth_vocab_ids(cfg, 7, seed=88)
    d = {k: v.to("cuda:0") for k, v in b.items()}
    outs = []
    for chain in (True, False, "tasks", "epilogue"):
        rec = _models(cfg, sd, chain)
        with torch.no_grad():
            seq = rec.bert(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"])[0]
            logits = rec(d["input_ids"], d["token_type_ids"], d["attention_mask"], img_feats=d["img_feats"],
                         mask_pos=d["mask_pos"], vocab_ids=vids.to("cuda:0"))[0]
            rec.bert.engine().check()
        outs.append((seq.cpu(), logits.cpu()))
        del rec
    (seq_c, log_c), (seq_u, log_u), (seq_t, log_t), (seq_e, log_e) = outs
    for other in (seq_c, seq_t, seq_e):
        assert (other - seq_u).abs().max().item() <= 2e-3 * seq_u.abs().max().item()
    assert not torch.equal(seq_c, seq_e)   # the switches really selected different paths
    with torch.no_grad():
        oseq, _, _ = O.bert_img_model(sd, cfg, b["input_ids"], b["token_type_ids"], b["attention_mask"],
                                      img_feats=b["img_feats"])
        rows = O.lm_head(sd, cfg, oseq[torch.arange(B), b["mask_pos"]])
    assert (seq_c - oseq).abs().max().item() <= 1e-3 * oseq.abs().max().item()
    row_max = rows.abs().max(dim=1, keepdim=True).values
    assert ((log_c - rows[:, vids]).abs() / row_max).max().item() <= 1e-3
