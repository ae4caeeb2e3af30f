"""GPU: each hand-written sm_100a kernel against a plain PyTorch fp32 statement of the same op, called through
the C ABI (cpt_b200.engine.Engine -> libcpt_b200.so)."""
import math

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


GEMM_CASES = [
    # M, N, K, epi, out_fp32, tile_cfg  (block_n + 1000 * CTAs-per-MMA; 0 = library default)
    (128, 128, 64, 0, True, 1128),
    (128, 128, 256, 0, True, 1128),
    (256, 256, 128, 0, True, 1256),
    (256, 64, 128, 0, True, 1064),
    (300, 200, 136, 0, True, 1128),     # ragged M, N, K (TMA zero-fill + masked epilogue)
    (300, 200, 136, 0, False, 1192),
    (1000, 2304, 768, 0, False, 0),      # QKV projection shape, 16-bit out, default tile
    (1000, 2304, 768, 0, False, 1128),
    (1000, 2304, 768, 0, False, 2256),  # CTA pair, tcgen05 cta_group::2
    (1000, 2304, 768, 0, False, 2128),  
    (1000, 2304, 768, 0, False, 2192),  
    (777, 3072, 768, 1, False, 0),       # FFN up: bias + erf-GELU
    (777, 3072, 768, 1, False, 2256),
    (777, 768, 3072, 2, True, 0),        # FFN down: bias + fp32 residual
    (777, 768, 3072, 2, True, 1064),
    (777, 768, 3072, 2, True, 2064),
    (777, 768, 3072, 2, True, 2128),
    (450, 768, 2054, 0, True, 0),        # region embedding: K = 2054 (tail of 6 in the last 64-wide k-block)
    (130, 1001, 768, 0, True, 0),        # vocabulary-decoder-like: N not a multiple of anything, unaligned rows
    (130, 1001, 768, 0, True, 2256),
    (7680, 768, 768, 2, True, 0),        # many tiles per CTA: exercises the TMEM double buffer + ring wrap
    (7680, 768, 768, 2, True, 2128),
    (7680, 2304, 768, 0, False, 2256),
]


@pytest.mark.parametrize("M,N,K,epi,out_fp32,cfg", GEMM_CASES)
def test_gemm_against_torch(eng, M, N, K, epi, out_fp32, cfg):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    Kp = (K + 7) // 8 * 8
    A = torch.zeros(M, Kp, device="cuda", dtype=torch.float16)
    W = torch.zeros(N, Kp, device="cuda", dtype=torch.float16)
    A[:, :K] = torch.randn(M, K, device="cuda", generator=g).half()
    W[:, :K] = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if epi == 2 else None
    out = eng.gemm(A[:, :K], W[:, :K], bias=bias, resid=resid, epi=epi, out_fp32=out_fp32, tile_cfg=cfg)
    torch.cuda.synchronize()
    ref = A[:, :K].double() @ W[:, :K].double().t() + bias.double()
    if epi == 1:
        ref = _gelu(ref)
    if epi == 2:
        ref = ref + resid.double()
    scale = ref.abs().max().item()
    err = (out.double() - ref).abs().max().item()
    tol = (2e-5 if out_fp32 else 1.2e-3) * scale   # fp32 accumulate / one 16-bit rounding of the output
    assert err <= tol, "max err %.3e (scale %.3e)" % (err, scale)


@pytest.mark.parametrize("B,S", [(2, 120), (3, 210), (1, 40), (2, 128), (2, 129), (1, 256), (2, 1), (5, 17), (64, 120),
                                 (16, 210), (3, 64), (2, 65), (2, 192)])
def test_attention_against_torch_and_simt(eng, B, S):
    H, nH, dH = 768, 12, 64
    g = torch.Generator(device="cuda").manual_seed(S * 31 + B)
    qkv = (torch.randn(B * S, 3 * H, device="cuda", generator=g) * 1.5).half()
    mask = (torch.rand(B, S, device="cuda", generator=g) > 0.3).long()
    mask[:, 0] = 1
    ext = (1.0 - mask.float()) * -10000.0
    ctx_tc = eng.attention(qkv, ext, B, S, impl=0)
    ctx_simt = eng.attention(qkv, ext, B, S, impl=1)
    ctx_tile = eng.attention(qkv, ext, B, S, impl=2)
    ctx_pipe = eng.attention(qkv, ext, B, S, impl=3)
    torch.cuda.synchronize()
    q, k, v = (t.float().view(B, S, nH, dH).permute(0, 2, 1, 3) for t in qkv.split(H, dim=1))
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0 + ext[:, None, None, :], -1)
    ref = (p @ v).permute(0, 2, 1, 3).reshape(B * S, H)
    scale = ref.abs().max().item()
    assert (ctx_simt.float() - ref).abs().max().item() <= 1.5e-3 * scale
    assert (ctx_tc.float() - ref).abs().max().item() <= 3e-3 * scale   # P is rounded to 16 bits before P.V
    assert (ctx_tc.float() - ctx_simt.float()).abs().max().item() <= 3e-3 * scale
    assert (ctx_tile.float() - ref).abs().max().item() <= 3e-3 * scale
    assert (ctx_pipe.float() - ref).abs().max().item() <= 3e-3 * scale


@pytest.mark.parametrize("M,N,K,cfg", [(768, 768, 7680, 0), (768, 3072, 1920, 0), (3072, 768, 840, 1256),
                                       (2304, 768, 75, 1128), (30522, 768, 64, 0), (128, 2056, 200, 1064),
                                       (100, 72, 33, 0)])
def test_gemm_transposed_operands_accumulate(eng, M, N, K, cfg):
    """the weight-gradient form: out[M,N] += A[K,M]^T . W[K,N] with both operands read through MN-major descriptors
    and the fp32 result added into `out` by TMA reduce-add stores"""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    Mp, Np = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    A = torch.randn(K, Mp, device="cuda", generator=g).half()[:, :M]
    W = torch.randn(K, Np, device="cuda", generator=g).half()[:, :N]
    ref = A.float().t() @ W.float()
    out = eng.gemm(A, W, out_fp32=True, tile_cfg=cfg, trans=True)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 2e-3 * scale
    if N % 4 == 0:
        acc = torch.randn(M, N, device="cuda", generator=g)
        want = acc + ref
        eng.gemm(A, W, tile_cfg=cfg, trans=True, accumulate_into=acc)
        torch.cuda.synchronize()
        assert (acc - want).abs().max().item() <= 2e-3 * scale
        # split-K: K cut into pieces that all add into the destination
        acc2 = want - ref
        eng.gemm(A, W, tile_cfg=cfg, trans=True, accumulate_into=acc2, ksplit=7)
        torch.cuda.synchronize()
        assert (acc2 - want).abs().max().item() <= 2e-3 * scale


@pytest.mark.parametrize("M,N,K,cfg", [(1920, 768, 3072, 0), (7680, 3072, 768, 1256), (75, 768, 2304, 0),
                                       (16, 768, 30528, 1128), (200, 72, 40, 1064)])
def test_gemm_transposed_weight_operand(eng, M, N, K, cfg):
    """the data-gradient form: out[M,N] = A[M,K] . W[K,N] with W read in place through an MN-major descriptor"""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = torch.randn(K, N, device="cuda", generator=g).half()
    ref = A.float() @ W.float()
    out = eng.gemm(A, W, out_fp32=True, tile_cfg=cfg, trans="b")
    out16 = eng.gemm(A, W, out_fp32=False, tile_cfg=cfg, trans="b")
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 2e-3 * scale
    assert (out16.float() - ref).abs().max().item() <= 4e-3 * scale


@pytest.mark.parametrize("M,N,K,ksplit", [(768, 768, 3072, 4), (100, 192, 7680, 30), (1920, 768, 768, 3)])
def test_gemm_split_k_accumulate(eng, M, N, K, ksplit):
    g = torch.Generator(device="cuda").manual_seed(K + ksplit)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = torch.randn(N, K, device="cuda", generator=g).half()
    acc = torch.randn(M, N, device="cuda", generator=g)
    want = acc + A.float() @ W.float().t()
    eng.gemm(A, W, accumulate_into=acc, ksplit=ksplit)
    torch.cuda.synchronize()
    assert (acc - want).abs().max().item() <= 2e-3 * want.abs().max().item()


@pytest.mark.parametrize("B,S", [(2, 120), (1, 128), (3, 17), (2, 64), (2, 65), (1, 1), (8, 120), (2, 210), (1, 256),
                                 (3, 129), (2, 192), (16, 210)])
def test_attention_backward_against_torch_autograd(eng, B, S):
    """d(qkv) from d(ctx): tensor-core kernel (S <= 128) and CUDA-core kernel against fp32 autograd on the same
    16-bit-rounded inputs."""
    H, nH, dH = 768, 12, 64
    g = torch.Generator(device="cuda").manual_seed(S * 7 + B)
    qkv = (torch.randn(B * S, 3 * H, device="cuda", generator=g)).half()
    dctx = (torch.randn(B * S, H, device="cuda", generator=g) * 0.1).half()
    mask = (torch.rand(B, S, device="cuda", generator=g) > 0.3).long()
    mask[:, 0] = 1
    ext = (1.0 - mask.float()) * -10000.0
    x = qkv.float().requires_grad_(True)
    q, k, v = (t.view(B, S, nH, dH).permute(0, 2, 1, 3) for t in x.split(H, dim=1))
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0 + ext[:, None, None, :], -1)
    ctx = (p @ v).permute(0, 2, 1, 3).reshape(B * S, H)
    ctx.backward(dctx.float())
    ref = x.grad
    impls = [1, 0]
    for impl in impls:
        got = eng.attention_backward(qkv, dctx, ext, B, S, impl=impl).float()
        torch.cuda.synchronize()
        for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
            scale = ref[:, sl].abs().max().item()
            err = (got[:, sl] - ref[:, sl]).abs().max().item()
            # P / dS are rounded to 16 bits before the second-stage products on the tensor-core path
            assert err <= (4e-3 if impl == 0 else 1.5e-3) * scale, (impl, name, err / scale)


def test_attention_fully_masked_row_matches_additive_mask_semantics(eng):
    # (1 - mask) * -10000 is ADDITIVE: an all-zero mask row yields softmax over the raw scores, not NaN
    B, S, H = 1, 24, 768
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn(B * S, 3 * H, device="cuda", generator=g).half()
    ext = torch.full((B, S), -10000.0, device="cuda")
    ctx = eng.attention(qkv, ext, B, S, impl=0).float()
    q, k, v = (t.float().view(B, S, 12, 64).permute(0, 2, 1, 3) for t in qkv.split(H, dim=1))
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0 + ext[:, None, None, :], -1)
    ref = (p @ v).permute(0, 2, 1, 3).reshape(B * S, H)
    assert torch.isfinite(ctx).all()
    assert (ctx - ref).abs().max().item() <= 5e-3 * ref.abs().max().item()


def test_layernorm_against_torch(eng):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(1001, 768, device="cuda", generator=g) * 3 + 0.5
    gamma = torch.rand(768, device="cuda", generator=g) + 0.5
    beta = torch.randn(768, device="cuda", generator=g)
    o32, o16 = eng.layernorm(x, gamma, beta, 1e-12)
    ref = torch.nn.functional.layer_norm(x, (768,), gamma, beta, 1e-12)
    assert (o32 - ref).abs().max().item() < 2e-5
    assert (o16.float() - ref).abs().max().item() < 4e-3
