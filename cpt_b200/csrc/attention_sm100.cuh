// Fused self-attention for the CPT sequence lengths (S = T + R <= 256: 120 for RefCOCO, 210 for GQA/VCR), one
// CTA per (head, 128-query tile, sample).  Restates CaptionBertSelfAttention.forward
// (/root/reference/Oscar/oscar/modeling/modeling_bert.py:47-67):
//     P = softmax(Q K^T / sqrt(dH) + ext_mask)   ext_mask = (1 - mask) * -10000  (additive, NOT -inf)
//     ctx = P V, heads merged back to [B*S, H]
// The whole K and V of a (sample, head) fit in shared memory, so the softmax is single-pass over the full row
// (no online rescaling): TMA stages Q/K/V tiles (128B swizzle) -> tcgen05.mma S = Q K^T into TMEM -> each of
// 128 threads owns one query row: max / exp / sum in fp32 from TMEM, P written back to smem as the 16-bit
// K-major A operand -> tcgen05.mma O = P V (V is the MN-major B operand, straight from the TMA tile) ->
// O / rowsum -> 16-bit ctx.  The [B,nH,S,S] score/probability tensors the reference materialises in HBM never
// leave the SM.
#pragma once
#include "ptx.cuh"

namespace cptk {

constexpr int kAttnThreads = 128;
constexpr int kAttnDH = 64;

struct AttnParams {
  int B, S, H, nH;
  const float* ext_mask;  // [B, S] additive fp32
  void* ctx;              // T16 [B*S, H]
  float scale;            // 1/sqrt(dH)
};

__host__ __device__ inline int attn_nkb(int S) { return (S + 63) / 64; }
inline size_t attn_smem_bytes(int S) {
  const int nkb = attn_nkb(S);
  // Q 16 KB | K nkb*8 KB | V nkb*8 KB | P nkb*16 KB | mask 1 KB | barriers
  return 1024 + 16384 + (size_t)nkb * 8192 * 2 + (size_t)nkb * 16384 + 1024 + 64;
}

template <typename T16>
__global__ void __launch_bounds__(kAttnThreads) attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                               const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int S = p.S;
  const int nkb = attn_nkb(S);
  const int NK = (S + 15) & ~15;  // UMMA N for S = Q K^T, and the K extent of O = P V
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sK = sQ + 16384;
  const uint32_t sV = sK + nkb * 8192;
  const uint32_t sP = sV + nkb * 8192;
  const uint32_t sMask = sP + nkb * 16384;
  const uint32_t bar_qk = sMask + 1024, bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_o = bar_qk + 24;
  const uint32_t tmem_slot = bar_qk + 32;
  float* maskp = reinterpret_cast<float*>(smem_raw + (sMask - smem_u32(smem_raw)));
  uint8_t* p_gen = smem_raw + (sP - smem_u32(smem_raw));

  const int h = blockIdx.x, i0 = blockIdx.y * 128, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (NK <= 128) ? 256u : 512u;
  const uint32_t o_col = (NK <= 128) ? 128u : 256u;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  for (int j = threadIdx.x; j < nkb * 64; j += kAttnThreads)
    maskp[j] = (j < S) ? p.ext_mask[(long long)b * S + j] : 0.f;
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (threadIdx.x == 0) {
    const int row0 = b * S;
    mbar_expect_tx(bar_qk, 16384 + nkb * 8192);
    tma_load_2d(sQ, &tmap_qkv, bar_qk, h * kAttnDH, row0 + i0);
    tma_load_2d(sQ + 8192, &tmap_qkv, bar_qk, h * kAttnDH, row0 + i0 + 64);
    for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sK + kb * 8192, &tmap_qkv, bar_qk, p.H + h * kAttnDH, row0 + kb * 64);
    mbar_expect_tx(bar_v, nkb * 8192);
    for (int kb = 0; kb < nkb; ++kb)
      tma_load_2d(sV + kb * 8192, &tmap_qkv, bar_v, 2 * p.H + h * kAttnDH, row0 + kb * 64);
    // S = Q K^T : M=128, N=NK, K=64 (4 x UMMA_K)
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, NK, Cvt<T16>::kFmt, 0, 0);
    const uint64_t qd = make_smem_desc(sQ, 16, 1024), kd = make_smem_desc(sK, 16, 1024);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16(tmem_base, qd + 2 * k, kd + 2 * k, idesc, k != 0);
    umma_commit(bar_s);
  }

  // ---- softmax: thread = one query row (TMEM lane), two passes over the score row held in TMEM
  mbar_wait(bar_s, 0);
  tc_fence_after();
  const int r = warp * 32 + lane;  // row within the tile == TMEM lane
  const uint32_t t_row = tmem_base + (uint32_t(warp * 32) << 16);
  const int nchunk = (NK + 31) / 32;
  float mx = -INFINITY;
  for (int c = 0; c < nchunk; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(t_row + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = c * 32 + j;
      const float t = (col < S) ? fmaf(__uint_as_float(v[j]), p.scale, maskp[col]) : -INFINITY;
      mx = fmaxf(mx, t);
    }
  }
  float sum = 0.f;
  for (int c = 0; c < nchunk; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(t_row + c * 32, v);
    tmem_ld_wait();
    float e[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = c * 32 + j;
      const float t = (col < S) ? fmaf(__uint_as_float(v[j]), p.scale, maskp[col]) : -INFINITY;
      e[j] = __expf(t - mx);
      sum += e[j];
    }
    // P[r, c*32 .. +31] -> K-major 128B-swizzled A-operand tile (64-key block kb, 16-byte chunk ^ (row & 7))
    const int kb = c >> 1;
    uint8_t* prow = p_gen + kb * 16384 + r * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = Cvt<T16>::pack2(e[8 * i + 0], e[8 * i + 1]);
      u.y = Cvt<T16>::pack2(e[8 * i + 2], e[8 * i + 3]);
      u.z = Cvt<T16>::pack2(e[8 * i + 4], e[8 * i + 5]);
      u.w = Cvt<T16>::pack2(e[8 * i + 6], e[8 * i + 7]);
      const int chunk = ((c & 1) * 4 + i) ^ (r & 7);
      *reinterpret_cast<uint4*>(prow + chunk * 16) = u;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    // O = P V : M=128, N=64 (dH), K=NK.  V tile rows are keys -> MN-major B operand.
    tc_fence_after();
    mbar_wait(bar_v, 0);
    const uint32_t idesc = make_idesc_f16(128, kAttnDH, Cvt<T16>::kFmt, 0, 1);
    for (int k = 0; k < NK / 16; ++k) {
      const uint64_t pd = make_smem_desc(sP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
      const uint64_t vd = make_smem_desc(sV + k * 2048, 1024, 1024);
      umma_f16(tmem_base + o_col, pd, vd, idesc, k != 0);
    }
    umma_commit(bar_o);
  }
  mbar_wait(bar_o, 0);
  tc_fence_after();
  {
    const float inv = 1.0f / sum;
    const int qi = i0 + r;
    T16* dst = reinterpret_cast<T16*>(p.ctx) + ((long long)b * S + qi) * p.H + h * kAttnDH;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + o_col + c * 32, v);
      tmem_ld_wait();
      if (qi < S) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
          u.y = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
          u.z = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
          u.w = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 32 + i * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Plain CUDA-core restatement of the same op (fp32 math on the same 16-bit Q/K/V): used by the GPU tests as an
// on-device cross-check of the tensor-core kernel, and selectable with CPT_B200_ATTN=simt for debugging.  It is
// a CUDA kernel, not a CPU fallback.
template <typename T16>
__global__ void __launch_bounds__(128) attn_simt_kernel(const T16* __restrict__ qkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int S = p.S, H = p.H;
  T16* sK = reinterpret_cast<T16*>(smem_raw);
  T16* sV = sK + (size_t)S * kAttnDH;
  float* sM = reinterpret_cast<float*>(sV + (size_t)S * kAttnDH);
  const int h = blockIdx.x, b = blockIdx.y;
  const T16* base = qkv + (long long)b * S * 3 * H;
  for (int i = threadIdx.x; i < S * 8; i += blockDim.x) {
    const int j = i >> 3, c = (i & 7) * 8;
    *reinterpret_cast<uint4*>(sK + j * kAttnDH + c) =
        *reinterpret_cast<const uint4*>(base + (long long)j * 3 * H + H + h * kAttnDH + c);
    *reinterpret_cast<uint4*>(sV + j * kAttnDH + c) =
        *reinterpret_cast<const uint4*>(base + (long long)j * 3 * H + 2 * H + h * kAttnDH + c);
  }
  for (int j = threadIdx.x; j < S; j += blockDim.x) sM[j] = p.ext_mask[(long long)b * S + j];
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    float q[kAttnDH];
    const T16* qp = base + (long long)i * 3 * H + h * kAttnDH;
#pragma unroll
    for (int d = 0; d < kAttnDH; ++d) q[d] = Cvt<T16>::to(qp[d]);
    float mx = -INFINITY;
    for (int j = 0; j < S; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) s = fmaf(q[d], Cvt<T16>::to(sK[j * kAttnDH + d]), s);
      mx = fmaxf(mx, fmaf(s, p.scale, sM[j]));
    }
    float o[kAttnDH];
#pragma unroll
    for (int d = 0; d < kAttnDH; ++d) o[d] = 0.f;
    float sum = 0.f;
    for (int j = 0; j < S; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) s = fmaf(q[d], Cvt<T16>::to(sK[j * kAttnDH + d]), s);
      const float e = expf(fmaf(s, p.scale, sM[j]) - mx);
      sum += e;
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) o[d] = fmaf(e, Cvt<T16>::to(sV[j * kAttnDH + d]), o[d]);
    }
    const float inv = 1.0f / sum;
    T16* dst = reinterpret_cast<T16*>(p.ctx) + ((long long)b * S + i) * H + h * kAttnDH;
#pragma unroll
    for (int d = 0; d < kAttnDH; ++d) dst[d] = Cvt<T16>::from(o[d] * inv);
  }
}

// K4: ext_mask = (1 - mask) * -10000 in fp32, exactly as modeling_bert.py:213-226 for a 2-D mask.
__global__ void ext_mask_kernel(const long long* __restrict__ mask, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (1.0f - (float)mask[i]) * -10000.0f;
}

}  // namespace cptk
