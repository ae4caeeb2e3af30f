// Self-attention backward on the tensor cores for S <= 128 (the RefCOCO geometry, S = 120), one CTA per
// (head, sample).  Probabilities are recomputed from Q, K (never stored by the forward):
//     S  = Q K^T                 (tcgen05, TMEM cols   0..127)      P  = softmax(S / sqrt(dH) + ext_mask)
//     dP = dO V^T                (tcgen05, TMEM cols 128..255)      D  = rowsum(P * dP)
//     dS = P * (dP - D) / sqrt(dH)
//     dQ = dS K                  (A = dS, K-major;  B = K tile, MN-major)          TMEM cols 256..319, lanes = queries
//     dV = P^T dO                (A = P  read MN-major: M = keys; B = dO MN-major)  TMEM cols 320..383, lanes = keys
//     dK = dS^T Q                (A = dS read MN-major;           B = Q  MN-major)  TMEM cols 384..447, lanes = keys
// With dropout on the probabilities (P~ = mask * P / (1-p) multiplied V in the forward): dP is masked and scaled the
// same way before D and dS, and the tile that feeds dV holds P~.
// Two threads own each query row for the softmax algebra (TMEM lane r of S, dP: warps w and w+4 share a lane quadrant
// and take 64 key columns each; row max / sum / D meet in shared memory) and each key row for the dV / dK read-out
// — with one warp per scheduler the kernel was issue-latency bound (ncu: 7.3 cycles per instruction, 13 % issue slots).
// P and dS are written once to shared memory as 16-bit [query][key] tiles in the 128B-swizzled layout; the same bytes
// serve as the K-major A operand of dQ and, through an MN-major descriptor, as the transposed A operand of dV / dK —
// no transpose pass.  Rows / keys beyond S are zeroed so the neighbouring sample's rows that the 128-row TMA boxes
// pull in never contribute.
#pragma once
#include "attention_sm100.cuh"
#include "ptx.cuh"

namespace cptk {

struct AttnBwdParams {
  int B, S, H, nH;
  const float* ext_mask;  // [B, S]
  void* dqkv;             // T16 [B*S, 3H]
  float scale;
  Drop drop;              // dropout that sat on the probabilities in the forward (thresh = 0 -> none)
};

constexpr int kAttnBwdThreads = 256;
constexpr int kAttnBwdSmem = 1024 + 4 * 16384 + 2 * 32768 + 512 + 3 * 1024 + 64;

template <typename T16>
__global__ void __launch_bounds__(kAttnBwdThreads) attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                          const __grid_constant__ CUtensorMap tmap_do,
                                                          const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int S = p.S;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = sQ + 16384, sV = sK + 16384, sdO = sV + 16384;
  const uint32_t sP = sdO + 16384, sdS = sP + 32768;
  const uint32_t sMask = sdS + 32768;
  const uint32_t sStat = sMask + 512;  // [3][2][128] fp32: partial row max, sum, sum(e * dP) of the two column halves
  const uint32_t bar_ld = sStat + 3072, bar_s = bar_ld + 8, bar_g = bar_ld + 16, tmem_slot = bar_ld + 32;
  float* maskp = reinterpret_cast<float*>(smem_raw + (sMask - smem_u32(smem_raw)));
  float* statp = reinterpret_cast<float*>(smem_raw + (sStat - smem_u32(smem_raw)));
  uint8_t* p_gen = smem_raw + (sP - smem_u32(smem_raw));
  uint8_t* ds_gen = smem_raw + (sdS - smem_u32(smem_raw));
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    mbar_init(bar_ld, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_g, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x < 128) maskp[threadIdx.x] = ((int)threadIdx.x < S) ? p.ext_mask[(long long)b * S + threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t cS = 0, cdP = 128, cdQ = 256, cdV = 320, cdK = 384;

  if (threadIdx.x == 0) {
    const int row0 = b * S;
    mbar_expect_tx(bar_ld, 4 * 16384);
    for (int rb = 0; rb < 2; ++rb) {
      tma_load_2d(sQ + rb * 8192, &tmap_qkv, bar_ld, h * kAttnDH, row0 + rb * 64);
      tma_load_2d(sK + rb * 8192, &tmap_qkv, bar_ld, p.H + h * kAttnDH, row0 + rb * 64);
      tma_load_2d(sV + rb * 8192, &tmap_qkv, bar_ld, 2 * p.H + h * kAttnDH, row0 + rb * 64);
      tma_load_2d(sdO + rb * 8192, &tmap_do, bar_ld, h * kAttnDH, row0 + rb * 64);
    }
    mbar_wait(bar_ld, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, 128, Cvt<T16>::kFmt, 0, 0);
    const uint64_t qd = make_smem_desc(sQ, 16, 1024), kd = make_smem_desc(sK, 16, 1024);
    const uint64_t od = make_smem_desc(sdO, 16, 1024), vd = make_smem_desc(sV, 16, 1024);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16(tmem_base + cS, qd + 2 * k, kd + 2 * k, idesc, k != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16(tmem_base + cdP, od + 2 * k, vd + 2 * k, idesc, k != 0);
    umma_commit(bar_s);
  }

  mbar_wait(bar_s, 0);
  tc_fence_after();
  const int q4 = warp & 3, half = warp >> 2;  // TMEM lane quadrant, column half
  const int r = q4 * 32 + lane;
  const uint32_t t_row = tmem_base + (uint32_t(q4 * 32) << 16);
  const int c0 = half * 2;  // this thread's two 32-column chunks
  const unsigned long long pidx0 = (((unsigned long long)b * p.nH + h) * S + r) * S;  // element index of P[b,h,r,0]
  const bool dropping = p.drop.thresh != 0u;
  float mx = -INFINITY;
  for (int c = c0; c < c0 + 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(t_row + cS + c * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = c * 32 + j;
      const float t = (col < S) ? fmaf(__uint_as_float(v[j]), p.scale, maskp[col]) : -INFINITY;
      mx = fmaxf(mx, t);
    }
  }
  statp[half * 128 + r] = mx;
  __syncthreads();
  mx = fmaxf(statp[r], statp[128 + r]);
  float sum = 0.f, dn = 0.f;
  for (int c = c0; c < c0 + 2; ++c) {
    uint32_t v[32], w[32];
    tmem_ld_32x32b_x32(t_row + cS + c * 32, v);
    tmem_ld_32x32b_x32(t_row + cdP + c * 32, w);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = c * 32 + j;
      const float e = (col < S) ? __expf(fmaf(__uint_as_float(v[j]), p.scale, maskp[col]) - mx) : 0.f;
      sum += e;
      float dp = __uint_as_float(w[j]);
      if (dropping) dp = drop_apply(p.drop, pidx0 + col, dp);
      dn = fmaf(e, dp, dn);
    }
  }
  statp[256 + half * 128 + r] = sum;
  statp[512 + half * 128 + r] = dn;
  __syncthreads();
  sum = statp[256 + r] + statp[384 + r];
  dn = statp[512 + r] + statp[640 + r];
  const bool valid = r < S;
  const float inv = valid ? 1.0f / sum : 0.f;
  const float D = dn * inv;
  for (int c = c0; c < c0 + 2; ++c) {
    uint32_t v[32], w[32];
    tmem_ld_32x32b_x32(t_row + cS + c * 32, v);
    tmem_ld_32x32b_x32(t_row + cdP + c * 32, w);
    tmem_ld_wait();
    float pr[32], ds[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = c * 32 + j;
      const float e = (col < S) ? __expf(fmaf(__uint_as_float(v[j]), p.scale, maskp[col]) - mx) : 0.f;
      pr[j] = e * inv;
      float dp = __uint_as_float(w[j]);
      if (dropping) dp = drop_apply(p.drop, pidx0 + col, dp);
      ds[j] = pr[j] * (dp - D) * p.scale;
      if (dropping) pr[j] = drop_apply(p.drop, pidx0 + col, pr[j]);  // the dV operand is P~
    }
    const int kb = c >> 1;
    uint8_t* prow = p_gen + kb * 16384 + r * 128;
    uint8_t* drow = ds_gen + kb * 16384 + r * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u, d;
      u.x = Cvt<T16>::pack2(pr[8 * i + 0], pr[8 * i + 1]);
      u.y = Cvt<T16>::pack2(pr[8 * i + 2], pr[8 * i + 3]);
      u.z = Cvt<T16>::pack2(pr[8 * i + 4], pr[8 * i + 5]);
      u.w = Cvt<T16>::pack2(pr[8 * i + 6], pr[8 * i + 7]);
      d.x = Cvt<T16>::pack2(ds[8 * i + 0], ds[8 * i + 1]);
      d.y = Cvt<T16>::pack2(ds[8 * i + 2], ds[8 * i + 3]);
      d.z = Cvt<T16>::pack2(ds[8 * i + 4], ds[8 * i + 5]);
      d.w = Cvt<T16>::pack2(ds[8 * i + 6], ds[8 * i + 7]);
      const int chunk = ((c & 1) * 4 + i) ^ (r & 7);
      *reinterpret_cast<uint4*>(prow + chunk * 16) = u;
      *reinterpret_cast<uint4*>(drow + chunk * 16) = d;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after();
    // dQ = dS K : A K-major [queries x keys], B = K tile rows = keys (the product's K dim) -> MN-major
    const uint32_t id_q = make_idesc_f16(128, kAttnDH, Cvt<T16>::kFmt, 0, 1);
    for (int k = 0; k < 8; ++k) {
      const uint64_t ad = make_smem_desc(sdS + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
      const uint64_t bd = make_smem_desc(sK + k * 2048, 1024, 1024);
      umma_f16(tmem_base + cdQ, ad, bd, id_q, k != 0);
    }
    // dV = P^T dO, dK = dS^T Q : A read MN-major (M = keys: two 64-key blocks 16 KB apart, 8-query groups 1 KB apart)
    const uint32_t id_t = make_idesc_f16(128, kAttnDH, Cvt<T16>::kFmt, 1, 1);
    for (int k = 0; k < 8; ++k) {
      const uint64_t ad = make_smem_desc(sP + k * 2048, 16384, 1024);
      const uint64_t bd = make_smem_desc(sdO + k * 2048, 1024, 1024);
      umma_f16(tmem_base + cdV, ad, bd, id_t, k != 0);
    }
    for (int k = 0; k < 8; ++k) {
      const uint64_t ad = make_smem_desc(sdS + k * 2048, 16384, 1024);
      const uint64_t bd = make_smem_desc(sQ + k * 2048, 1024, 1024);
      umma_f16(tmem_base + cdK, ad, bd, id_t, k != 0);
    }
    umma_commit(bar_g);
  }
  mbar_wait(bar_g, 0);
  tc_fence_after();
  {
    // tcgen05.ld is warp-collective: every lane reads, only rows < S store
    T16* dst = reinterpret_cast<T16*>(p.dqkv) + ((long long)b * S + r) * 3 * p.H + h * kAttnDH;
    const uint32_t cols[3] = {cdQ, cdK, cdV};
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      {
        const int c = half;  // this thread's 32 of the 64 head-dim columns
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + cols[which] + c * 32, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]));
            u.y = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
            u.z = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
            u.w = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
            *reinterpret_cast<uint4*>(dst + which * p.H + c * 32 + i * 8) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// 128 < S <= 256 (GQA / VCR: S = 210): two 128-row query tiles x two 128-key tiles per (head, sample), still one CTA.
// TMEM cannot hold everything at once (S + dP over 256 keys already fill its 512 columns), so the kernel runs
//   a pre-pass per query tile : S = Q K^T and dP = dO V^T over ALL keys (2 x 256 columns) -> row max, 1 / row sum and
//                               D = sum_j P dP~ kept in shared memory;
//   the main pass, key tile outer / query tile inner : S and dP recomputed for the 128 x 128 block (256 columns),
//                               P~ and dS tiles written to shared memory, then  dQ[qt] += dS K[kt]  (2 x 64 columns,
//                               live for the whole kernel),  dV[kt] += P~^T dO[qt],  dK[kt] += dS^T Q[qt]  (2 x 64
//                               columns, read out after the two query tiles of the key tile).
// 256 + 128 + 128 = 512 columns.  Operand forms, thread mapping and the dropout replay are those of the kernel above.
constexpr int kAttnBwd2Smem = 1024 + 4 * 32768 + 2 * 32768 + 1024 + 3072 + 3072 + 64;

template <typename T16>
__global__ void __launch_bounds__(kAttnBwdThreads) attn_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                       const __grid_constant__ CUtensorMap tmap_do,
                                                                       const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int S = p.S;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = sQ + 32768, sV = sK + 32768, sdO = sV + 32768;
  const uint32_t sP = sdO + 32768, sdS = sP + 32768;
  const uint32_t sMask = sdS + 32768;  // 256 floats
  const uint32_t sRow = sMask + 1024;  // [3][256]: row max, 1 / row sum, D
  const uint32_t sStat = sRow + 3072;  // [3][2][128]: partials of the two column halves
  const uint32_t bar_ld = sStat + 3072, bar_s = bar_ld + 8, bar_g = bar_ld + 16, tmem_slot = bar_ld + 32;
  const uint32_t s0 = smem_u32(smem_raw);
  float* maskp = reinterpret_cast<float*>(smem_raw + (sMask - s0));
  float* rowp = reinterpret_cast<float*>(smem_raw + (sRow - s0));
  float* statp = reinterpret_cast<float*>(smem_raw + (sStat - s0));
  uint8_t* p_gen = smem_raw + (sP - s0);
  uint8_t* ds_gen = smem_raw + (sdS - s0);
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    mbar_init(bar_ld, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_g, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  maskp[threadIdx.x] = ((int)threadIdx.x < S) ? p.ext_mask[(long long)b * S + threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t cS = 0, cdP = 128, cdQ = 256, cdK = 384, cdV = 448;

  const int q4 = warp & 3, half = warp >> 2;
  const int r = q4 * 32 + lane;  // row inside a 128-row tile == TMEM lane
  const uint32_t t_row = tmem_base + (uint32_t(q4 * 32) << 16);
  const bool dropping = p.drop.thresh != 0u;
  const unsigned long long pbase = ((unsigned long long)b * p.nH + h) * S * S;
  uint32_t ph_s = 0, ph_g = 0;

  if (threadIdx.x == 0) {
    const int row0 = b * S;
    mbar_expect_tx(bar_ld, 4 * 32768);
    for (int rb = 0; rb < 4; ++rb) {
      tma_load_2d(sQ + rb * 8192, &tmap_qkv, bar_ld, h * kAttnDH, row0 + rb * 64);
      tma_load_2d(sK + rb * 8192, &tmap_qkv, bar_ld, p.H + h * kAttnDH, row0 + rb * 64);
      tma_load_2d(sV + rb * 8192, &tmap_qkv, bar_ld, 2 * p.H + h * kAttnDH, row0 + rb * 64);
      tma_load_2d(sdO + rb * 8192, &tmap_do, bar_ld, h * kAttnDH, row0 + rb * 64);
    }
    mbar_wait(bar_ld, 0);
  }

  // ---- pre-pass: row statistics over all 256 key columns
  for (int qt = 0; qt < 2; ++qt) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16(128, 256, Cvt<T16>::kFmt, 0, 0);
      const uint64_t qd = make_smem_desc(sQ + qt * 16384, 16, 1024), kd = make_smem_desc(sK, 16, 1024);
      const uint64_t od = make_smem_desc(sdO + qt * 16384, 16, 1024), vd = make_smem_desc(sV, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base, qd + 2 * k, kd + 2 * k, idesc, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base + 256, od + 2 * k, vd + 2 * k, idesc, k != 0);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, ph_s);
    ph_s ^= 1u;
    tc_fence_after();
    const int g = qt * 128 + r;  // query row within the sample
    float mx = -INFINITY;
    for (int c = half * 4; c < half * 4 + 4; ++c) {
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
    statp[half * 128 + r] = mx;
    __syncthreads();
    mx = fmaxf(statp[r], statp[128 + r]);
    float sum = 0.f, dn = 0.f;
    for (int c = half * 4; c < half * 4 + 4; ++c) {
      uint32_t v[32], w[32];
      tmem_ld_32x32b_x32(t_row + c * 32, v);
      tmem_ld_32x32b_x32(t_row + 256 + c * 32, w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = c * 32 + j;
        const float e = (col < S) ? __expf(fmaf(__uint_as_float(v[j]), p.scale, maskp[col]) - mx) : 0.f;
        sum += e;
        float dp = __uint_as_float(w[j]);
        if (dropping) dp = drop_apply(p.drop, pbase + (unsigned long long)g * S + col, dp);
        dn = fmaf(e, dp, dn);
      }
    }
    statp[256 + half * 128 + r] = sum;
    statp[512 + half * 128 + r] = dn;
    __syncthreads();
    if (half == 0) {
      const float l = statp[256 + r] + statp[384 + r];
      const float inv = (g < S) ? 1.0f / l : 0.f;
      rowp[g] = mx;
      rowp[256 + g] = inv;
      rowp[512 + g] = (statp[512 + r] + statp[640 + r]) * inv;
    }
    tc_fence_before();
    __syncthreads();  // TMEM reads done before the next MMAs overwrite it; row statistics visible
  }

  // ---- main pass
  T16* const out = reinterpret_cast<T16*>(p.dqkv) + (long long)b * S * 3 * p.H + h * kAttnDH;
  auto store32 = [&](uint32_t tcol, long long row, int which, bool ok) {
    // every lane reads (tcgen05.ld is warp-collective); rows beyond S do not store
    uint32_t v[32];
    tmem_ld_32x32b_x32(t_row + tcol + half * 32, v);
    tmem_ld_wait();
    if (ok) {
      T16* dst = out + row * 3 * p.H + which * p.H + half * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]));
        u.y = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
        u.z = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
        u.w = Cvt<T16>::pack2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
        *reinterpret_cast<uint4*>(dst + i * 8) = u;
      }
    }
  };
  for (int kt = 0; kt < 2; ++kt) {
    for (int qt = 0; qt < 2; ++qt) {
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16(128, 128, Cvt<T16>::kFmt, 0, 0);
        const uint64_t qd = make_smem_desc(sQ + qt * 16384, 16, 1024), kd = make_smem_desc(sK + kt * 16384, 16, 1024);
        const uint64_t od = make_smem_desc(sdO + qt * 16384, 16, 1024), vd = make_smem_desc(sV + kt * 16384, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base + cS, qd + 2 * k, kd + 2 * k, idesc, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base + cdP, od + 2 * k, vd + 2 * k, idesc, k != 0);
        umma_commit(bar_s);  // arrives after the previous block's dQ / dV / dK products too (in-order pipe)
      }
      mbar_wait(bar_s, ph_s);
      ph_s ^= 1u;
      tc_fence_after();
      const int g = qt * 128 + r;
      const float mx = rowp[g], inv = rowp[256 + g], D = rowp[512 + g];
      for (int c = half * 2; c < half * 2 + 2; ++c) {
        uint32_t v[32], w[32];
        tmem_ld_32x32b_x32(t_row + cS + c * 32, v);
        tmem_ld_32x32b_x32(t_row + cdP + c * 32, w);
        tmem_ld_wait();
        float pr[32], ds[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = kt * 128 + c * 32 + j;
          const float e = (col < S) ? __expf(fmaf(__uint_as_float(v[j]), p.scale, maskp[col]) - mx) : 0.f;
          pr[j] = e * inv;
          float dp = __uint_as_float(w[j]);
          if (dropping) {
            const unsigned long long e_idx = pbase + (unsigned long long)g * S + col;
            dp = drop_apply(p.drop, e_idx, dp);
            ds[j] = pr[j] * (dp - D) * p.scale;
            pr[j] = drop_apply(p.drop, e_idx, pr[j]);
          } else {
            ds[j] = pr[j] * (dp - D) * p.scale;
          }
        }
        const int kb = c >> 1;
        uint8_t* prow = p_gen + kb * 16384 + r * 128;
        uint8_t* drow = ds_gen + kb * 16384 + r * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u, d;
          u.x = Cvt<T16>::pack2(pr[8 * i + 0], pr[8 * i + 1]);
          u.y = Cvt<T16>::pack2(pr[8 * i + 2], pr[8 * i + 3]);
          u.z = Cvt<T16>::pack2(pr[8 * i + 4], pr[8 * i + 5]);
          u.w = Cvt<T16>::pack2(pr[8 * i + 6], pr[8 * i + 7]);
          d.x = Cvt<T16>::pack2(ds[8 * i + 0], ds[8 * i + 1]);
          d.y = Cvt<T16>::pack2(ds[8 * i + 2], ds[8 * i + 3]);
          d.z = Cvt<T16>::pack2(ds[8 * i + 4], ds[8 * i + 5]);
          d.w = Cvt<T16>::pack2(ds[8 * i + 6], ds[8 * i + 7]);
          const int chunk = ((c & 1) * 4 + i) ^ (r & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = u;
          *reinterpret_cast<uint4*>(drow + chunk * 16) = d;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t id_q = make_idesc_f16(128, kAttnDH, Cvt<T16>::kFmt, 0, 1);
        for (int k = 0; k < 8; ++k) {
          const uint64_t ad = make_smem_desc(sdS + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          const uint64_t bd = make_smem_desc(sK + kt * 16384 + k * 2048, 1024, 1024);
          umma_f16(tmem_base + cdQ + qt * 64, ad, bd, id_q, (kt | k) != 0);
        }
        const uint32_t id_t = make_idesc_f16(128, kAttnDH, Cvt<T16>::kFmt, 1, 1);
        for (int k = 0; k < 8; ++k) {
          const uint64_t ad = make_smem_desc(sP + k * 2048, 16384, 1024);
          const uint64_t bd = make_smem_desc(sdO + qt * 16384 + k * 2048, 1024, 1024);
          umma_f16(tmem_base + cdV, ad, bd, id_t, (qt | k) != 0);
        }
        for (int k = 0; k < 8; ++k) {
          const uint64_t ad = make_smem_desc(sdS + k * 2048, 16384, 1024);
          const uint64_t bd = make_smem_desc(sQ + qt * 16384 + k * 2048, 1024, 1024);
          umma_f16(tmem_base + cdK, ad, bd, id_t, (qt | k) != 0);
        }
        if (qt == 1) umma_commit(bar_g);
      }
    }
    // dK, dV of this key tile: lanes = keys
    mbar_wait(bar_g, ph_g);
    ph_g ^= 1u;
    tc_fence_after();
    store32(cdK, kt * 128 + r, 1, kt * 128 + r < S);
    store32(cdV, kt * 128 + r, 2, kt * 128 + r < S);
    tc_fence_before();
    __syncthreads();
  }
  // dQ of both query tiles (every MMA has completed: bar_g of the last key tile)
  tc_fence_after();
  store32(cdQ, r, 0, r < S);
  store32(cdQ + 64, 128 + r, 0, 128 + r < S);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace cptk
