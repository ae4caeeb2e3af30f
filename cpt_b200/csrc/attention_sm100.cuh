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
  long long* trace;       // optional [gridDim.x][8] cycle counters (debug)
  Drop drop;              // training only (attn_tc_kernel): dropout on the probabilities, thresh = 0 -> off
  // attn_pp_kernel after a chain launch: per 128-row tile, the chain's "QKV rows published" counter and the value that
  // means complete.  The kernel then does NOT wait for the whole previous grid (griddepcontrol.wait): its CTAs take the
  // SMs the chain's pairs free one by one and start on the samples whose rows are there.  NULL: plain PDL ordering.
  const unsigned* qkv_ready = nullptr;
  unsigned qkv_target = 0;
  // ... and before a chain launch that starts early itself: per 128-row tile of the context rows, this kernel counts
  // rows x heads written (complete at 128 * nH; the partial last tile is padded up once by block 0)
  unsigned* ctx_done = nullptr;
};

__device__ __forceinline__ void attn_rows_wait(const unsigned* addr, unsigned target) {
  unsigned v, spins = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    if (v >= target) break;
    if (++spins > (1u << 23)) {
      printf("cpt_b200: attention waited for QKV rows that never came (block %d: have %u need %u)\n", blockIdx.x, v, target);
      __trap();
    }
    __nanosleep(64);
  }
}

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

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  pdl_wait();
  for (int j = threadIdx.x; j < nkb * 64; j += kAttnThreads)
    maskp[j] = (j < S) ? p.ext_mask[(long long)b * S + j] : 0.f;
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
    if (p.drop.thresh) {
      // training: P~ = mask * P / (1 - p) multiplies V; the row sum stays that of the unmasked probabilities
      const unsigned long long pidx = (((unsigned long long)b * p.nH + h) * S + (i0 + r)) * S + c * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) e[j] = drop_apply(p.drop, pidx + j, e[j]);
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
// Production kernel: the same math as attn_tc_kernel, restructured as a persistent, warp-specialised pipeline
// (one CTA per SM, work item = (sample, head, 128-query tile)) so that TMA, the two tensor-core GEMMs and the
// softmax of consecutive items overlap:
//   warp 0      : TMA producer            (Q, K, V tiles up to NSLOT-1 items ahead)
//   warp 1      : tcgen05.mma issuer      (S(i+1) = Q K^T is issued before waiting for P(i); O(i) = P V)
//   warps 2..9  : softmax + output        (2 threads per query row, each owning half of the keys / of dH; the
//                                          half-row of scores is read from TMEM ONCE and stays in registers for
//                                          max, exp and sum; with >= 3 slots the O(i) read-out is deferred until
//                                          after the softmax of item i+1 so it never waits for the P V MMA)
// Per slot the probability tile P aliases the Q/K tiles (dead once S is in TMEM) and the output accumulator O
// aliases the first 64 TMEM columns of S (dead once P is in smem).  NCH = 64-key blocks (= 32-column score chunks
// per thread): S <= 128 -> NCH 2, 4 slots (192 KB smem, 512 TMEM columns); S <= 256 -> NCH 4, 2 slots.
constexpr int kAttn2Threads = 64 + 256;

template <int NCH>
struct Attn2Cfg {
  static constexpr int kQK = 16384 + NCH * 8192, kP = NCH * 16384;
  static constexpr int kSlotBytes = (kQK > kP ? kQK : kP) + NCH * 8192;
  static constexpr int kVOff = kSlotBytes - NCH * 8192;
  static constexpr int kTmemStride = NCH * 64;                 // score columns per slot (O aliases the first 64)
  static constexpr int kSlots = (NCH <= 2) ? 4 : 2;
  static constexpr int kSmemBytes = 1024 + kSlots * kSlotBytes + 1024 /*mask*/ + 2048 /*row stats*/ + 256;
  static_assert(kSlots * kTmemStride <= 512, "TMEM");
};

template <typename T16, int NCH>
__global__ void __launch_bounds__(kAttn2Threads, 1) attn_pipe_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                     const AttnParams p) {
  using Cfg = Attn2Cfg<NCH>;
  constexpr int NSLOT = Cfg::kSlots;
  constexpr bool kDefer = NSLOT >= 3;
  extern __shared__ uint8_t smem_raw[];
  const int S = p.S;
  // UMMA N of S = Q K^T and K extent of O = P V: whole 32-column softmax chunks, so every score column the softmax
  // reads was written by the MMA (extra key rows are other samples' rows or TMA zero fill; their mask is -inf)
  const int NK = (S + 31) & ~31;
  const int nchunk = (NK + 31) / 32;      // live 32-column score chunks (<= 2 * NCH)
  const int n_mt = (S + 127) / 128;
  const int n_items = p.B * p.nH * n_mt;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* mask_s = reinterpret_cast<float*>(gen + NSLOT * Cfg::kSlotBytes);
  float* hmax = mask_s + 256;   // [2][128]
  float* hsum = hmax + 256;     // [2][128]
  const uint32_t bars = base + NSLOT * Cfg::kSlotBytes + 1024 + 2048;
  enum { QK_FULL = 0, V_FULL, S_FULL, P_READY, O_FULL, SLOT_FREE };
  auto bar = [&](int which, int s) { return bars + 8u * (which * NSLOT + s); };
  const uint32_t tmem_slot = bars + 8u * 6 * NSLOT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(bar(QK_FULL, s), 1);
      mbar_init(bar(V_FULL, s), 1);
      mbar_init(bar(S_FULL, s), 1);
      mbar_init(bar(P_READY, s), 1);
      mbar_init(bar(O_FULL, s), 1);
      mbar_init(bar(SLOT_FREE, s), 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto decode = [&](int item, int& b, int& h, int& mt) {
    mt = item % n_mt;
    h = (item / n_mt) % p.nH;
    b = item / (n_mt * p.nH);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int i = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
        const int s = i % NSLOT;
        const uint32_t par = (i / NSLOT) & 1u;
        int b, h, mt;
        decode(item, b, h, mt);
        const int row0 = b * S;
        const uint32_t sQ = base + s * Cfg::kSlotBytes, sK = sQ + 16384, sV = sQ + Cfg::kVOff;
        const long long t0 = clock64();
        mbar_wait(bar(SLOT_FREE, s), par ^ 1u);
        if (p.trace) p.trace[blockIdx.x * 16 + 0] += clock64() - t0;  // producer: waiting for a free slot
        mbar_expect_tx(bar(QK_FULL, s), 16384 + NCH * 8192);
        tma_load_2d(sQ, &tmap_qkv, bar(QK_FULL, s), h * kAttnDH, row0 + mt * 128);
        tma_load_2d(sQ + 8192, &tmap_qkv, bar(QK_FULL, s), h * kAttnDH, row0 + mt * 128 + 64);
#pragma unroll
        for (int kb = 0; kb < NCH; ++kb)
          tma_load_2d(sK + kb * 8192, &tmap_qkv, bar(QK_FULL, s), p.H + h * kAttnDH, row0 + kb * 64);
        mbar_expect_tx(bar(V_FULL, s), NCH * 8192);
#pragma unroll
        for (int kb = 0; kb < NCH; ++kb)
          tma_load_2d(sV + kb * 8192, &tmap_qkv, bar(V_FULL, s), 2 * p.H + h * kAttnDH, row0 + kb * 64);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, NK, Cvt<T16>::kFmt, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, kAttnDH, Cvt<T16>::kFmt, 0, 1);
      const int n_mine = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      auto issue_qk = [&](int i) {
        const int s = i % NSLOT;
        const uint32_t par = (i / NSLOT) & 1u;
        const long long t0 = clock64();
        mbar_wait(bar(QK_FULL, s), par);
        if (p.trace) p.trace[blockIdx.x * 16 + 1] += clock64() - t0;  // MMA: waiting for Q/K tiles (TMA)
        tc_fence_after();
        const uint32_t sQ = base + s * Cfg::kSlotBytes;
        const uint64_t qd = make_smem_desc(sQ, 16, 1024), kd = make_smem_desc(sQ + 16384, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem_base + s * Cfg::kTmemStride, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
        umma_commit(bar(S_FULL, s));
      };
      if (n_mine > 0) issue_qk(0);
      for (int i = 0; i < n_mine; ++i) {
        const int s = i % NSLOT;
        const uint32_t par = (i / NSLOT) & 1u;
        if (i + 1 < n_mine) issue_qk(i + 1);
        const long long t0 = clock64();
        mbar_wait(bar(P_READY, s), par);
        if (p.trace) p.trace[blockIdx.x * 16 + 2] += clock64() - t0;  // MMA: waiting for the softmax
        mbar_wait(bar(V_FULL, s), par);
        tc_fence_after();
        const uint32_t sP = base + s * Cfg::kSlotBytes, sV = sP + Cfg::kVOff;
        for (int k = 0; k < NK / 16; ++k) {
          const uint64_t pd = make_smem_desc(sP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          const uint64_t vd = make_smem_desc(sV + k * 2048, 1024, 1024);
          umma_f16(tmem_base + s * Cfg::kTmemStride, pd, vd, idesc_o, k != 0);
        }
        umma_commit(bar(O_FULL, s));
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + output (256 threads)
    const int ew = warp - 2, q = warp & 3, half = ew >> 2;
    const int tid = threadIdx.x - 64;
    const int r = q * 32 + lane;        // query row within the tile == TMEM lane
    const int c0 = half * NCH;          // my chunks: [c0, c0 + NCH), live while < nchunk
    // read-out of one finished item: O / rowsum -> 16-bit ctx
    auto read_out = [&](int i, int item, float inv) {
      const int s = i % NSLOT;
      const uint32_t par = (i / NSLOT) & 1u;
      int b, h, mt;
      decode(item, b, h, mt);
      const long long t0 = clock64();
      mbar_wait(bar(O_FULL, s), par);
      if (p.trace && tid == 0) p.trace[blockIdx.x * 16 + 4] += clock64() - t0;  // softmax warps: waiting for O
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + s * Cfg::kTmemStride + half * 32, v);
      tmem_ld_wait();
      const int qi = mt * 128 + r;
      if (qi < S) {
        T16* dst = reinterpret_cast<T16*>(p.ctx) + ((long long)b * S + qi) * p.H + h * kAttnDH + half * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 u;
          u.x = Cvt<T16>::pack2(__uint_as_float(v[8 * k + 0]) * inv, __uint_as_float(v[8 * k + 1]) * inv);
          u.y = Cvt<T16>::pack2(__uint_as_float(v[8 * k + 2]) * inv, __uint_as_float(v[8 * k + 3]) * inv);
          u.z = Cvt<T16>::pack2(__uint_as_float(v[8 * k + 4]) * inv, __uint_as_float(v[8 * k + 5]) * inv);
          u.w = Cvt<T16>::pack2(__uint_as_float(v[8 * k + 6]) * inv, __uint_as_float(v[8 * k + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + k * 8) = u;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(SLOT_FREE, s));
    };

    int i = 0, prev_item = -1;
    float prev_inv = 0.f;
    const long long t_begin = clock64();
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
      const int s = i % NSLOT;
      const uint32_t par = (i / NSLOT) & 1u;
      const int b = item / (n_mt * p.nH);
      if (tid < NCH * 64)
        mask_s[tid] = (tid < S) ? p.ext_mask[(long long)b * S + tid] * 1.4426950408889634f : -INFINITY;
      named_bar_sync(1, 256);
      const long long t0 = clock64();
      mbar_wait(bar(S_FULL, s), par);
      if (p.trace && tid == 0) p.trace[blockIdx.x * 16 + 3] += clock64() - t0;  // softmax warps: waiting for S
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + s * Cfg::kTmemStride;
      uint32_t v[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if (c0 + c < nchunk) tmem_ld_32x32b_x32(t_row + (c0 + c) * 32, v[c]);
      tmem_ld_wait();
      // t = (s / sqrt(dH) + ext_mask) * log2 e (modeling_bert.py:47-50), kept in place; the mask of padded key
      // columns is -inf so they are excluded outright
      const float sc2 = p.scale * 1.4426950408889634f;
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c0 + c < nchunk) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 m4 = *reinterpret_cast<const float4*>(mask_s + (c0 + c) * 32 + j);
            const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float t = fmaf(__uint_as_float(v[c][j + e]), sc2, mm[e]);
              v[c][j + e] = __float_as_uint(t);
              mx4[e] = fmaxf(mx4[e], t);
            }
          }
        }
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      hmax[half * 128 + r] = mx;
      named_bar_sync(1, 256);
      mx = fmaxf(hmax[r], hmax[128 + r]);
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      uint8_t* p_gen = gen + s * Cfg::kSlotBytes;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c0 + c < nchunk) {
          float e[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            e[j] = ex2_approx(__uint_as_float(v[c][j]) - mx);
            sum4[j & 3] += e[j];
          }
          // P[r, chunk] -> K-major 128B-swizzled A-operand tile (64-key block, 16-byte piece ^ (row & 7))
          const int cc = c0 + c;
          uint8_t* prow = p_gen + (cc >> 1) * 16384 + r * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 u;
            u.x = Cvt<T16>::pack2(e[8 * k + 0], e[8 * k + 1]);
            u.y = Cvt<T16>::pack2(e[8 * k + 2], e[8 * k + 3]);
            u.z = Cvt<T16>::pack2(e[8 * k + 4], e[8 * k + 5]);
            u.w = Cvt<T16>::pack2(e[8 * k + 6], e[8 * k + 7]);
            const int piece = ((cc & 1) * 4 + k) ^ (r & 7);
            *reinterpret_cast<uint4*>(prow + piece * 16) = u;
          }
        }
      }
      hsum[half * 128 + r] = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1, 256);
      if (tid == 0) mbar_arrive(bar(P_READY, s));
      const float inv = 1.0f / (hsum[r] + hsum[128 + r]);
      if (kDefer) {
        if (prev_item >= 0) read_out(i - 1, prev_item, prev_inv);
        prev_item = item;
        prev_inv = inv;
      } else {
        read_out(i, item, inv);
      }
    }
    if (kDefer && prev_item >= 0) read_out(i - 1, prev_item, prev_inv);
    if (p.trace && tid == 0) {
      p.trace[blockIdx.x * 16 + 5] = clock64() - t_begin;  // softmax warps: total
      p.trace[blockIdx.x * 16 + 7] = i;                    // items
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Production kernel (S <= 256): ping-pong variant of the pipelined kernel.  Two softmax groups of 4 warps each take
// alternate work items (sample, head, 128-query tile); a thread owns a WHOLE query row, walked 32 score columns at a
// time in two passes straight out of TMEM (max, then exp/sum/P), so there is no cross-thread exchange and no CTA-wide
// barrier in the loop, and while one group is in its MUFU-bound exp phase the other is loading / reading out.
// NCH = 64-key blocks: S <= 128 -> 4 slots, S <= 256 -> 2 slots.  P aliases the dead Q/K tiles, O aliases the first 64
// TMEM columns of S, the 16-bit read-out is staged in the dead P tile and leaves as one bulk tensor store per warp
// (a [B][S][H] map clips the rows of the tile that lie beyond S).
template <int NCH>
struct AttnPPCfg {
  using Base = Attn2Cfg<NCH>;
  static constexpr int kSlots = Base::kSlots;
  static constexpr int kMaskBytes = 8 * NCH * 64 * 4;  // one private copy of the key mask per softmax warp
  static constexpr int kSmemBytes = 1024 + kSlots * Base::kSlotBytes + kMaskBytes + 256;
};

template <typename T16, int NCH>
__global__ void __launch_bounds__(kAttn2Threads, 1) attn_pp_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                   const __grid_constant__ CUtensorMap tmap_ctx,
                                                                   const AttnParams p) {
  using Cfg = Attn2Cfg<NCH>;
  constexpr int NSLOT = Cfg::kSlots;
  constexpr bool kDefer = NSLOT >= 4;
  constexpr int kMaskVec = (NCH * 64 + 127) / 128;  // float4 per lane covering the key mask
  extern __shared__ uint8_t smem_raw[];
  const int S = p.S;
  const int NK = (S + 31) & ~31;          // whole 32-column chunks: every score column read was written by the MMA
  const int nchunk = NK / 32;             // <= 2 * NCH
  const int n_mt = (S + 127) / 128;
  const int n_items = p.B * p.nH * n_mt;
  const int n_mine = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* mask_s = reinterpret_cast<float*>(gen + NSLOT * Cfg::kSlotBytes);
  const uint32_t bars = base + NSLOT * Cfg::kSlotBytes + AttnPPCfg<NCH>::kMaskBytes;
  enum { QK_FULL = 0, V_FULL, S_FULL, P_READY, O_FULL, SLOT_FREE };
  auto bar = [&](int which, int s) { return bars + 8u * (which * NSLOT + s); };
  const uint32_t tmem_slot = bars + 8u * 6 * NSLOT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_ctx);
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(bar(QK_FULL, s), 1);
      mbar_init(bar(V_FULL, s), 1);
      mbar_init(bar(S_FULL, s), 1);
      mbar_init(bar(P_READY, s), 4);
      mbar_init(bar(O_FULL, s), 1);
      mbar_init(bar(SLOT_FREE, s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.qkv_ready == nullptr) pdl_wait();
  if (p.ctx_done != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && ((p.B * S) & 127) != 0)
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.ctx_done + ((p.B * S) >> 7)),
                 "r"((128 - ((p.B * S) & 127)) * p.nH) : "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int i = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++i) {
        const int s = i % NSLOT;
        const uint32_t par = (i / NSLOT) & 1u;
        const int mt = item % n_mt, h = (item / n_mt) % p.nH, b = item / (n_mt * p.nH);
        const int row0 = b * S;
        const uint32_t sQ = base + s * Cfg::kSlotBytes, sK = sQ + 16384, sV = sQ + Cfg::kVOff;
        if (p.qkv_ready != nullptr) {   // the sample's rows, published tile by tile by the chain kernel still running
          for (int t = row0 / 128; t <= (row0 + S - 1) / 128; ++t) attn_rows_wait(p.qkv_ready + t, p.qkv_target);
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        mbar_wait(bar(SLOT_FREE, s), par ^ 1u);
        mbar_expect_tx(bar(QK_FULL, s), 16384 + NCH * 8192);
        tma_load_2d(sQ, &tmap_qkv, bar(QK_FULL, s), h * kAttnDH, row0 + mt * 128);
        tma_load_2d(sQ + 8192, &tmap_qkv, bar(QK_FULL, s), h * kAttnDH, row0 + mt * 128 + 64);
#pragma unroll
        for (int kb = 0; kb < NCH; ++kb)
          tma_load_2d(sK + kb * 8192, &tmap_qkv, bar(QK_FULL, s), p.H + h * kAttnDH, row0 + kb * 64);
        mbar_expect_tx(bar(V_FULL, s), NCH * 8192);
#pragma unroll
        for (int kb = 0; kb < NCH; ++kb)
          tma_load_2d(sV + kb * 8192, &tmap_qkv, bar(V_FULL, s), 2 * p.H + h * kAttnDH, row0 + kb * 64);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: an event loop over the two things
    // it can be asked to do (S = Q K^T of the next loaded item, O = P V of the next finished softmax), so that
    // neither waits behind the other
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, NK, Cvt<T16>::kFmt, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, kAttnDH, Cvt<T16>::kFmt, 0, 1);
      int next_qk = 0, next_pv = 0;
      uint32_t spins = 0;
      while (next_pv < n_mine) {
        bool did = false;
        if (next_qk < n_mine) {
          const int s = next_qk % NSLOT;
          if (mbar_try_wait(bar(QK_FULL, s), (next_qk / NSLOT) & 1u)) {
            tc_fence_after();
            const uint32_t sQ = base + s * Cfg::kSlotBytes;
            const uint64_t qd = make_smem_desc(sQ, 16, 1024), kd = make_smem_desc(sQ + 16384, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + s * Cfg::kTmemStride, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
            umma_commit(bar(S_FULL, s));
            ++next_qk;
            did = true;
          }
        }
        {
          const int s = next_pv % NSLOT;
          const uint32_t par = (next_pv / NSLOT) & 1u;
          if (next_pv < next_qk && mbar_try_wait(bar(P_READY, s), par) && mbar_try_wait(bar(V_FULL, s), par)) {
            tc_fence_after();
            const uint32_t sP = base + s * Cfg::kSlotBytes, sV = sP + Cfg::kVOff;
            for (int k = 0; k < NK / 16; ++k) {
              const uint64_t pd = make_smem_desc(sP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
              const uint64_t vd = make_smem_desc(sV + k * 2048, 1024, 1024);
              umma_f16(tmem_base + s * Cfg::kTmemStride, pd, vd, idesc_o, k != 0);
            }
            umma_commit(bar(O_FULL, s));
            ++next_pv;
            did = true;
          }
        }
        if (did) {
          spins = 0;
        } else if (++spins > (1u << 26)) {
          printf("cpt_b200: attention MMA issuer stalled (block %d, qk %d pv %d of %d)\n", blockIdx.x, next_qk, next_pv,
                 n_mine);
          __trap();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ two softmax groups (4 warps each)
    const int grp = (warp - 2) >> 2, q = warp & 3;
    const int r = q * 32 + lane;             // query row == TMEM lane
    float* gmask = mask_s + (warp - 2) * (NCH * 64);  // this warp's private copy of the item's key mask (no group barrier)
    int pub_g0 = 0, pub_valid = 0;   // context rows of this warp's last store that are not yet counted in ctx_done
    auto publish_ctx = [&](int g0, int valid) {   // after cp.async.bulk.wait_group 0: rows [g0, g0 + valid) are written
      asm volatile("fence.proxy.async.global;" ::: "memory");
      const int in0 = min(valid, 128 - (g0 & 127));
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.ctx_done + (g0 >> 7)), "r"(in0) : "memory");
      if (valid > in0)
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.ctx_done + (g0 >> 7) + 1), "r"(valid - in0) : "memory");
    };
    const bool tr = p.trace != nullptr && warp == 2 && lane == 0;
    long long tq = tr ? clock64() : 0;
    auto lap = [&](int slot) {
      if (tr) {
        const long long now = clock64();
        p.trace[blockIdx.x * 16 + slot] += now - tq;
        tq = now;
      }
    };
    // this thread's mask element of the group's NEXT item is fetched one item ahead (a global load at item start sat
    // on the critical path); pre-multiplied by log2(e): softmax(t) = 2^((t - max) log2 e), exp is a bare ex2.approx
    auto fetch_mask = [&](int i, int kk) -> float4 {
      float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (i < n_mine) {
        const float* src = p.ext_mask + (long long)((blockIdx.x + i * gridDim.x) / (p.nH * n_mt)) * S;
        const int c = (kk * 32 + lane) * 4;
        if (c + 0 < S) m.x = src[c + 0];  // raw values: the scaling happens when they are consumed, one item later,
        if (c + 1 < S) m.y = src[c + 1];  // so nothing waits on these loads
        if (c + 2 < S) m.z = src[c + 2];
        if (c + 3 < S) m.w = src[c + 3];
      }
      return m;
    };
    float4 next_mask[kMaskVec];
#pragma unroll
    for (int kk = 0; kk < kMaskVec; ++kk) next_mask[kk] = fetch_mask(grp, kk);
    int pending_free = -1;  // slot whose read-out store is still draining its smem reads
    const float sc2 = p.scale * 1.4426950408889634f;
    for (int i = grp; i < n_mine; i += 2) {
      const int item = blockIdx.x + i * gridDim.x;
      const int s = i % NSLOT;
      const uint32_t par = (i / NSLOT) & 1u;
      const int mt = item % n_mt, h = (item / n_mt) % p.nH, b = item / (n_mt * p.nH);
#pragma unroll
      for (int kk = 0; kk < kMaskVec; ++kk) {
        if ((kk * 32 + lane) * 4 < NCH * 64) {  // stay inside this warp's private copy
          *reinterpret_cast<float4*>(gmask + (kk * 32 + lane) * 4) =
              make_float4(next_mask[kk].x * 1.4426950408889634f, next_mask[kk].y * 1.4426950408889634f,
                          next_mask[kk].z * 1.4426950408889634f, next_mask[kk].w * 1.4426950408889634f);
          next_mask[kk] = fetch_mask(i + 2, kk);
        }
      }
      __syncwarp();
      lap(0);  // mask
      mbar_wait(bar(S_FULL, s), par);
      tc_fence_after();
      lap(1);  // wait S
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + s * Cfg::kTmemStride;
      // Two passes over the score row, 32 columns at a time, re-reading TMEM (a 32-column read costs ~30 cycles;
      // holding all 128 scores in registers spilled).  t = (s / sqrt(dH) + ext_mask) * log2(e); padded key columns
      // carry mask = -inf and are excluded outright.  Packed fp32x2 FMAs, 4 independent max / sum chains.
      const f32x2 sc22 = pack_f2(sc2, sc2);
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 2 * NCH; ++c) {
        if (c < nchunk) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 m4 = *reinterpret_cast<const float4*>(gmask + c * 32 + j);
            float t0, t1, t2, t3;
            unpack_f2(fma_f2(pack_f2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc22, pack_f2(m4.x, m4.y)), t0, t1);
            unpack_f2(fma_f2(pack_f2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), sc22, pack_f2(m4.z, m4.w)), t2,
                      t3);
            mx4[0] = fmaxf(mx4[0], t0);
            mx4[1] = fmaxf(mx4[1], t1);
            mx4[2] = fmaxf(mx4[2], t2);
            mx4[3] = fmaxf(mx4[3], t3);
          }
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      lap(3);  // scale + mask + max
      const f32x2 nmx2 = pack_f2(-mx, -mx);
      f32x2 sum01 = pack_f2(0.f, 0.f), sum23 = pack_f2(0.f, 0.f);
      uint8_t* p_gen = gen + s * Cfg::kSlotBytes;
#pragma unroll
      for (int c = 0; c < 2 * NCH; ++c) {
        if (c < nchunk) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c * 32, v);
          tmem_ld_wait();
          uint8_t* prow = p_gen + (c >> 1) * 16384 + r * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float e[8];
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
              const float4 m4 = *reinterpret_cast<const float4*>(gmask + c * 32 + 8 * k + j);
              // (v * sc2 + (mask - max)) -> ex2
              float t0, t1, t2, t3;
              unpack_f2(fma_f2(pack_f2(__uint_as_float(v[8 * k + j]), __uint_as_float(v[8 * k + j + 1])), sc22,
                               add_f2(pack_f2(m4.x, m4.y), nmx2)), t0, t1);
              unpack_f2(fma_f2(pack_f2(__uint_as_float(v[8 * k + j + 2]), __uint_as_float(v[8 * k + j + 3])), sc22,
                               add_f2(pack_f2(m4.z, m4.w), nmx2)), t2, t3);
              e[j] = ex2_approx(t0);
              e[j + 1] = ex2_approx(t1);
              e[j + 2] = ex2_approx(t2);
              e[j + 3] = ex2_approx(t3);
              sum01 = add_f2(sum01, pack_f2(e[j], e[j + 1]));
              sum23 = add_f2(sum23, pack_f2(e[j + 2], e[j + 3]));
            }
            uint4 u;
            u.x = Cvt<T16>::pack2(e[0], e[1]);
            u.y = Cvt<T16>::pack2(e[2], e[3]);
            u.z = Cvt<T16>::pack2(e[4], e[5]);
            u.w = Cvt<T16>::pack2(e[6], e[7]);
            const int piece = ((c & 1) * 4 + k) ^ (r & 7);
            *reinterpret_cast<uint4*>(prow + piece * 16) = u;
          }
        }
      }
      float sum4[4];
      unpack_f2(sum01, sum4[0], sum4[1]);
      unpack_f2(sum23, sum4[2], sum4[3]);
      lap(4);  // exp + P -> smem
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(P_READY, s));
      if (kDefer && pending_free >= 0) {  // the previous item's bulk store has long read its staging tile
        if (lane == 0) {
          tma_store_wait_read<0>();
          mbar_arrive(bar(SLOT_FREE, pending_free));
        }
        pending_free = -1;
      }
      lap(5);  // fences + arrive
      const float inv = 1.0f / ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
      mbar_wait(bar(O_FULL, s), par);
      tc_fence_after();
      lap(6);  // wait O
      {
        // O / rowsum -> 16-bit, staged in the slot's dead P tile (this warp's 32 rows x 128 B, 128B-swizzled) and
        // written out by ONE bulk tensor store per warp; the [B][S][H] map clips the tile's rows >= S.
        const uint32_t stage_u32 = base + s * Cfg::kSlotBytes + q * 4096;
        uint8_t* srow = gen + s * Cfg::kSlotBytes + q * 4096 + lane * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(t_row + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 u;
            u.x = Cvt<T16>::pack2(__uint_as_float(o[8 * k + 0]) * inv, __uint_as_float(o[8 * k + 1]) * inv);
            u.y = Cvt<T16>::pack2(__uint_as_float(o[8 * k + 2]) * inv, __uint_as_float(o[8 * k + 3]) * inv);
            u.z = Cvt<T16>::pack2(__uint_as_float(o[8 * k + 4]) * inv, __uint_as_float(o[8 * k + 5]) * inv);
            u.w = Cvt<T16>::pack2(__uint_as_float(o[8 * k + 6]) * inv, __uint_as_float(o[8 * k + 7]) * inv);
            *reinterpret_cast<uint4*>(srow + (((c * 4 + k) ^ (lane & 7)) * 16)) = u;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (p.ctx_done != nullptr && pub_valid > 0) {   // the PREVIOUS item's store has had a whole item to land
            tma_store_wait<0>();
            publish_ctx(pub_g0, pub_valid);
          }
          tma_store_3d(&tmap_ctx, stage_u32, h * kAttnDH, mt * 128 + q * 32, b);
          tma_store_commit();
          pub_g0 = b * S + mt * 128 + q * 32;
          pub_valid = max(0, min(32, S - (mt * 128 + q * 32)));
          if (!kDefer) {  // two slots only: the slot must be refilled as soon as possible
            tma_store_wait_read<0>();
            mbar_arrive(bar(SLOT_FREE, s));
          }
        }
        if (kDefer) pending_free = s;  // released in the middle of this group's next item (or after the loop)
      }
      lap(7);  // read-out + store
      if (tr) p.trace[blockIdx.x * 16 + 8] += 1;
    }
    if (lane == 0) {
      tma_store_wait<0>();
      if (p.ctx_done != nullptr && pub_valid > 0) publish_ctx(pub_g0, pub_valid);
      if (pending_free >= 0) mbar_arrive(bar(SLOT_FREE, pending_free));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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
  pdl_launch_dependents();
  pdl_wait();
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
  pdl_launch_dependents();
  pdl_wait();
  if (i < n) out[i] = (1.0f - (float)mask[i]) * -10000.0f;
}

}  // namespace cptk
