// Persistent warp-specialised tcgen05 GEMM for sm_100a:   D[M,N] = epi(A[M,K] . B[N,K]^T)
//   A, B : 16-bit (fp16 or bf16), K-major, staged by TMA (128B swizzle) through a kStages mbarrier ring
//   D    : fp32 accumulator in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps the
//          main loop of tile i+1
//   roles: warp 0 = TMA producer (1 lane), warp 1 = TMEM owner + MMA issuer (1 lane), warps 2..9 = epilogue
//          (TMEM lane quadrant = warp_idx % 4, two warps per quadrant splitting the BN columns)
//   PAIR : thread-block cluster of 2 CTAs issuing tcgen05.mma.cta_group::2 (UMMA M = 256).  Each CTA stages its
//          own 128 A rows but only HALF of the B tile, so a CTA ingests (128 + BN/2) x 64 operands per k-block
//          instead of (128 + BN) x 64.  Measured on B200 this kernel is bound by per-SM operand ingest
//          (~41.5 B/clk/SM on every shape tried; TMA multicast inside a 2/4-CTA cluster did not move it), so the
//          pair raises the ceiling by the same factor.  The leader CTA (cluster rank 0) issues all MMAs; both
//          CTAs' TMA loads signal the leader's full barrier; stage release / accumulator-ready are multicast
//          tcgen05.commit; both CTAs' epilogues release the accumulator on the leader's barrier.
// Epilogues (fused; the reference runs them as separate ATen ops):
//   EPI_BIAS       : + bias                      (QKV projection, region embedding, vocabulary decoder)
//   EPI_BIAS_GELU  : erf-GELU(+ bias)            (BertIntermediate)
//   EPI_BIAS_RESID : + bias + fp32 residual      (BertSelfOutput / BertOutput dense, pre-LayerNorm)
// The epilogue transposes each 32x32 accumulator block through a per-warp smem pad so global loads (residual)
// and stores are row-contiguous 128-byte (fp32) / 64-byte (16-bit) segments.
#pragma once
#include "ptx.cuh"

namespace cptk {

enum GemmEpilogue { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_RESID = 2 };

struct GemmParams {
  int M, N, K;
  void* out;          // OutT [rows, ldo]
  long long ldo;      // elements
  const float* bias;  // [N] or nullptr
  const float* resid; // fp32 [M, ldr] (EPI_BIAS_RESID)
  long long ldr;
  int tma_store;        // outputs leave through tmap_out (set by the host when `out` is 16-byte aligned and pitched)
  int tma_reduce;       // ... as out += tile (cp.reduce.async.bulk .add at L2): `out` already holds the residual
  long long* trace;     // optional [gridDim.x][16] cycle counters (debug): see cpt_gemm_trace
  // trans bit 0: A is given TRANSPOSED in memory, as [K, M] row-major; bit 1: W is given as [K, N] row-major.
  //   3: out = A^T . W  — the weight-gradient product dW = dY^T X straight from the row-major activations;
  //   2: out = A . W    — the data-gradient product dX = dY W straight from the nn.Linear weight [out, in].
  // A transposed operand's tiles are staged as [64-wide MN block][64 k rows][128 B] and read through MN-major UMMA
  // descriptors.  Single-CTA tiles only.
  int trans;
  // ksplit > 1: the K range is cut into ksplit pieces handled as separate work items that all ADD their partial
  // product into `out` (needs tma_reduce; the bias rides with the first piece): fills the machine when M x N is a
  // handful of tiles and K is long
  // (weight gradients: 768 x 768 outputs over K = 7680 rows)
  int ksplit;
  // ---- LayerNorm folding (see DESIGN.md "LayerNorm folding"); all optional (nullptr = off)
  // EPI_BIAS / EPI_BIAS_GELU: the A operand is a PRE-LayerNorm tensor x (16-bit) and W already carries gamma;
  //   out = rstd_m * (acc - mu_m * gvec_n) + bias_n, with (mu, rstd) from nstats[m] = (sum x, sum x^2) over nH
  const float* nstats;
  const float* gvec;
  float neps;
  int nH;
  // EPI_BIAS_RESID: the residual is LN(resid_raw) applied on the fly from rstats/rgamma/rbeta; the sum x is also
  //   written as 16 bits (next GEMM's A operand) and its row statistics accumulated into stats_out (atomicAdd)
  const float* rstats;
  const float* rgamma;
  const float* rbeta;
  float reps;
  void* out16;
  long long ldo16;
  float* stats_out;
};

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;  // 64 x 2 B = one 128-byte swizzle row
constexpr int kGemmEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kGemmEpiWarps;
constexpr int kSmemLimit = 232448;  // 227 KB

template <int BN, int OutBytes, bool PAIR, int NVEC = 3>
struct GemmCfg {
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN must be 64, 128, 192 or 256");
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;
  static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kGemmBK * 2;  // per CTA
  static constexpr int kStageBytes = kABytes + kBBytes;
  // per epilogue warp: a staging area for one 32x32 output block.  TMA-store path: the block in the TMA box layout
  // (fp32: one 4 KB buffer, 128B swizzle; 16-bit: one 2 KB buffer, 64B swizzle).  LSU path (unaligned outputs,
  // residual epilogue): a padded transpose area (pitch 144 B / 80 B).
  static constexpr int kPadPitch = (OutBytes == 4) ? 144 : 80;
  static constexpr int kStageEpi = (OutBytes == 4) ? 5120 : 3072;     // multiple of 1024, >= 32 * kPadPitch
  static constexpr int kBiasBytes = NVEC * (BN / 2) * 4;              // per epilogue warp: NVEC column vectors
  static constexpr int kEpiBytes = kGemmEpiWarps * (kStageEpi + kBiasBytes);
  static constexpr int kFixed = 1024 /*align slack*/ + 256 /*barriers*/ + kEpiBytes;
  static constexpr int kStagesFit = (kSmemLimit - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixed;
  static_assert(kStages >= 3, "pipeline too shallow");
};

// erf-form GELU, x * Phi(x) = x * 0.5 * (1 + erf(x / sqrt 2)) (hidden_act == "gelu"), evaluated as
//     x * sigmoid(p(x)),  p(x) = x (c0 + c1 x^2 + c2 x^4)  = a minimax fit of logit(Phi(x)) on |x| <= 6
// (argument clamped to +-6, where Phi is 0 / 1 to 1e-9).  Max abs error vs the exact erf form 2.6e-5 in fp32
// (tools/fit_gelu.py), i.e. ~5x below the 16-bit rounding of the value it produces; 10 instructions and two
// MUFU ops per element instead of ~20 for erff(): the FFN-up epilogue must retire 128 x 256 activations per
// 6144 tensor-core cycles.  The coefficients below are c_i * -log2(e) so the exponential is a bare ex2.
__device__ __forceinline__ float gelu_exp_arg(float x) {  // q(x): gelu(x) = x / (1 + 2^q)
  const float xc = fminf(fmaxf(x, -6.0f), 6.0f);
  const float u = xc * xc;
  float q = fmaf(u, 0.0010142651153728366f, -0.10677573829889297f);
  q = fmaf(u, q, -2.301121234893799f);
  return q * xc;
}
__device__ __forceinline__ float gelu_erf(float x) { return x * rcp_approx(1.0f + ex2_approx(gelu_exp_arg(x))); }
// Four at once with ONE reciprocal: the epilogue is MUFU-bound (2 MUFU ops per element at 4 lanes/clk/SMSP), and
// 1/a, 1/b, 1/c, 1/d = r*(b*cd), r*(a*cd), r*(ab*d), r*(ab*c) with r = 1/(ab*cd) trades 3 MUFU.RCP for FMULs.
// Each denominator is 1 + 2^q <= 1 + 2^29 (q is clamped), so the product stays far below 2^127.  The polynomial
// and the products run as packed fp32x2 instructions (FFMA2): half the FMA-pipe issue slots.
__device__ __forceinline__ void gelu_erf4(float& x0, float& x1, float& x2, float& x3) {
  const float c0 = fminf(fmaxf(x0, -6.0f), 6.0f), c1 = fminf(fmaxf(x1, -6.0f), 6.0f);
  const float c2 = fminf(fmaxf(x2, -6.0f), 6.0f), c3 = fminf(fmaxf(x3, -6.0f), 6.0f);
  const f32x2 xc01 = pack_f2(c0, c1), xc23 = pack_f2(c2, c3);
  const f32x2 k2 = pack_f2(0.0010142651153728366f, 0.0010142651153728366f);
  const f32x2 k1 = pack_f2(-0.10677573829889297f, -0.10677573829889297f);
  const f32x2 k0 = pack_f2(-2.301121234893799f, -2.301121234893799f);
  const f32x2 u01 = mul_f2(xc01, xc01), u23 = mul_f2(xc23, xc23);
  const f32x2 q01 = mul_f2(fma_f2(u01, fma_f2(u01, k2, k1), k0), xc01);
  const f32x2 q23 = mul_f2(fma_f2(u23, fma_f2(u23, k2, k1), k0), xc23);
  float q0, q1, q2, q3;
  unpack_f2(q01, q0, q1);
  unpack_f2(q23, q2, q3);
  const float a = 1.0f + ex2_approx(q0), b = 1.0f + ex2_approx(q1);
  const float c = 1.0f + ex2_approx(q2), d = 1.0f + ex2_approx(q3);
  float ab, cd;
  unpack_f2(mul_f2(pack_f2(a, c), pack_f2(b, d)), ab, cd);
  const float r = rcp_approx(ab * cd);
  float rab, rcd;
  unpack_f2(mul_f2(pack_f2(r, r), pack_f2(ab, cd)), rab, rcd);
  unpack_f2(mul_f2(mul_f2(pack_f2(x0, x1), pack_f2(rcd, rcd)), pack_f2(b, a)), x0, x1);
  unpack_f2(mul_f2(mul_f2(pack_f2(x2, x3), pack_f2(rab, rab)), pack_f2(d, c)), x2, x3);
}

template <int BN, bool PAIR, int EPI, typename OutT, typename T16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_out, const GemmParams p) {
  using Cfg = GemmCfg<BN, (int)sizeof(OutT), PAIR, (EPI == EPI_BIAS_RESID) ? 3 : 2>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kCluster = PAIR ? 2 : 1;
  static_assert(!PAIR || BN % 32 == 0, "cta_group::2 needs N % 16 == 0 and whole swizzle atoms per half");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + kStages * Cfg::kStageBytes + Cfg::kEpiBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  uint8_t* epi_gen = smem_gen + kStages * Cfg::kStageBytes;  // 1024-aligned: the TMA-store staging needs it

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + kGemmBM - 1) / kGemmBM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int ct_m = (m_tiles + kCluster - 1) / kCluster;
  const int mn_ctiles = ct_m * n_tiles;  // cluster tiles: (kCluster*128) x BN
  const int num_kb = (p.K + kGemmBK - 1) / kGemmBK;
  const int ksplit = p.ksplit > 1 ? p.ksplit : 1;
  const int num_ctiles = mn_ctiles * ksplit;  // work items: (tile, K piece); piece index = ct / mn_ctiles
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  const int cluster_id = blockIdx.x / kCluster, num_clusters = gridDim.x / kCluster;

  long long t_entry = 0, g_entry = 0;
  if (p.trace && threadIdx.x == 0) {
    t_entry = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_entry));
  }
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), kCluster * kGemmEpiWarps);
      }
      fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) {
      tmem_alloc_2cta(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 16 + 8] = clock64() - t_entry;   // prologue (before PDL wait)
  pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x * 16 + 9] = clock64() - t_entry;   // ... incl. the PDL wait
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long t_wait = 0, t_start = clock64();
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
        const int tile = ct % mn_ctiles, ks = ct / mn_ctiles;
        const int m0 = ((tile / n_tiles) * kCluster + (int)crank) * kGemmBM;
        const int n0 = (tile % n_tiles) * BN;
        const int kb_end = (int)((long long)(ks + 1) * num_kb / ksplit);
        for (int kb = (int)((long long)ks * num_kb / ksplit); kb < kb_end; ++kb) {
          const long long t0 = clock64();
          mbar_wait(empty_bar(stage), phase ^ 1u);
          t_wait += clock64() - t0;
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (!PAIR && p.trans) {
            mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
            if (p.trans & 1) {
#pragma unroll
              for (int blk = 0; blk < kGemmBM / 64; ++blk)
                tma_load_2d(sa + blk * 8192, &tmap_a, full_bar(stage), m0 + blk * 64, kb * kGemmBK);
            } else {
              tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kGemmBK, m0);
            }
            if (p.trans & 2) {
#pragma unroll
              for (int blk = 0; blk < BN / 64; ++blk)
                tma_load_2d(sb + blk * 8192, &tmap_b, full_bar(stage), n0 + blk * 64, kb * kGemmBK);
            } else {
              tma_load_2d(sb, &tmap_b, full_bar(stage), kb * kGemmBK, n0);
            }
          } else if (!PAIR) {
            mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kGemmBK, m0);
            tma_load_2d(sb, &tmap_b, full_bar(stage), kb * kGemmBK, n0);
          } else {
            // both CTAs' bytes are counted on the leader's barrier (the leader alone issues the MMAs)
            if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
            const uint32_t lbar = mapa_cluster(full_bar(stage), 0);
            tma_load_2d_2cta(sa, &tmap_a, lbar, kb * kGemmBK, m0);
            tma_load_2d_2cta(sb, &tmap_b, lbar, kb * kGemmBK, n0 + (int)crank * (BN / 2));
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (p.trace) {
        p.trace[blockIdx.x * 16 + 0] = clock64() - t_start;  // producer: total
        p.trace[blockIdx.x * 16 + 1] = t_wait;               // producer: waiting for a free stage
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only when PAIR)
    if (lane == 0 && leader) {
      const bool ta = !PAIR && (p.trans & 1), tb = !PAIR && (p.trans & 2);
      const uint32_t idesc = make_idesc_f16(kGemmBM * kCluster, BN, Cvt<T16>::kFmt, ta ? 1 : 0, tb ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long t_full = 0, t_tmem = 0, t_start = clock64();
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        long long t0 = clock64();
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        t_tmem += clock64() - t0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int ks = ct / mn_ctiles;
        const int kb_begin = (int)((long long)ks * num_kb / ksplit), kb_end = (int)((long long)(ks + 1) * num_kb / ksplit);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          t0 = clock64();
          mbar_wait(full_bar(stage), phase);
          t_full += clock64() - t0;
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          // K-major: +32 B per UMMA_K=16 step inside the 128-byte swizzle row (descriptor address is in 16-B units);
          // MN-major: 64-wide MN blocks 8 KB apart (LBO), 8-row k groups 1 KB apart (SBO), +2 KB per 16 k rows
          const uint64_t adesc = ta ? make_smem_desc(sa, 8192, 1024) : make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc =
              tb ? make_smem_desc(sa + Cfg::kABytes, 8192, 1024) : make_smem_desc(sa + Cfg::kABytes, 16, 1024);
          const uint32_t ka = ta ? 128u : 2u, kbs = tb ? 128u : 2u;
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            if (PAIR) umma_f16_2cta(d_tmem, adesc + ka * k, bdesc + kbs * k, idesc, ((kb - kb_begin) | k) != 0);
            else umma_f16(d_tmem, adesc + ka * k, bdesc + kbs * k, idesc, ((kb - kb_begin) | k) != 0);
          }
          if (PAIR) umma_commit_2cta_mc(empty_bar(stage), 3);
          else umma_commit(empty_bar(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (PAIR) umma_commit_2cta_mc(tfull_bar(acc), 3);
        else umma_commit(tfull_bar(acc));
      }
      if (p.trace) {
        p.trace[blockIdx.x * 16 + 2] = clock64() - t_start;  // MMA issuer: total
        p.trace[blockIdx.x * 16 + 3] = t_full;               // ... waiting for operands (TMA)
        p.trace[blockIdx.x * 16 + 4] = t_tmem;               // ... waiting for a drained accumulator (epilogue)
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp - 2;
    const int q = warp & 3;        // TMEM lane quadrant this warp may read
    const int half = ew >> 2;      // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    uint8_t* pad = epi_gen + ew * Cfg::kStageEpi;
    const uint32_t pad_u32 = smem_base + kStages * Cfg::kStageBytes + ew * Cfg::kStageEpi;
    float* sv0 = reinterpret_cast<float*>(epi_gen + kGemmEpiWarps * Cfg::kStageEpi + ew * Cfg::kBiasBytes);  // bias / c_n
    float* sv1 = sv0 + kColsPerWarp;                              // gvec  | residual-LN gamma
    float* sv2 = sv1 + kColsPerWarp;                              //       | residual-LN beta (EPI_BIAS_RESID only)
    constexpr bool kResid = (EPI == EPI_BIAS_RESID);
    const bool vec_ok = ((p.ldo * (long long)sizeof(OutT)) % 16 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const bool norm = !kResid && p.nstats != nullptr;
    const bool tma_out = !kResid && p.tma_store != 0;  // aligned output: async TMA stores instead of LSU stores
    int n_stores = 0;                                  // TMA stores issued by this warp (16-bit: 2 buffers)
    const bool rnorm = kResid && p.rstats != nullptr;
    // this warp's slice of a per-column vector -> smem (zero beyond N)
    auto stage_vec = [&](float* dst, const float* src, int n_first) {
      const int nb = n_first + lane * 4;
      if (lane * 4 < kColsPerWarp) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src != nullptr) {
          if (nb + 3 < p.N) {
            b4 = __ldg(reinterpret_cast<const float4*>(src + nb));
          } else {
            if (nb < p.N) b4.x = __ldg(src + nb);
            if (nb + 1 < p.N) b4.y = __ldg(src + nb + 1);
            if (nb + 2 < p.N) b4.z = __ldg(src + nb + 2);
          }
        }
        *reinterpret_cast<float4*>(dst + lane * 4) = b4;
      }
    };
    int it = 0;
    long long t_acc = 0, t_start = clock64();
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int tile = ct % mn_ctiles;
      const int m0 = ((tile / n_tiles) * kCluster + (int)crank) * kGemmBM;
      const int n0 = (tile % n_tiles) * BN;
      const int mrow0 = m0 + q * 32;  // first row of this warp's 32-row band
      const int ncol0 = n0 + half * kColsPerWarp;

      // per-column vectors -> smem while the tile's MMAs are still running (a global load per chunk sat on the
      // critical path of every chunk: ncu long_scoreboard on the bias FADDs)
      stage_vec(sv0, (ct / mn_ctiles) == 0 ? p.bias : nullptr, ncol0);  // split-K: the first K piece carries the bias
      if (norm) stage_vec(sv1, p.gvec, ncol0);
      if (kResid && rnorm) {
        stage_vec(sv1, p.rgamma, ncol0);
        stage_vec(sv2, p.rbeta, ncol0);
      }
      __syncwarp();
      // folded LayerNorm of the A operand: this thread's row (TMEM lane) statistics
      float mu = 0.f, rstd = 1.f;
      if (norm) {
        const int m = min(mrow0 + lane, p.M - 1);
        const float2 st = __ldg(reinterpret_cast<const float2*>(p.nstats) + m);
        mu = st.x / (float)p.nH;
        rstd = rsqrtf(fmaxf(st.y / (float)p.nH - mu * mu, 0.f) + p.neps);
      }
      // on-the-fly LayerNorm of the residual: statistics of the 8 rows this thread touches after the transpose
      float rmu[kResid ? 8 : 1], rrs[kResid ? 8 : 1], acc_s[kResid ? 8 : 1], acc_q[kResid ? 8 : 1];
      if (kResid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rmu[i] = 0.f; rrs[i] = 1.f; acc_s[i] = 0.f; acc_q[i] = 0.f;
          if (rnorm) {
            const int m = min(mrow0 + i * 4 + (lane >> 3), p.M - 1);
            const float2 st = __ldg(reinterpret_cast<const float2*>(p.rstats) + m);
            rmu[i] = st.x / (float)p.nH;
            rrs[i] = rsqrtf(fmaxf(st.y / (float)p.nH - rmu[i] * rmu[i], 0.f) + p.reps);
          }
        }
      }
      const long long t0 = clock64();
      mbar_wait(tfull_bar(acc), acc_phase);
      t_acc += clock64() - t0;
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN + half * kColsPerWarp;
      if (mrow0 < p.M) {
        constexpr int NC = kColsPerWarp / 32;
        // chunk c+1's TMEM load and residual rows are issued as soon as chunk c sits in the transpose pad
        uint32_t rbuf[32];
        float4 res[2][kResid ? 8 : 1];
        const bool res_vec = kResid && (p.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.resid) & 15) == 0);
        auto prefetch_resid = [&](int c, float4* dst) {
          const int nc = ncol0 + c * 32;
          if (kResid && res_vec && nc + 32 <= p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int m = mrow0 + i * 4 + (lane >> 3), n = nc + (lane & 7) * 4;
              dst[i] = (m < p.M) ? *reinterpret_cast<const float4*>(p.resid + (long long)m * p.ldr + n)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        };
        tmem_ld_32x32b_x32(t_row, rbuf);
        prefetch_resid(0, res[0]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int nc = ncol0 + c * 32;
          const bool live = nc < p.N;           // warp-uniform
          const bool full = (nc + 32 <= p.N);
          const long long tp0 = p.trace ? clock64() : 0;
          tmem_ld_wait();
          const long long tp1 = p.trace ? clock64() : 0;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sv0 + c * 32 + j);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
            if (norm) {
              const float4 g4 = *reinterpret_cast<const float4*>(sv1 + c * 32 + j);
              const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                v[j + e] = fmaf(rstd, fmaf(-mu, gg[e], __uint_as_float(rbuf[j + e])), bb[e]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[j + e] = __uint_as_float(rbuf[j + e]) + bb[e];
            }
          }
          if (EPI == EPI_BIAS_GELU) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) gelu_erf4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (tma_out) {
            // ---- async path: the 32x32 block goes to smem in the TMA box layout (row = this thread; 16-byte pieces
            // XOR-swizzled so the quarter-warp stores are bank-conflict free) and one lane issues a bulk tensor
            // store.  The LSU store loop this replaces was the epilogue's critical path (cpt_gemm_trace: ~1100
            // cycles per block waiting on global stores while TMA loads saturate the memory system).
            constexpr int kBufBytes = 32 * 32 * (int)sizeof(OutT);
            constexpr int kNBuf = 1;  // one staging buffer: a second one costs a pipeline stage and buys nothing
            const int buf = n_stores % kNBuf;
            if (n_stores >= kNBuf) {  // the store that last read this buffer must have drained its smem reads
              if (lane == 0) {
                if (kNBuf == 1) tma_store_wait_read<0>();
                else tma_store_wait_read<1>();
              }
              __syncwarp();
            }
            if (c + 1 < NC) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, rbuf);
            if (live) {
              uint8_t* brow = pad + buf * kBufBytes + lane * (32 * (int)sizeof(OutT));
              if (sizeof(OutT) == 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *reinterpret_cast<float4*>(brow + ((j ^ (lane & 7)) * 16)) =
                      make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 u;
                  u.x = Cvt<T16>::pack2(v[8 * j + 0], v[8 * j + 1]);
                  u.y = Cvt<T16>::pack2(v[8 * j + 2], v[8 * j + 3]);
                  u.z = Cvt<T16>::pack2(v[8 * j + 4], v[8 * j + 5]);
                  u.w = Cvt<T16>::pack2(v[8 * j + 6], v[8 * j + 7]);
                  *reinterpret_cast<uint4*>(brow + ((j ^ ((lane >> 1) & 3)) * 16)) = u;
                }
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (sizeof(OutT) == 4 && p.tma_reduce) tma_reduce_add_2d(&tmap_out, pad_u32 + buf * kBufBytes, nc, mrow0);
                else tma_store_2d(&tmap_out, pad_u32 + buf * kBufBytes, nc, mrow0);
                tma_store_commit();
              }
              ++n_stores;
            }
          } else if (sizeof(OutT) == 4) {
            // ---- LSU path: transpose through the per-warp pad: thread = row -> thread = (row group, 16-byte piece)
            if (live) {
              float* prow = reinterpret_cast<float*>(pad + lane * Cfg::kPadPitch);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(prow + 4 * j) =
                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            __syncwarp();
            if (c + 1 < NC) {
              tmem_ld_32x32b_x32(t_row + (c + 1) * 32, rbuf);
              prefetch_resid(c + 1, res[(c + 1) & 1]);
            }
            if (live) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {  // 4 rows x 128 B per warp instruction
                const int rr = i * 4 + (lane >> 3), cc = (lane & 7) * 4;
                const int m = mrow0 + rr, n = nc + cc;
                float4 x = *reinterpret_cast<const float4*>(pad + rr * Cfg::kPadPitch + cc * 4);
                if (m < p.M && n < p.N) {
                  float* o = reinterpret_cast<float*>(p.out) + (long long)m * p.ldo + n;
                  if (full && vec_ok && (!kResid || res_vec)) {
                    if (kResid) {
                      float4 r4 = res[c & 1][i];
                      if (rnorm) {
                        const float4 g4 = *reinterpret_cast<const float4*>(sv1 + c * 32 + cc);
                        const float4 e4 = *reinterpret_cast<const float4*>(sv2 + c * 32 + cc);
                        const float a = rrs[i], b = -rmu[i] * rrs[i];
                        r4.x = fmaf(fmaf(r4.x, a, b), g4.x, e4.x);
                        r4.y = fmaf(fmaf(r4.y, a, b), g4.y, e4.y);
                        r4.z = fmaf(fmaf(r4.z, a, b), g4.z, e4.z);
                        r4.w = fmaf(fmaf(r4.w, a, b), g4.w, e4.w);
                      }
                      x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
                      if (p.stats_out != nullptr) {
                        acc_s[i] += (x.x + x.y) + (x.z + x.w);
                        acc_q[i] += fmaf(x.x, x.x, x.y * x.y) + fmaf(x.z, x.z, x.w * x.w);
                      }
                      if (p.out16 != nullptr) {
                        uint2 u;
                        u.x = Cvt<T16>::pack2(x.x, x.y);
                        u.y = Cvt<T16>::pack2(x.z, x.w);
                        *reinterpret_cast<uint2*>(reinterpret_cast<T16*>(p.out16) + (long long)m * p.ldo16 + n) = u;
                      }
                    }
                    *reinterpret_cast<float4*>(o) = x;
                  } else {
                    const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                      if (n + e < p.N) {
                        float y = xs[e];
                        if (kResid) y += p.resid[(long long)m * p.ldr + n + e];
                        o[e] = y;
                      }
                  }
                }
              }
            }
          } else {
            if (live) {
              uint32_t* prow = reinterpret_cast<uint32_t*>(pad + lane * Cfg::kPadPitch);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = Cvt<T16>::pack2(v[8 * j + 0], v[8 * j + 1]);
                u.y = Cvt<T16>::pack2(v[8 * j + 2], v[8 * j + 3]);
                u.z = Cvt<T16>::pack2(v[8 * j + 4], v[8 * j + 5]);
                u.w = Cvt<T16>::pack2(v[8 * j + 6], v[8 * j + 7]);
                *reinterpret_cast<uint4*>(prow + 4 * j) = u;
              }
            }
            __syncwarp();
            const long long tp2 = p.trace ? clock64() : 0;
            if (c + 1 < NC) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, rbuf);
            if (p.trace && ew == 0 && lane == 0) {
              p.trace[blockIdx.x * 16 + 13] += tp1 - tp0;
              p.trace[blockIdx.x * 16 + 14] += tp2 - tp1;
            }
            if (live) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {  // 8 rows x 64 B per warp instruction
                const int rr = i * 8 + (lane >> 2), cc = (lane & 3) * 8;
                const int m = mrow0 + rr, n = nc + cc;
                const uint4 u = *reinterpret_cast<const uint4*>(pad + rr * Cfg::kPadPitch + cc * 2);
                if (m < p.M && n < p.N) {
                  T16* o = reinterpret_cast<T16*>(p.out) + (long long)m * p.ldo + n;
                  if (full && vec_ok) {
                    *reinterpret_cast<uint4*>(o) = u;
                  } else {
                    const T16* us = reinterpret_cast<const T16*>(&u);
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                      if (n + e < p.N) o[e] = us[e];
                  }
                }
              }
            }
            if (p.trace && ew == 0 && lane == 0) p.trace[blockIdx.x * 16 + 15] += clock64() - tp2;
          }
          __syncwarp();
        }
        tmem_ld_wait();
        if (kResid && p.stats_out != nullptr) {
          // row statistics of x over this warp's columns: 8 lanes share a row -> shuffle-reduce, one atomic pair
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float s = acc_s[i], qq = acc_q[i];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              s += __shfl_xor_sync(0xffffffffu, s, o);
              qq += __shfl_xor_sync(0xffffffffu, qq, o);
            }
            const int m = mrow0 + i * 4 + (lane >> 3);
            if ((lane & 7) == 0 && m < p.M) {
              atomicAdd(p.stats_out + 2 * (long long)m, s);
              atomicAdd(p.stats_out + 2 * (long long)m + 1, qq);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!PAIR || leader) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_cluster(mapa_cluster(tempty_bar(acc), 0));
      }
    }
    if (tma_out && lane == 0) tma_store_wait<0>();  // all bulk stores of this warp have completed
    if (p.trace && ew == 0 && lane == 0) {
      p.trace[blockIdx.x * 16 + 5] = clock64() - t_start;  // epilogue warp 0: total
      p.trace[blockIdx.x * 16 + 6] = t_acc;                // ... waiting for a finished accumulator (MMA)
      p.trace[blockIdx.x * 16 + 7] = it;                   // tiles processed by this CTA
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no CTA exits (or frees TMEM) while its peer may still signal it / read its smem
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
  if (p.trace && threadIdx.x == 0) {
    long long g_exit;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_exit));
    p.trace[blockIdx.x * 16 + 10] = clock64() - t_entry;  // CTA lifetime (cycles)
    p.trace[blockIdx.x * 16 + 11] = g_entry;              // globaltimer at entry / exit (ns)
    p.trace[blockIdx.x * 16 + 12] = g_exit;
  }
}

}  // namespace cptk
