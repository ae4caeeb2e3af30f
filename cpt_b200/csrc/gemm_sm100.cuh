// Persistent warp-specialised tcgen05 GEMM for sm_100a:   D[M,N] = epi(A[M,K] . B[N,K]^T)
//   A, B : 16-bit (fp16 or bf16), K-major, staged by TMA (128B swizzle) through a kStages mbarrier ring
//   D    : fp32 accumulator in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps the
//          main loop of tile i+1
//   roles: warp 0 = TMA producer (1 lane), warp 1 = TMEM owner + MMA issuer (1 lane), warps 2..9 = epilogue
//          (TMEM lane quadrant = warp_idx % 4, two warps per quadrant splitting the BN columns)
//   cluster CM x CN (thread-block cluster of CM*CN CTAs): the CM CTAs that share an N tile each fetch 1/CM of the B
//          tile and TMA-multicast it to the others; the CN CTAs that share an M tile do the same with A.  The
//          kernel is bound by L2->SM operand bandwidth at these shapes (K = 768), so every multicast halves the
//          bytes one operand costs.  Stage release is a multicast tcgen05.commit to every CTA that writes into
//          this CTA's ring.
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
  int rin, rout, roff;  // output row remap: (m / rin) * rout + roff + m % rin   (identity when rin == 0)
};

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;  // 64 x 2 B = one 128-byte swizzle row
constexpr int kGemmEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kGemmEpiWarps;
constexpr int kSmemLimit = 232448;  // 227 KB

template <int BN, int OutBytes>
struct GemmCfg {
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN must be 64, 128, 192 or 256");
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kPadPitch = (OutBytes == 4) ? 144 : 80;        // bytes per staged row (32 outputs + 16 B)
  static constexpr int kPadBytes = 32 * kPadPitch;                    // per epilogue warp
  static constexpr int kEpiBytes = kGemmEpiWarps * kPadBytes;
  static constexpr int kFixed = 1024 /*align slack*/ + 256 /*barriers*/ + kEpiBytes;
  static constexpr int kStagesFit = (kSmemLimit - kFixed) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixed;
  static_assert(kStages >= 3, "pipeline too shallow");
};

// erf-form GELU, x * 0.5 * (1 + erf(x / sqrt 2)).  erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7), which is
// two orders below the 16-bit rounding of the value it feeds and ~2x cheaper than erff() in the epilogue.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  p *= t;
  const float e = fmaf(-p, __expf(-z * z), 1.0f);  // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

template <int BN, int CM, int CN, int EPI, typename OutT, typename T16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const GemmParams p) {
  using Cfg = GemmCfg<BN, (int)sizeof(OutT)>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kCluster = CM * CN;
  static_assert((kGemmBM / CN) % 8 == 0 && (BN / CM) % 8 == 0, "multicast slices must be whole swizzle atoms");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + kStages * Cfg::kStageBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  uint8_t* epi_gen = smem_gen + kStages * Cfg::kStageBytes + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + kGemmBM - 1) / kGemmBM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int ct_m = (m_tiles + CM - 1) / CM, ct_n = (n_tiles + CN - 1) / CN;
  const int num_ctiles = ct_m * ct_n;  // cluster tiles: (CM*128) x (CN*BN)
  const int num_kb = (p.K + kGemmBK - 1) / kGemmBK;
  const uint32_t crank = (kCluster > 1) ? cluster_ctarank() : 0u;
  const int rm = crank % CM, rn = crank / CM;
  const int cluster_id = blockIdx.x / kCluster, num_clusters = gridDim.x / kCluster;
  uint16_t mask_a = 0, mask_b = 0;  // CTAs sharing my A tile (same rm) / my B tile (same rn)
#pragma unroll
  for (int j = 0; j < CN; ++j) mask_a |= uint16_t(1u << (rm + j * CM));
#pragma unroll
  for (int i = 0; i < CM; ++i) mask_b |= uint16_t(1u << (i + rn * CM));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), CM + CN - 1);  // every CTA whose ring my multicasts land in releases the stage
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), kGemmEpiWarps);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
        const int m0 = ((ct / ct_n) * CM + rm) * kGemmBM;
        const int n0 = ((ct % ct_n) * CN + rn) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
          if (CN == 1) {
            tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kGemmBK, m0);
          } else {
            constexpr int rows = kGemmBM / CN;
            tma_load_2d_mc(sa + rn * rows * 128, &tmap_a, full_bar(stage), kb * kGemmBK, m0 + rn * rows, mask_a);
          }
          if (CM == 1) {
            tma_load_2d(sb, &tmap_b, full_bar(stage), kb * kGemmBK, n0);
          } else {
            constexpr int rows = BN / CM;
            tma_load_2d_mc(sb + rm * rows * 128, &tmap_b, full_bar(stage), kb * kGemmBK, n0 + rm * rows, mask_b);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kGemmBM, BN, Cvt<T16>::kFmt, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sa + Cfg::kABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            // +32 B per UMMA_K=16 step inside the 128-byte swizzle row (descriptor address is in 16-B units)
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          if (kCluster == 1) umma_commit(empty_bar(stage));
          else umma_commit_mc(empty_bar(stage), uint16_t(mask_a | mask_b));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp - 2;
    const int q = warp & 3;        // TMEM lane quadrant this warp may read
    const int half = ew >> 2;      // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    uint8_t* pad = epi_gen + ew * Cfg::kPadBytes;
    const bool vec_ok = ((p.ldo * (long long)sizeof(OutT)) % 16 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    int it = 0;
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int m0 = ((ct / ct_n) * CM + rm) * kGemmBM;
      const int n0 = ((ct % ct_n) * CN + rn) * BN;
      const int mrow0 = m0 + q * 32;  // first row of this warp's 32-row band

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN + half * kColsPerWarp;
      if (mrow0 < p.M) {
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp / 32; ++c) {
          const int nc = n0 + half * kColsPerWarp + c * 32;
          if (nc >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          const bool full = (nc + 32 <= p.N);
          if (p.bias != nullptr) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nc + j < p.N) v[j] += __ldg(p.bias + nc + j);
            }
          }
          if (EPI == EPI_BIAS_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          }
          // ---- transpose through the per-warp pad: thread = row  ->  thread = (row group, 16-byte column chunk)
          if (sizeof(OutT) == 4) {
            float* prow = reinterpret_cast<float*>(pad + lane * Cfg::kPadPitch);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(prow + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {  // 4 rows x 128 B per warp instruction
              const int rr = i * 4 + (lane >> 3), cc = (lane & 7) * 4;
              const int m = mrow0 + rr, n = nc + cc;
              float4 x = *reinterpret_cast<const float4*>(pad + rr * Cfg::kPadPitch + cc * 4);
              if (m < p.M && n < p.N) {
                long long orow = m;
                if (p.rin > 0) orow = (long long)(m / p.rin) * p.rout + p.roff + (m % p.rin);
                float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + n;
                if (full && vec_ok) {
                  if (EPI == EPI_BIAS_RESID) {
                    const float4 r4 = *reinterpret_cast<const float4*>(p.resid + (long long)m * p.ldr + n);
                    x.x += r4.x; x.y += r4.y; x.z += r4.z; x.w += r4.w;
                  }
                  *reinterpret_cast<float4*>(o) = x;
                } else {
                  const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (n + e < p.N) {
                      float y = xs[e];
                      if (EPI == EPI_BIAS_RESID) y += p.resid[(long long)m * p.ldr + n + e];
                      o[e] = y;
                    }
                }
              }
            }
          } else {
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
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // 8 rows x 64 B per warp instruction
              const int rr = i * 8 + (lane >> 2), cc = (lane & 3) * 8;
              const int m = mrow0 + rr, n = nc + cc;
              const uint4 u = *reinterpret_cast<const uint4*>(pad + rr * Cfg::kPadPitch + cc * 2);
              if (m < p.M && n < p.N) {
                long long orow = m;
                if (p.rin > 0) orow = (long long)(m / p.rin) * p.rout + p.roff + (m % p.rin);
                T16* o = reinterpret_cast<T16*>(p.out) + orow * p.ldo + n;
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
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // no CTA exits while a peer may still multicast into / signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace cptk
