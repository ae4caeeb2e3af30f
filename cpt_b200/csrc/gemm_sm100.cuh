// Persistent warp-specialised tcgen05 GEMM for sm_100a:   D[M,N] = epi(A[M,K] . B[N,K]^T)
//   A, B : 16-bit (fp16 or bf16), K-major, staged by TMA (128B swizzle) through a kStages mbarrier ring
//   D    : fp32 accumulator in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps the
//          main loop of tile i+1
//   roles: warp 0 = TMA producer (1 lane), warp 1 = TMEM owner + MMA issuer (1 lane), warps 2..5 = epilogue
//          (TMEM lane quadrant = warp_idx % 4; one thread owns one output row)
// Epilogues (fused; the reference runs them as separate ATen ops):
//   EPI_BIAS       : + bias                      (QKV projection, region embedding, vocabulary decoder)
//   EPI_BIAS_GELU  : erf-GELU(+ bias)            (BertIntermediate)
//   EPI_BIAS_RESID : + bias + fp32 residual      (BertSelfOutput / BertOutput dense, pre-LayerNorm)
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
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;  // power of two >= 32 for BN in {64,128,256}
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
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

template <int BN, int EPI, typename OutT, typename T16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + kStages * Cfg::kStageBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + kGemmBM - 1) / kGemmBM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + kGemmBK - 1) / kGemmBK;

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
        mbar_init(tempty_bar(a), 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * kGemmBM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
          tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kGemmBK, m0);
          tma_load_2d(sa + Cfg::kABytes, &tmap_b, full_bar(stage), kb * kGemmBK, n0);
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
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
          umma_commit(empty_bar(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps)
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const bool vec_ok = ((p.ldo * (long long)sizeof(OutT)) % 16 == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int m0 = (tile / n_tiles) * kGemmBM;
      const int n0 = (tile % n_tiles) * BN;
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      long long orow = m;
      if (p.rin > 0) orow = (long long)(m / p.rin) * p.rout + p.roff + (m % p.rin);
      OutT* out_row = reinterpret_cast<OutT*>(p.out) + orow * p.ldo;
      const float* res_row = (EPI == EPI_BIAS_RESID) ? p.resid + (long long)m * p.ldr : nullptr;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int nc = n0 + c * 32;
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
        if (row_ok) {
          if (EPI == EPI_BIAS_RESID) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 r4 = *reinterpret_cast<const float4*>(res_row + nc + j);
                v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nc + j < p.N) v[j] += res_row[nc + j];
            }
          }
          if (full && vec_ok) {
            if (sizeof(OutT) == 4) {
              float4* o = reinterpret_cast<float4*>(out_row + nc);
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              uint4* o = reinterpret_cast<uint4*>(out_row + nc);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = Cvt<T16>::pack2(v[8 * j + 0], v[8 * j + 1]);
                u.y = Cvt<T16>::pack2(v[8 * j + 2], v[8 * j + 3]);
                u.z = Cvt<T16>::pack2(v[8 * j + 4], v[8 * j + 5]);
                u.w = Cvt<T16>::pack2(v[8 * j + 6], v[8 * j + 7]);
                o[j] = u;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (nc + j < p.N) {
                if (sizeof(OutT) == 4) reinterpret_cast<float*>(out_row)[nc + j] = v[j];
                else reinterpret_cast<T16*>(out_row)[nc + j] = Cvt<T16>::from(v[j]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace cptk
