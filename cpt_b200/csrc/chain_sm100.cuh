// Dataflow "chain" kernel for sm_100a: SEVERAL dependent stages of one encoder layer in ONE persistent launch.
//
// The reference runs one CaptionBertLayer (Oscar/oscar/modeling/modeling_bert.py:139-147) as ~25 ATen launches; round 1
// of this library ran it as 7 (QKV GEMM | attention | attention-out GEMM | LayerNorm | FFN-up GEMM | FFN-down GEMM |
// LayerNorm), every kernel boundary costing a drain (the slowest CTA's last epilogue), a grid-wide dependency wait and a
// pipeline refill: ~35 % of each short-K GEMM launch.  Here the stages that follow the attention kernel,
//     attention.output.dense (+bias +residual)   -> BertSelfOutput.LayerNorm
//  -> intermediate.dense (+bias, erf-GELU)       -> output.dense (+bias +residual) -> BertOutput.LayerNorm
//  -> the NEXT layer's fused query/key/value projection,
// are ONE launch: a stage is a list of TASKS (a 256 x 256 output tile computed by a CTA pair with
// tcgen05.mma.cta_group::2, or 64 rows of LayerNorm), every task waits for the rows it reads through counters in global
// memory (L2) and publishes the rows it wrote the same way, and the host gives every CTA pair an ordered task list
// (list scheduling in stage-major order, so a task only ever waits for tasks that precede it in every list: no
// deadlock with all pairs resident).  There is no grid-wide barrier anywhere: while the last tiles of one stage are
// still in their epilogues, pairs that are done already run the next stage's tiles on rows that are complete.
//
//   roles per CTA (as gemm_sm100.cuh): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (leader CTA), warps 2..9
//   = epilogue (TMEM -> registers -> bias / GELU -> swizzled staging -> TMA bulk store, or cp.reduce.async.bulk .add
//   into the fp32 residual stream) AND the LayerNorm tasks (a warp owns a row; two-pass statistics in registers).
//   The MMA issuer runs up to two accumulators ahead of the epilogue warps, so LayerNorm tasks hide behind main loops.
//
//   cross-CTA ordering:  writer: bulk stores complete (cp.async.bulk.wait_group 0) / generic stores fenced
//                                -> fence.proxy.async -> red.release.gpu on the counter
//                        reader: ld.acquire.gpu spin on the counter -> fence.proxy.async -> TMA loads / ld.global.cg
#pragma once
#include "gemm_sm100.cuh"
#include "rowwise.cuh"

namespace cptk {

constexpr int kChainMaxStages = 8;
constexpr int kChainMaxMaps = 4;
constexpr int kChainBN = 256;
constexpr int kChainLnRows = 32;     // rows per CTA per LayerNorm task (4 per epilogue warp); a pair task = 64 rows
constexpr int kChainMaxTasks = 512;  // per CTA pair (the list is staged in shared memory)

enum ChainKind { CHAIN_GEMM = 0, CHAIN_LN = 1 };

struct ChainStage {
  int kind;
  // CHAIN_GEMM: out[M,N] (+)= epi(A[M,K] . W[N,K]^T + bias);  CHAIN_LN: out[M,N] = LayerNorm(ln_in[M,N]) gamma + beta
  int M, N, K;
  int gelu;      // erf-GELU after the bias (BertIntermediate)
  int out_fp32;  // 1: the fp32 tile is ADDED into the destination by the TMA store (residual already there); 0: 16-bit
  int ksplit;    // K cut into pieces that all add into the destination (out_fp32 only)
  int map;       // index of this stage's tensor maps in ChainMaps
  const float* bias;
  // CHAIN_GEMM with ln != 0 — dense + bias + residual + LayerNorm in the tile's epilogue (BertSelfOutput / BertOutput):
  //   x = A W^T + bias + resid, out32 / out16 = LayerNorm(x) gamma + beta.  A row's statistics span the stage's N tiles
  //   (other CTA pairs): every (tile, column half) publishes (mean, M2) of its 128 columns per row in `part`, bumps
  //   sflag[m tile], waits until all 2 * n_tiles partials of the M tile are there and merges them (Chan et al.).
  // ln == 2 — the LayerNorm is DEFERRED to the consumers (round 1's "LayerNorm folding", here with TMA stores and no
  //   atomics): x = A W^T + bias + LN_prev(resid) is written PRE-LayerNorm (out32 fp32 / out16 raw 16-bit) together with
  //   the (mean, M2) partials of its rows in `part`; nobody waits for anyone inside an epilogue.  LN_prev is the
  //   LayerNorm of the residual's own rows applied on the fly from `rpart` (NULL: the residual is used as it is) with
  //   gamma / beta / eps.  A consumer GEMM reads the raw 16-bit rows as its A operand with gamma folded into its weight
  //   (host: fold_weight_kernel) and finishes the normalisation in its epilogue: `apart` != NULL ->
  //   out = rstd_m (acc - mean_m gvec_n) + bias_n, (mean, rstd) merged from the producer's partials.
  int ln;
  int map2;              // tensor map of out16 (16-bit), -1: none
  int map_r;             // tensor map of resid (fp32, 32 x 32 boxes) for ln == 2
  const float2* rpart;   // ln == 2: partials of the residual's rows (NULL: residual already normalised)
  const float2* apart;   // consumer side: partials of the A operand's rows (NULL: A is final)
  const float* gvec;     // consumer side: g_n = sum_k fp16(gamma_k W_nk)
  const float* resid;    // fp32 [M, ldr]
  long long ldr;
  float2* part;          // [n_tiles * 2][rows padded to the M-tile grid] (mean, M2) of 128 columns
  unsigned* sflag;       // [m_tiles] += 1 per epilogue warp per tile once its partials are published
  const float* ln_in;
  const float* gamma;
  const float* beta;
  float eps;
  float* out32;  // LayerNorm outputs (either may be NULL)
  void* out16;
  // readiness counters, one per 128-row M tile
  const unsigned* dep;  // rows this stage READS are complete when dep[mt] >= target (NULL: produced by an earlier launch)
  unsigned dep_target;  // 0 = "the number of rows in the M tile" (the producer is a LayerNorm stage)
  unsigned* done;       // this stage adds here: GEMM +1 per epilogue warp per (tile, K piece); LayerNorm + rows
};

struct ChainMaps {
  CUtensorMap a[kChainMaxMaps], b[kChainMaxMaps], o[kChainMaxMaps], o2[kChainMaxMaps], r[kChainMaxMaps];
};

struct ChainParams {
  int n_stages;
  ChainStage st[kChainMaxStages];
  const int* tasks;  // [pairs][pitch]: (stage << 24) | index, -1 terminated
  int pitch;
  int no_wait;       // chain2: stage 0 reads rows a STILL RUNNING attention launch publishes: no griddepcontrol.wait
  // optional event log (CPT_B200_CHAIN_TRACE=1), leader CTAs only: trace[(pair * pitch + i) * 16 + k] in SM clocks since
  // the CTA's entry: 0 producer starts waiting for the rows | 1 rows ready | 2 last load issued | 3 accumulator free
  // | 4 last MMA committed | 5 epilogue / LayerNorm starts waiting | 6 ready | 7 outputs published | 8 task code
  // | 9 tile's stores issued | 10..13 fused LayerNorm epilogue: accumulator ready, pass 1 done, statistics of the row
  // complete, pass 2 done; trace_hdr[pair * 2] = globaltimer (ns) and clock64 at entry
  long long* trace;
  long long* trace_hdr;
};

struct ChainCfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;           // this CTA's 128 rows
  static constexpr int kBBytes = (kChainBN / 2) * kGemmBK * 2;    // this CTA's half of the 256 weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = 4;
  static constexpr int kStageEpi = 4096;                          // one 32 x 32 fp32 block per epilogue warp
  static constexpr int kStage16 = 2048;                           // ... and one 32 x 32 16-bit block (fused LayerNorm)
  static constexpr int kStageRes = 4096;                          // ... and the landing block of the residual (TMA)
  static constexpr int kBiasBytes = 3 * (kChainBN / 2) * 4;       // bias, gamma, beta slices of the warp's 128 columns
  static constexpr int kEpiBytes = kGemmEpiWarps * (kStageEpi + kStage16 + kStageRes + kBiasBytes);
  static constexpr int kTaskBytes = kChainMaxTasks * 4;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiBytes + 512 + kTaskBytes;
  static_assert(kSmemBytes <= kSmemLimit, "shared memory");
};

__host__ __device__ inline int chain_rows_padded(int M) { return 2 * ((((M + kGemmBM - 1) / kGemmBM) + 1) / 2) * kGemmBM; }
inline int chain_m_tiles2(int M) { return 2 * ((((M + kGemmBM - 1) / kGemmBM) + 1) / 2); }  // counters per stage
// per stage: the "rows published" counters and the "row statistics published" counters of a fused LayerNorm
inline size_t chain_counter_bytes(int M, int n_stages) {
  return (size_t)n_stages * 2 * chain_m_tiles2(M) * sizeof(unsigned);
}
// scratch of the fused LayerNorm epilogues: (mean, M2) per row per 128-column slice, for up to two such stages
inline size_t chain_part_bytes(int M, int N) {
  return (size_t)2 * (2 * ((N + kChainBN - 1) / kChainBN)) * chain_rows_padded(M) * sizeof(float2);
}

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void flag_add_release(unsigned* addr, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
// Bounded like mbar_wait: a scheduling bug becomes a trapped launch, not a hung GPU.
__device__ __forceinline__ void flag_wait_ge(const unsigned* addr, unsigned target) {
  unsigned v, spins = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    if (v >= target) break;
    if (++spins > (1u << 23)) {
      printf("cpt_b200: chain dependency wait timed out (block %d thread %d: have %u need %u)\n", blockIdx.x, threadIdx.x,
             v, target);
      __trap();
    }
    __nanosleep(32);
  }
}

// LayerNorm of up to 4 rows by one warp (BertLayerNorm: biased variance, eps inside the sqrt, two-pass statistics).
// Rows come straight from L2 (ld.global.cg: another SM wrote them during this launch).  ALL loads of the ROWS rows are
// issued before the first use: a CTA has only 8 such warps, so memory-level parallelism per warp is what sets the
// rate (the first version, two rows in flight, took 11 us per 32-row task; the event log showed it on the critical
// path of the whole chain).
template <typename T16, int NV, int ROWS>
__device__ __noinline__ void chain_ln_batch(const ChainStage& s, int row0, int nrows, int lane, long long* tmark) {
  constexpr int H = NV * 128;
  const long long t_begin = tmark ? clock64() : 0;
  float4 x[ROWS][NV];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (r < nrows) {
      const float4* src = reinterpret_cast<const float4*>(s.ln_in + (long long)(row0 + r) * H);
#pragma unroll
      for (int i = 0; i < NV; ++i) x[r][i] = __ldcg(src + i * 32 + lane);
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (r < nrows) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) sum += (x[r][i].x + x[r][i].y) + (x[r][i].z + x[r][i].w);
      const float mean = warp_sum(sum) / (float)H;
      if (tmark && r == 0) tmark[0] = clock64() - t_begin;   // row 0 arrived
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float a = x[r][i].x - mean, b = x[r][i].y - mean, c = x[r][i].z - mean, d = x[r][i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + s.eps);
      if (tmark && r == 0) tmark[3] = clock64() - t_begin;   // row 0 statistics done
      const long long orow = (long long)(row0 + r) * H;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(s.gamma + col));
        const float4 b = __ldg(reinterpret_cast<const float4*>(s.beta + col));
        float4 y;
        y.x = (x[r][i].x - mean) * rstd * g.x + b.x;
        y.y = (x[r][i].y - mean) * rstd * g.y + b.y;
        y.z = (x[r][i].z - mean) * rstd * g.z + b.z;
        y.w = (x[r][i].w - mean) * rstd * g.w + b.w;
        if (s.out32) *reinterpret_cast<float4*>(s.out32 + orow + col) = y;
        if (s.out16) {
          uint2 u;
          u.x = Cvt<T16>::pack2(y.x, y.y);
          u.y = Cvt<T16>::pack2(y.z, y.w);
          *reinterpret_cast<uint2*>(reinterpret_cast<T16*>(s.out16) + orow + col) = u;
        }
      }
      if (tmark && r == 0) tmark[4] = clock64() - t_begin;   // row 0 stored
    }
  }
}
template <typename T16>
__device__ __forceinline__ void chain_ln_rows(const ChainStage& s, int row0, int nrows, int lane, long long* tmark) {
  switch (s.N >> 7) {  // float4 per lane; the host admits these widths only
    case 1: chain_ln_batch<T16, 1, 4>(s, row0, nrows, lane, tmark); break;
    case 2: chain_ln_batch<T16, 2, 4>(s, row0, nrows, lane, tmark); break;
    case 4: chain_ln_batch<T16, 4, 4>(s, row0, nrows, lane, tmark); break;
    case 6: chain_ln_batch<T16, 6, 4>(s, row0, nrows, lane, tmark); break;
    default:  // 8 (H = 1024): two rows at a time keep the register count of the 4 x 6 case
      chain_ln_batch<T16, 8, 2>(s, row0, nrows < 2 ? nrows : 2, lane, tmark);
      if (nrows > 2) chain_ln_batch<T16, 8, 2>(s, row0 + 2, nrows - 2, lane, nullptr);
      break;
  }
}

// (mean, rstd) of one row from the (mean, M2) partials of its 128-column slices (Chan et al. pairwise merge)
__device__ __forceinline__ float2 chain_row_stats(const float2* part, int width, int m_pad, int row, float eps) {
  float cnt = 0.f, mean = 0.f, m2 = 0.f;
  for (int t = 0; t * (kChainBN / 2) < width; ++t) {
    const float nb = (float)min(kChainBN / 2, width - t * (kChainBN / 2));
    const float2 pm = __ldcg(part + (long long)t * m_pad + row);
    const float tot = cnt + nb, delta = pm.x - mean;
    mean += delta * (nb / tot);
    m2 += pm.y + delta * delta * (cnt * nb / tot);
    cnt = tot;
  }
  return make_float2(mean, 1.0f / sqrtf(m2 / (float)width + eps));
}

// per-warp shared-memory areas of the epilogue
struct ChainEpiSmem {
  uint8_t* pad;        // 4 KB: a 32 x 32 fp32 block (128B-swizzled TMA box / transpose area)
  uint32_t pad_u32;
  uint8_t* pad16;      // 2 KB: a 32 x 32 16-bit block (64B-swizzled TMA box)
  uint32_t pad16_u32;
  float *sv0, *sv1, *sv2;  // bias | gamma | beta of the warp's 128 columns
  uint8_t* padr;       // 4 KB: the residual's 32 x 32 fp32 block, landed by TMA (ln == 2)
  uint32_t padr_u32;
  uint32_t rbar;       // this warp's mbarrier for the residual loads
};

// Epilogue of a dense + bias + LN_prev(residual) tile with the LayerNorm of the RESULT deferred (ChainStage::ln == 2).
// One pass: the residual block of chunk c + 1 is in flight (TMA) while chunk c is combined, its statistics folded into
// the running (count, mean, M2) of this warp's 128 columns, and written out pre-LayerNorm as fp32 and raw 16-bit.
template <typename T16>
__device__ __forceinline__ void chain_epilogue_defer(const ChainStage& s, const CUtensorMap* mo32, const CUtensorMap* mo16,
                                                     const CUtensorMap* mr, int mrow0, int ncol0, int n0, uint32_t t_row,
                                                     uint32_t tfull, uint32_t tfull_phase, const ChainEpiSmem& e, int lane,
                                                     bool& staging_busy, uint32_t& rphase) {
  constexpr int BN = kChainBN;
  const int m_pad = chain_rows_padded(s.M);
  const int n_live = max(0, min(BN / 2, s.N - ncol0));
  const bool rows_ok = mrow0 < s.M;
  const int row = mrow0 + lane;
  const bool early = s.dep == nullptr;  // residual (and its statistics) come from an earlier launch
  auto load_resid = [&](int c) {
    if (lane == 0) {
      mbar_expect_tx(e.rbar, 4096);
      tma_load_2d(e.padr_u32, mr, e.rbar, ncol0 + c * 32, mrow0);
    }
  };
  const bool work = rows_ok && n_live > 0;
  float2 rs = make_float2(0.f, 1.f);
  if (work && early) {
    load_resid(0);
    if (s.rpart) rs = chain_row_stats(s.rpart, s.N, m_pad, row, s.eps);
  }
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  if (work && !early) {
    if (lane == 0) fence_proxy_async_global();
    load_resid(0);
    if (s.rpart) rs = chain_row_stats(s.rpart, s.N, m_pad, row, s.eps);
  }
  float cnt = 0.f, mean = 0.f, m2 = 0.f;
  if (work) {
    const float ra = s.rpart ? rs.y : 1.f, rb = s.rpart ? -rs.x * rs.y : 0.f;  // LN_prev(r) = (r * ra + rb) * gamma + beta
    uint32_t rbuf[32];
    tmem_ld_32x32b_x32(t_row, rbuf);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int nc = ncol0 + c * 32;
      const bool live = nc < s.N;  // warp-uniform; N is a multiple of 128: a live chunk is whole
      if (live) {
        mbar_wait(e.rbar, rphase);
        rphase ^= 1u;
      }
      tmem_ld_wait();
      float x[32];
      if (live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 r4 = *reinterpret_cast<const float4*>(e.padr + lane * 128 + ((j ^ (lane & 7)) << 4));
          const float4 b4 = *reinterpret_cast<const float4*>(e.sv0 + c * 32 + 4 * j);
          if (s.rpart) {
            const float4 g4 = *reinterpret_cast<const float4*>(e.sv1 + c * 32 + 4 * j);
            const float4 e4 = *reinterpret_cast<const float4*>(e.sv2 + c * 32 + 4 * j);
            r4.x = fmaf(fmaf(r4.x, ra, rb), g4.x, e4.x);
            r4.y = fmaf(fmaf(r4.y, ra, rb), g4.y, e4.y);
            r4.z = fmaf(fmaf(r4.z, ra, rb), g4.z, e4.z);
            r4.w = fmaf(fmaf(r4.w, ra, rb), g4.w, e4.w);
          }
          x[4 * j] = __uint_as_float(rbuf[4 * j]) + b4.x + r4.x;
          x[4 * j + 1] = __uint_as_float(rbuf[4 * j + 1]) + b4.y + r4.y;
          x[4 * j + 2] = __uint_as_float(rbuf[4 * j + 2]) + b4.z + r4.z;
          x[4 * j + 3] = __uint_as_float(rbuf[4 * j + 3]) + b4.w + r4.w;
        }
      }
      __syncwarp();  // every lane has read the residual block: the next one may land
      if (c + 1 < 4 && nc + 32 < s.N) load_resid(c + 1);
      if (c + 1 < 4) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, rbuf);
      if (live) {
        // statistics of these 32 values (two passes in registers), merged into the running ones
        float sm = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) sm += x[j];
        const float mc = sm * (1.0f / 32.0f);
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float d0 = x[j] - mc, d1 = x[j + 1] - mc;
          q0 = fmaf(d0, d0, q0);
          q1 = fmaf(d1, d1, q1);
        }
        const float tot = cnt + 32.f, delta = mc - mean;
        mean += delta * (32.f / tot);
        m2 += (q0 + q1) + delta * delta * (cnt * 32.f / tot);
        cnt = tot;
        if (staging_busy) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          staging_busy = false;
        }
        if (mo32 != nullptr) {
          uint8_t* brow = e.pad + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(brow + ((j ^ (lane & 7)) * 16)) =
                make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
        }
        if (mo16 != nullptr) {
          uint8_t* brow = e.pad16 + lane * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = Cvt<T16>::pack2(x[8 * j + 0], x[8 * j + 1]);
            u.y = Cvt<T16>::pack2(x[8 * j + 2], x[8 * j + 3]);
            u.z = Cvt<T16>::pack2(x[8 * j + 4], x[8 * j + 5]);
            u.w = Cvt<T16>::pack2(x[8 * j + 6], x[8 * j + 7]);
            *reinterpret_cast<uint4*>(brow + ((j ^ ((lane >> 1) & 3)) * 16)) = u;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (mo32 != nullptr) tma_store_2d(mo32, e.pad_u32, nc, mrow0);
          if (mo16 != nullptr) tma_store_2d(mo16, e.pad16_u32, nc, mrow0);
          tma_store_commit();
        }
        staging_busy = true;
      }
      __syncwarp();
    }
    tmem_ld_wait();
  }
  if (rows_ok)  // this slice's partial: (mean, M2) of its live columns (an empty slice is never read)
    s.part[(long long)(((n0 / BN) * 2 + (ncol0 - n0) / (BN / 2))) * m_pad + row] = make_float2(mean, m2);
  __threadfence();  // ordered before this warp's "rows published" increment (the caller issues it)
  __syncwarp();
}


// Epilogue of a dense + bias + residual + LayerNorm tile (see ChainStage::ln).  The warp owns rows mrow0..+31 (thread =
// row = TMEM lane) and 128 columns from ncol0; x lives in TMEM (written back over the accumulator) between the passes.
template <typename T16>
__device__ __forceinline__ void chain_epilogue_ln(const ChainStage& s, const CUtensorMap* mo32, const CUtensorMap* mo16,
                                                  int mrow0, int ncol0, int n0, int mt, uint32_t t_row,
                                                  uint32_t tfull, uint32_t tfull_phase, const ChainEpiSmem& e, int lane,
                                                  bool& staging_busy, long long* tmark, long long c_entry) {
  constexpr int BN = kChainBN;
  const int n_tiles = (s.N + BN - 1) / BN;
  const int m_pad = chain_rows_padded(s.M);
  const int n_live = max(0, min(BN / 2, s.N - ncol0));   // live columns of this warp's half (warp-uniform)
  const bool rows_ok = mrow0 < s.M;
  const int row = mrow0 + lane;
  float4 rr[8];
  auto load_resid = [&](int c) {  // 4 rows x 128 B per warp instruction: coalesced; transposed through the pad below
    const int col = ncol0 + c * 32 + (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = mrow0 + i * 4 + (lane >> 3);
      rr[i] = (r < s.M && col < s.N) ? __ldcg(reinterpret_cast<const float4*>(s.resid + (long long)r * s.ldr + col))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // the first residual chunk is in flight while the tile's MMAs finish — unless the residual itself is produced by an
  // earlier stage of THIS launch: then only the tile's completion (its producer waited for those rows) orders the read
  const bool early = s.dep == nullptr;
  if (rows_ok && early) load_resid(0);
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
  if (tmark) tmark[10] = clock64() - c_entry;
  if (rows_ok && !early) load_resid(0);
  float mean_l = 0.f, m2_l = 0.f;
  uint32_t rbuf[32];
  if (rows_ok) {
    // ---- pass 1a: x = acc + bias + residual -> TMEM, row sums
    float sum = 0.f;
    tmem_ld_32x32b_x32(t_row, rbuf);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = i * 4 + (lane >> 3), piece = lane & 7;
        *reinterpret_cast<float4*>(e.pad + rl * 128 + ((piece ^ (rl & 7)) << 4)) = rr[i];
      }
      __syncwarp();
      if (c + 1 < 4) load_resid(c + 1);
      tmem_ld_wait();
      uint32_t x[32];
      const int nl = max(0, min(32, n_live - c * 32));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 r4 = *reinterpret_cast<const float4*>(e.pad + lane * 128 + ((j ^ (lane & 7)) << 4));
        const float4 b4 = *reinterpret_cast<const float4*>(e.sv0 + c * 32 + 4 * j);
        const float x0 = __uint_as_float(rbuf[4 * j]) + b4.x + r4.x, x1 = __uint_as_float(rbuf[4 * j + 1]) + b4.y + r4.y;
        const float x2 = __uint_as_float(rbuf[4 * j + 2]) + b4.z + r4.z, x3 = __uint_as_float(rbuf[4 * j + 3]) + b4.w + r4.w;
        x[4 * j] = __float_as_uint(x0);
        x[4 * j + 1] = __float_as_uint(x1);
        x[4 * j + 2] = __float_as_uint(x2);
        x[4 * j + 3] = __float_as_uint(x3);
        if (4 * j < nl) sum += (x0 + x1) + (x2 + x3);  // N is a multiple of 128: whole groups of 4
      }
      if (c + 1 < 4) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, rbuf);
      tmem_st_32x32b_x32(t_row + c * 32, x);
      __syncwarp();  // every lane has read its pad row before the next chunk's residual lands there
    }
    tmem_ld_wait();
    tmem_st_wait();
    // ---- pass 1b: centred second moment of the same 128 columns
    mean_l = n_live > 0 ? sum / (float)n_live : 0.f;
    tmem_ld_32x32b_x32(t_row, rbuf);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      tmem_ld_wait();
      const int nl = max(0, min(32, n_live - c * 32));
      float q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        if (j < nl) {
          const float d0 = __uint_as_float(rbuf[j]) - mean_l, d1 = __uint_as_float(rbuf[j + 1]) - mean_l;
          q0 = fmaf(d0, d0, q0);
          q1 = fmaf(d1, d1, q1);
        }
      }
      if (c + 1 < 4) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, rbuf);
      m2_l += q0 + q1;
    }
    tmem_ld_wait();
    s.part[(long long)(((n0 / BN) * 2 + (ncol0 - n0) / (BN / 2))) * m_pad + row] = make_float2(mean_l, m2_l);
  }
  if (tmark) tmark[11] = clock64() - c_entry;
  __threadfence();
  __syncwarp();
  if (lane == 0) flag_add_release(s.sflag + mt, 1u);
  if (!rows_ok) return;
  if (lane == 0) flag_wait_ge(s.sflag + mt, (unsigned)(n_tiles * kGemmEpiWarps));
  __syncwarp();
  if (tmark) tmark[12] = clock64() - c_entry;
  // ---- merge the 2 * n_tiles partials of this row: n, mean, M2 (Chan et al. pairwise update)
  float cnt = 0.f, mean = 0.f, m2 = 0.f;
  for (int t = 0; t < 2 * n_tiles; ++t) {
    const float nb = (float)max(0, min(BN / 2, s.N - t * (BN / 2)));
    if (nb > 0.f) {
      const float2 pm = __ldcg(s.part + (long long)t * m_pad + row);
      const float tot = cnt + nb, delta = pm.x - mean;
      mean += delta * (nb / tot);
      m2 += pm.y + delta * delta * (cnt * nb / tot);
      cnt = tot;
    }
  }
  const float rstd = 1.0f / sqrtf(m2 / (float)s.N + s.eps);
  // ---- pass 2: normalise, scale, shift -> fp32 stream and 16-bit shadow through TMA bulk stores
  tmem_ld_32x32b_x32(t_row, rbuf);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int nc = ncol0 + c * 32;
    const bool live = nc < s.N;
    tmem_ld_wait();
    float y[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 g4 = *reinterpret_cast<const float4*>(e.sv1 + c * 32 + j);
      const float4 b4 = *reinterpret_cast<const float4*>(e.sv2 + c * 32 + j);
      y[j] = (__uint_as_float(rbuf[j]) - mean) * rstd * g4.x + b4.x;
      y[j + 1] = (__uint_as_float(rbuf[j + 1]) - mean) * rstd * g4.y + b4.y;
      y[j + 2] = (__uint_as_float(rbuf[j + 2]) - mean) * rstd * g4.z + b4.z;
      y[j + 3] = (__uint_as_float(rbuf[j + 3]) - mean) * rstd * g4.w + b4.w;
    }
    if (staging_busy) {
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
      staging_busy = false;
    }
    if (c + 1 < 4) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, rbuf);
    if (live) {
      if (mo32 != nullptr) {
        uint8_t* brow = e.pad + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(brow + ((j ^ (lane & 7)) * 16)) =
              make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
      }
      if (mo16 != nullptr) {
        uint8_t* brow = e.pad16 + lane * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = Cvt<T16>::pack2(y[8 * j + 0], y[8 * j + 1]);
          u.y = Cvt<T16>::pack2(y[8 * j + 2], y[8 * j + 3]);
          u.z = Cvt<T16>::pack2(y[8 * j + 4], y[8 * j + 5]);
          u.w = Cvt<T16>::pack2(y[8 * j + 6], y[8 * j + 7]);
          *reinterpret_cast<uint4*>(brow + ((j ^ ((lane >> 1) & 3)) * 16)) = u;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (mo32 != nullptr) tma_store_2d(mo32, e.pad_u32, nc, mrow0);
        if (mo16 != nullptr) tma_store_2d(mo16, e.pad16_u32, nc, mrow0);
        tma_store_commit();
      }
      staging_busy = true;
    }
    __syncwarp();
  }
  tmem_ld_wait();
  if (tmark) tmark[13] = clock64() - c_entry;
}

template <typename T16>
__global__ void __launch_bounds__(kGemmThreads, 1)
chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  using Cfg = ChainCfg;
  constexpr int kStages = Cfg::kStages;
  constexpr int BN = kChainBN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t epi_base = smem_base + kStages * Cfg::kStageBytes;
  const uint32_t bars = epi_base + Cfg::kEpiBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  uint8_t* epi_gen = smem_gen + kStages * Cfg::kStageBytes;
  const int* task_s = reinterpret_cast<const int*>(epi_gen + Cfg::kEpiBytes + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int pair_id = blockIdx.x >> 1;

  const bool tracing = p.trace != nullptr && leader;
  long long c_entry = 0;
  if (tracing) {
    c_entry = clock64();
    if (threadIdx.x == 0) {
      long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      p.trace_hdr[pair_id * 2] = gt;
      p.trace_hdr[pair_id * 2 + 1] = c_entry;
    }
  }
  auto mark = [&](int i, int k) {
    if (tracing) p.trace[((long long)pair_id * p.pitch + i) * 16 + k] = clock64() - c_entry;
  };
  pdl_launch_dependents();
  // the task list is static data (uploaded when the schedule was built): staged before the dependency wait
  {
    int* dst = reinterpret_cast<int*>(epi_gen + Cfg::kEpiBytes + 512);
    const int* src = p.tasks + (long long)pair_id * p.pitch;
    for (int i = threadIdx.x; i < p.pitch; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_stages; ++s)
      if (p.st[s].kind == CHAIN_GEMM) {
        tma_prefetch_desc(&maps.a[p.st[s].map]);
        tma_prefetch_desc(&maps.b[p.st[s].map]);
        tma_prefetch_desc(&maps.o[p.st[s].map]);
        if (p.st[s].ln && p.st[s].map2 >= 0) tma_prefetch_desc(&maps.o2[p.st[s].map2]);
        if (p.st[s].ln == 2) tma_prefetch_desc(&maps.r[p.st[s].map_r]);
      }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 2 * kGemmEpiWarps);
      }
      for (int w = 0; w < kGemmEpiWarps; ++w) mbar_init(bars + 8u * (2 * kStages + 6 + w), 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  pdl_wait();  // outputs of the previous launch (attention context, residual stream) are visible from here on
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // decode of a GEMM task: cluster tile (256 rows x 256 columns) and K piece
  struct TileAt {
    int m0, n0, kb_begin, kb_end, ks;
  };
  auto decode = [&](const ChainStage& s, int idx) {
    const int m_tiles = (s.M + kGemmBM - 1) / kGemmBM, n_tiles = (s.N + BN - 1) / BN;
    const int mn = ((m_tiles + 1) >> 1) * n_tiles;
    const int ksplit = s.ksplit > 1 ? s.ksplit : 1;
    const int num_kb = (s.K + kGemmBK - 1) / kGemmBK;
    const int tile = idx % mn, ks = idx / mn;
    TileAt t;
    t.m0 = ((tile / n_tiles) * 2 + (int)crank) * kGemmBM;
    t.n0 = (tile % n_tiles) * BN;
    t.kb_begin = (int)((long long)ks * num_kb / ksplit);
    t.kb_end = (int)((long long)(ks + 1) * num_kb / ksplit);
    t.ks = ks;
    return t;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0;; ++i) {
        const int task = task_s[i];
        if (task < 0) break;
        const ChainStage& s = p.st[task >> 24];
        if (s.kind != CHAIN_GEMM) continue;
        const TileAt t = decode(s, task & 0xFFFFFF);
        const CUtensorMap* ma = &maps.a[s.map];
        const CUtensorMap* mb = &maps.b[s.map];
        mark(i, 0);
        if (s.dep != nullptr && t.m0 < s.M) {  // this CTA's 128 rows of the A operand
          const unsigned target = s.dep_target ? s.dep_target : (unsigned)min(kGemmBM, s.M - t.m0);
          flag_wait_ge(s.dep + t.m0 / kGemmBM, target);
          fence_proxy_async_global();
        }
        mark(i, 1);
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          // both CTAs' bytes are counted on the leader's barrier (the leader alone issues the MMAs)
          if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
          const uint32_t lbar = mapa_cluster(full_bar(stage), 0);
          tma_load_2d_2cta(sa, ma, lbar, kb * kGemmBK, t.m0);
          tma_load_2d_2cta(sb, mb, lbar, kb * kGemmBK, t.n0 + (int)crank * (BN / 2));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        mark(i, 2);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && leader) {
      const uint32_t idesc = make_idesc_f16(2 * kGemmBM, BN, Cvt<T16>::kFmt, 0, 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int i = 0;; ++i) {
        const int task = task_s[i];
        if (task < 0) break;
        const ChainStage& s = p.st[task >> 24];
        if (s.kind != CHAIN_GEMM) continue;
        const TileAt t = decode(s, task & 0xFFFFFF);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        ++it;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        mark(i, 3);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = t.kb_begin; kb < t.kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sa + Cfg::kABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k)
            umma_f16_2cta(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, ((kb - t.kb_begin) | k) != 0);
          umma_commit_2cta_mc(empty_bar(stage), 3);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_2cta_mc(tfull_bar(acc), 3);
        mark(i, 4);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue + LayerNorm (8 warps)
    const int ew = warp - 2;
    const int q = warp & 3;    // TMEM lane quadrant this warp may read
    const int half = ew >> 2;  // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    uint8_t* pad = epi_gen + ew * Cfg::kStageEpi;
    const uint32_t pad_u32 = epi_base + ew * Cfg::kStageEpi;
    float* sv0 = reinterpret_cast<float*>(epi_gen + kGemmEpiWarps * (Cfg::kStageEpi + Cfg::kStage16 + Cfg::kStageRes) +
                                          ew * Cfg::kBiasBytes);
    ChainEpiSmem es;
    es.pad = pad;
    es.pad_u32 = pad_u32;
    es.pad16 = epi_gen + kGemmEpiWarps * Cfg::kStageEpi + ew * Cfg::kStage16;
    es.pad16_u32 = epi_base + kGemmEpiWarps * Cfg::kStageEpi + ew * Cfg::kStage16;
    es.sv0 = sv0;
    es.sv1 = sv0 + kColsPerWarp;
    es.sv2 = sv0 + 2 * kColsPerWarp;
    es.padr = epi_gen + kGemmEpiWarps * (Cfg::kStageEpi + Cfg::kStage16) + ew * Cfg::kStageRes;
    es.padr_u32 = epi_base + kGemmEpiWarps * (Cfg::kStageEpi + Cfg::kStage16) + ew * Cfg::kStageRes;
    es.rbar = bars + 8u * (2 * kStages + 6 + ew);
    uint32_t rphase = 0;
    bool staging_busy = false;  // a bulk store may still be reading this warp's staging block
    int it = 0;
    for (int i = 0;; ++i) {
      const int task = task_s[i];
      if (task < 0) break;
      const ChainStage& s = p.st[task >> 24];
      const bool tr = tracing && ew == 0 && lane == 0;
      if (tr) {
        mark(i, 5);
        p.trace[((long long)pair_id * p.pitch + i) * 16 + 8] = task;
      }
      if (s.kind == CHAIN_LN) {
        // ---- 64 rows per pair task: this CTA's 32, 4 per warp
        const int row0 = ((task & 0xFFFFFF) * 2 + (int)crank) * kChainLnRows + ew * 4;
        if (row0 < s.M) {
          const int nrows = min(4, s.M - row0);
          const int mt = row0 / kGemmBM;
          if (s.dep != nullptr) {
            if (lane == 0) {
              const unsigned target = s.dep_target ? s.dep_target : (unsigned)min(kGemmBM, s.M - mt * kGemmBM);
              flag_wait_ge(s.dep + mt, target);
            }
            __syncwarp();
          }
          if (tr) mark(i, 6);
          chain_ln_rows<T16>(s, row0, nrows, lane,
                             tr ? p.trace + ((long long)pair_id * p.pitch + i) * 16 : nullptr);
          if (tr) mark(i, 1);
          __threadfence();
          if (tr) mark(i, 2);
          __syncwarp();
          if (lane == 0 && s.done != nullptr) {
            fence_proxy_async_global();
            flag_add_release(s.done + mt, (unsigned)nrows);
          }
          if (tr) mark(i, 7);
        }
        continue;
      }
      // ---- GEMM tile epilogue
      const TileAt t = decode(s, task & 0xFFFFFF);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      ++it;
      const CUtensorMap* mo = &maps.o[s.map];
      const int mrow0 = t.m0 + q * 32;
      const int ncol0 = t.n0 + half * kColsPerWarp;
      const bool f32out = s.out_fp32 != 0;
      {  // this warp's bias slice -> smem while the tile's MMAs are still running (the first K piece carries the bias)
        const float* src = (t.ks == 0) ? s.bias : nullptr;
        const int nb = ncol0 + lane * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src != nullptr) {
          if (nb + 3 < s.N) {
            b4 = __ldg(reinterpret_cast<const float4*>(src + nb));
          } else {
            if (nb < s.N) b4.x = __ldg(src + nb);
            if (nb + 1 < s.N) b4.y = __ldg(src + nb + 1);
            if (nb + 2 < s.N) b4.z = __ldg(src + nb + 2);
          }
        }
        *reinterpret_cast<float4*>(sv0 + lane * 4) = b4;
        if (s.apart) {  // consumer of a deferred LayerNorm: g_n of the same columns
          const bool in = nb + 3 < s.N;
          *reinterpret_cast<float4*>(es.sv1 + lane * 4) =
              in ? __ldg(reinterpret_cast<const float4*>(s.gvec + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (s.ln && s.gamma != nullptr) {  // LayerNorm scale / shift of the same columns (N is a multiple of 128 here)
          const bool in = nb + 3 < s.N;
          *reinterpret_cast<float4*>(es.sv1 + lane * 4) =
              in ? __ldg(reinterpret_cast<const float4*>(s.gamma + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(es.sv2 + lane * 4) =
              in ? __ldg(reinterpret_cast<const float4*>(s.beta + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
      }
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN + half * kColsPerWarp;
      if (s.ln == 2) {
        chain_epilogue_defer<T16>(s, s.out32 ? mo : nullptr, s.map2 >= 0 ? &maps.o2[s.map2] : nullptr, &maps.r[s.map_r],
                                  mrow0, ncol0, t.n0, t_row, tfull_bar(acc), acc_phase, es, lane, staging_busy, rphase);
        if (tr) mark(i, 6);
      } else if (s.ln) {
        chain_epilogue_ln<T16>(s, s.out32 ? mo : nullptr, s.map2 >= 0 ? &maps.o2[s.map2] : nullptr, mrow0, ncol0, t.n0,
                               t.m0 / kGemmBM, t_row, tfull_bar(acc), acc_phase, es, lane, staging_busy,
                               tr ? p.trace + ((long long)pair_id * p.pitch + i) * 16 : nullptr, c_entry);
        if (tr) mark(i, 6);
      } else {
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (tr) mark(i, 6);
      if (mrow0 < s.M) {
        constexpr int NC = kColsPerWarp / 32;
        uint32_t rbuf[32];
        tmem_ld_32x32b_x32(t_row, rbuf);
        // A operand = raw pre-LayerNorm rows, gamma folded into W: finish the normalisation with this row's statistics
        float fa = 1.f, fb = 0.f;  // out = fa * acc + fb * g_n + c_n
        if (s.apart) {
          const float2 st = chain_row_stats(s.apart, s.K, chain_rows_padded(s.M), min(mrow0 + lane, s.M - 1), s.eps);
          fa = st.y;
          fb = -st.x * st.y;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int nc = ncol0 + c * 32;
          const bool live = nc < s.N;  // warp-uniform
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sv0 + c * 32 + j);
            if (s.apart) {
              const float4 g4 = *reinterpret_cast<const float4*>(es.sv1 + c * 32 + j);
              v[j] = fmaf(fa, __uint_as_float(rbuf[j]), fmaf(fb, g4.x, b4.x));
              v[j + 1] = fmaf(fa, __uint_as_float(rbuf[j + 1]), fmaf(fb, g4.y, b4.y));
              v[j + 2] = fmaf(fa, __uint_as_float(rbuf[j + 2]), fmaf(fb, g4.z, b4.z));
              v[j + 3] = fmaf(fa, __uint_as_float(rbuf[j + 3]), fmaf(fb, g4.w, b4.w));
            } else {
              v[j] = __uint_as_float(rbuf[j]) + b4.x;
              v[j + 1] = __uint_as_float(rbuf[j + 1]) + b4.y;
              v[j + 2] = __uint_as_float(rbuf[j + 2]) + b4.z;
              v[j + 3] = __uint_as_float(rbuf[j + 3]) + b4.w;
            }
          }
          if (s.gelu) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) gelu_erf4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (staging_busy) {  // the store that last read the staging block must have drained its smem reads
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            staging_busy = false;
          }
          if (c + 1 < NC) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, rbuf);
          if (live) {
            if (f32out) {
              uint8_t* brow = pad + lane * 128;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(brow + ((j ^ (lane & 7)) * 16)) =
                    make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              uint8_t* brow = pad + lane * 64;
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
              if (f32out) tma_reduce_add_2d(mo, pad_u32, nc, mrow0);
              else tma_store_2d(mo, pad_u32, nc, mrow0);
              tma_store_commit();
            }
            staging_busy = true;
          }
          __syncwarp();
        }
        tmem_ld_wait();
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_cluster(mapa_cluster(tempty_bar(acc), 0));
        if (tr) mark(i, 9);
        // publish: this warp's blocks of the tile are in L2 (a stage nobody reads within this launch only needs its
        // staging block back: the kernel's end waits for the stores themselves)
        if (s.done != nullptr) {
          tma_store_wait<0>();
          fence_proxy_async_global();
          flag_add_release(s.done + t.m0 / kGemmBM, 1u);
        } else {
          tma_store_wait_read<0>();
        }
        if (tr) mark(i, 7);
      }
      staging_busy = false;
      __syncwarp();
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA exits (or frees TMEM) while its peer may still signal it / read its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace cptk
