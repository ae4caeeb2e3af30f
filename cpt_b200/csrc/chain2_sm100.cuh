// Production form of the dataflow chain (see chain_sm100.cuh for the scheme): the LayerNorms of a layer are DEFERRED
// to their consumers, so the launch is four dependent GEMM stages and nothing else —
//     x1 = ctx Wo^T + b + LN2'(x2 of the previous layer)        "F" stage: fp32 pre-LayerNorm rows + raw 16-bit rows
//                                                                 + (mean, M2) partials of every row
//     inter = gelu(LN1(x1) W1^T + b1)                           "C" stage: A = raw rows, W carries gamma, epilogue
//                                                                 finishes the normalisation from the partials
//     x2 = inter W2^T + b2 + LN1(x1)                            "F" stage, residual normalised on the fly
//     qkv' = LN2(x2) Wqkv'^T + b'                               "C" stage of the NEXT layer
// What the event log of the first version (chain_sm100.cuh, ln == 1 / 2) showed: every synchronous global-memory access
// inside an epilogue costs 1.5 - 3 us while the TMA producers keep the memory system full, and an epilogue does four
// of them per tile.  So here nothing in an epilogue waits for global memory:
//   * the fp32 RESIDUAL tile of an F stage travels through the operand ring: after the tile's k-blocks the producer
//     loads it as four more ring slots (128 rows x 64 fp32 columns each), in flight while the main loop finishes; the
//     MMA issuer skips those slots, the epilogue warps read them in place and release them;
//   * row statistics of a consumer are fetched BEFORE the accumulator wait whenever the epilogue warp can see that the
//     producing stage has already published the rows (it polls the same counter the TMA producer waited on);
//   * stores are TMA bulk stores out of double-buffered staging blocks (cp.async.bulk.wait_group.read 1).
#pragma once
#include "chain_sm100.cuh"

namespace cptk {

struct Chain2Cfg {
  static constexpr int kStages = 4;
  static constexpr int kStageBytes = ChainCfg::kStageBytes;     // 32 KB: A 128 x 64 + B 128 x 64 (16-bit) = a residual slot
  static constexpr int kPad = 4096;                             // two of these per warp: staging blocks
  static constexpr int kPad16 = 2048;
  static constexpr int kVecBytes = 3 * (kChainBN / 2) * 4;
  static constexpr int kEpiBytes = kGemmEpiWarps * (2 * kPad + kPad16 + kVecBytes);
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiBytes + 512 + kChainMaxTasks * 4;
  static_assert(kSmemBytes <= kSmemLimit, "shared memory");
};

template <typename T16>
__global__ void __launch_bounds__(kGemmThreads, 1)
chain2_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  using Cfg = Chain2Cfg;
  constexpr int kStages = Cfg::kStages;
  constexpr int BN = kChainBN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t epi_base = smem_base + kStages * Cfg::kStageBytes;
  const uint32_t bars = epi_base + Cfg::kEpiBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto rfull_bar = [&](int s) { return bars + 8u * (2 * kStages + s); };  // residual slots: local to each CTA
  auto tfull_bar = [&](int a) { return bars + 8u * (3 * kStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (3 * kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (3 * kStages + 4);
  uint8_t* epi_gen = smem_gen + kStages * Cfg::kStageBytes;
  int* rel_cnt = reinterpret_cast<int*>(epi_gen + Cfg::kEpiBytes + 8 * (3 * kStages + 6));  // [kStages]
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
  {
    int* dst = reinterpret_cast<int*>(epi_gen + Cfg::kEpiBytes + 512);
    const int* src = p.tasks + (long long)pair_id * p.pitch;
    for (int i = threadIdx.x; i < p.pitch; i += blockDim.x) dst[i] = __ldg(src + i);
    if (threadIdx.x < kStages) rel_cnt[threadIdx.x] = 0;
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.n_stages; ++s) {
      tma_prefetch_desc(&maps.a[p.st[s].map]);
      tma_prefetch_desc(&maps.b[p.st[s].map]);
      if (p.st[s].ln != 2 || p.st[s].out32 != nullptr) tma_prefetch_desc(&maps.o[p.st[s].map]);
      if (p.st[s].map2 >= 0) tma_prefetch_desc(&maps.o2[p.st[s].map2]);
      if (p.st[s].ln == 2) tma_prefetch_desc(&maps.r[p.st[s].map_r]);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
        mbar_init(rfull_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 2 * kGemmEpiWarps);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  struct TileAt {
    int m0, n0, nkb;
  };
  auto decode = [&](const ChainStage& s, int idx) {
    const int m_tiles = (s.M + kGemmBM - 1) / kGemmBM, n_tiles = (s.N + BN - 1) / BN;
    TileAt t;
    t.m0 = ((idx / n_tiles) * 2 + (int)crank) * kGemmBM;
    t.n0 = (idx % n_tiles) * BN;
    t.nkb = (s.K + kGemmBK - 1) / kGemmBK;
    (void)m_tiles;
    return t;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      };
      for (int i = 0;; ++i) {
        const int task = task_s[i];
        if (task < 0) break;
        const ChainStage& s = p.st[task >> 24];
        const TileAt t = decode(s, task & 0xFFFFFF);
        const CUtensorMap* ma = &maps.a[s.map];
        const CUtensorMap* mb = &maps.b[s.map];
        mark(i, 0);
        if (s.dep != nullptr && t.m0 < s.M) {
          flag_wait_ge(s.dep + t.m0 / kGemmBM, s.dep_target);
          fence_proxy_async_global();
        }
        mark(i, 1);
        for (int kb = 0; kb < t.nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
          const uint32_t lbar = mapa_cluster(full_bar(stage), 0);
          tma_load_2d_2cta(sa, ma, lbar, kb * kGemmBK, t.m0);
          tma_load_2d_2cta(sa + ChainCfg::kABytes, mb, lbar, kb * kGemmBK, t.n0 + (int)crank * (BN / 2));
          advance();
        }
        if (s.ln == 2) {  // this CTA's 128 rows of the residual tile: four slots of two 128 x 32 fp32 boxes
          const CUtensorMap* mr = &maps.r[s.map_r];
          for (int r = 0; r < 4; ++r) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
            mbar_expect_tx(rfull_bar(stage), Cfg::kStageBytes);
            tma_load_2d(sa, mr, rfull_bar(stage), t.n0 + r * 64, t.m0);
            tma_load_2d(sa + 16384, mr, rfull_bar(stage), t.n0 + r * 64 + 32, t.m0);
            advance();
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
      uint32_t fbits = 0;  // per ring slot: parity of its next OPERAND fill (residual fills use rfull barriers)
      for (int i = 0;; ++i) {
        const int task = task_s[i];
        if (task < 0) break;
        const ChainStage& s = p.st[task >> 24];
        const TileAt t = decode(s, task & 0xFFFFFF);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        ++it;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        mark(i, 3);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < t.nkb; ++kb) {
          mbar_wait(full_bar(stage), (fbits >> stage) & 1u);
          fbits ^= 1u << stage;
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sa + ChainCfg::kABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) umma_f16_2cta(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
          umma_commit_2cta_mc(empty_bar(stage), 3);
          if (++stage == kStages) stage = 0;
        }
        umma_commit_2cta_mc(tfull_bar(acc), 3);
        if (s.ln == 2) stage = (stage + 4) % kStages;  // the residual slots: filled and released without the tensor cores
        mark(i, 4);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    constexpr int kColsPerWarp = BN / 2;
    uint8_t* padA = epi_gen + ew * (2 * Cfg::kPad);
    const uint32_t padA_u32 = epi_base + ew * (2 * Cfg::kPad);
    uint8_t* pad16 = epi_gen + kGemmEpiWarps * 2 * Cfg::kPad + ew * Cfg::kPad16;
    const uint32_t pad16_u32 = epi_base + kGemmEpiWarps * 2 * Cfg::kPad + ew * Cfg::kPad16;
    float* sv0 = reinterpret_cast<float*>(epi_gen + kGemmEpiWarps * (2 * Cfg::kPad + Cfg::kPad16) + ew * Cfg::kVecBytes);
    float* sv1 = sv0 + kColsPerWarp;
    float* sv2 = sv1 + kColsPerWarp;
    int it = 0, cursor = 0;   // cursor: ring slot the next task's first k-block goes to
    uint32_t rbits = 0;       // per ring slot: parity of its next RESIDUAL fill
    int issued = 0;           // bulk-store groups of the current tile committed so far
    for (int i = 0;; ++i) {
      const int task = task_s[i];
      if (task < 0) break;
      const ChainStage& s = p.st[task >> 24];
      const bool tr = tracing && ew == 0 && lane == 0;
      if (tr) {
        mark(i, 5);
        p.trace[((long long)pair_id * p.pitch + i) * 16 + 8] = task;
      }
      const TileAt t = decode(s, task & 0xFFFFFF);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      ++it;
      const int res0 = (cursor + t.nkb) % kStages;  // first residual slot of an F tile
      cursor = (cursor + t.nkb + (s.ln == 2 ? 4 : 0)) % kStages;
      const int mrow0 = t.m0 + q * 32;
      const int ncol0 = t.n0 + half * kColsPerWarp;
      const int mt = t.m0 / kGemmBM;
      const bool rows_ok = mrow0 < s.M;
      const int n_live = max(0, min(kColsPerWarp, s.N - ncol0));
      const int m_pad = chain_rows_padded(s.M);
      {  // per-column vectors of this warp's 128 columns -> smem
        const int nb = ncol0 + lane * 4;
        const bool in = nb + 3 < s.N;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s.bias != nullptr) {
          if (in) {
            b4 = __ldg(reinterpret_cast<const float4*>(s.bias + nb));
          } else {
            if (nb < s.N) b4.x = __ldg(s.bias + nb);
            if (nb + 1 < s.N) b4.y = __ldg(s.bias + nb + 1);
            if (nb + 2 < s.N) b4.z = __ldg(s.bias + nb + 2);
          }
        }
        *reinterpret_cast<float4*>(sv0 + lane * 4) = b4;
        if (s.apart != nullptr)
          *reinterpret_cast<float4*>(sv1 + lane * 4) =
              in ? __ldg(reinterpret_cast<const float4*>(s.gvec + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (s.ln == 2 && s.rpart != nullptr) {
          *reinterpret_cast<float4*>(sv1 + lane * 4) =
              in ? __ldg(reinterpret_cast<const float4*>(s.gamma + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(sv2 + lane * 4) =
              in ? __ldg(reinterpret_cast<const float4*>(s.beta + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
      }
      // row statistics this epilogue needs (of the A operand's rows for a consumer, of the residual's rows for an F
      // stage): fetched now if the producing stage has already published the rows, else after the accumulator wait
      const float2* stats_src = s.ln == 2 ? s.rpart : s.apart;
      const int stats_width = s.ln == 2 ? s.N : s.K;
      float fa = 1.f, fb = 0.f;
      bool have_stats = stats_src == nullptr;
      if (!have_stats && rows_ok) {
        unsigned ok = 1u;
        if (s.dep != nullptr) {
          if (lane == 0) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(s.dep + mt) : "memory");
            ok = v >= s.dep_target ? 1u : 0u;
          }
          ok = __shfl_sync(0xffffffffu, ok, 0);
        }
        if (ok) {
          const float2 st = chain_row_stats(stats_src, stats_width, m_pad, min(mrow0 + lane, s.M - 1), s.eps);
          fa = st.y;
          fb = -st.x * st.y;
          have_stats = true;
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (tr) mark(i, 6);
      if (!have_stats && rows_ok) {
        const float2 st = chain_row_stats(stats_src, stats_width, m_pad, min(mrow0 + lane, s.M - 1), s.eps);
        fa = st.y;
        fb = -st.x * st.y;
      }
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN + half * kColsPerWarp;
      const CUtensorMap* mo = &maps.o[s.map];
      issued = 0;
      if (s.ln == 2) {
        // ================= F stage: x = acc + bias + LN_prev(residual) -> fp32 + raw 16-bit + (mean, M2) partials
        const CUtensorMap* mo32 = s.out32 ? mo : nullptr;
        const CUtensorMap* mo16 = s.map2 >= 0 ? &maps.o2[s.map2] : nullptr;
        const bool work = rows_ok && n_live > 0;
        float cnt = 0.f, mean = 0.f, m2 = 0.f;
        uint32_t rbuf[32];
        if (work) tmem_ld_32x32b_x32(t_row, rbuf);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int sidx = (res0 + 2 * half + (c >> 1)) % kStages;
          if ((c & 1) == 0) {  // this warp's two residual slots, each waited for once
            mbar_wait(rfull_bar(sidx), (rbits >> sidx) & 1u);
          }
          const int nc = ncol0 + c * 32;
          const bool live = work && nc < s.N;
          float x[32];
          if (live) {
            tmem_ld_wait();
            const uint8_t* rrow = smem_gen + sidx * Cfg::kStageBytes + (c & 1) * 16384 + (q * 32 + lane) * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 r4 = *reinterpret_cast<const float4*>(rrow + ((j ^ (lane & 7)) << 4));
              const float4 b4 = *reinterpret_cast<const float4*>(sv0 + c * 32 + 4 * j);
              if (s.rpart != nullptr) {  // LN_prev(r) = (r * fa + fb) * gamma + beta
                const float4 g4 = *reinterpret_cast<const float4*>(sv1 + c * 32 + 4 * j);
                const float4 e4 = *reinterpret_cast<const float4*>(sv2 + c * 32 + 4 * j);
                r4.x = fmaf(fmaf(r4.x, fa, fb), g4.x, e4.x);
                r4.y = fmaf(fmaf(r4.y, fa, fb), g4.y, e4.y);
                r4.z = fmaf(fmaf(r4.z, fa, fb), g4.z, e4.z);
                r4.w = fmaf(fmaf(r4.w, fa, fb), g4.w, e4.w);
              }
              x[4 * j] = __uint_as_float(rbuf[4 * j]) + b4.x + r4.x;
              x[4 * j + 1] = __uint_as_float(rbuf[4 * j + 1]) + b4.y + r4.y;
              x[4 * j + 2] = __uint_as_float(rbuf[4 * j + 2]) + b4.z + r4.z;
              x[4 * j + 3] = __uint_as_float(rbuf[4 * j + 3]) + b4.w + r4.w;
            }
            if (c + 1 < 4) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, rbuf);
          }
          __syncwarp();
          if ((c & 1) == 1 && lane == 0) {  // both boxes of the slot have been read by this warp: 4 warps share a slot
            if (atomicAdd(&rel_cnt[sidx], 1) == 3) {
              rel_cnt[sidx] = 0;
              mbar_arrive(empty_bar(sidx));
            }
          }
          if (live) {
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
            // stores: [16-bit block, group][fp32 block, group]; the fp32 staging alternates, so only the most recent
            // group (the previous chunk's fp32 store) may still be reading shared memory
            if (issued > 0) {
              if (lane == 0) tma_store_wait_read<1>();
              __syncwarp();
            }
            uint8_t* p32 = padA + (c & 1) * Cfg::kPad;
            if (mo16 != nullptr) {
              uint8_t* brow = pad16 + lane * 64;
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
            if (mo32 != nullptr) {
              uint8_t* brow = p32 + lane * 128;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(brow + ((j ^ (lane & 7)) * 16)) =
                    make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (mo16 != nullptr) {
                tma_store_2d(mo16, pad16_u32, nc, mrow0);
                tma_store_commit();
              }
              if (mo32 != nullptr) {
                tma_store_2d(mo32, padA_u32 + (c & 1) * Cfg::kPad, nc, mrow0);
                tma_store_commit();
              }
            }
            ++issued;
          }
          __syncwarp();
        }
        rbits ^= (1u << res0) | (1u << ((res0 + 1) % kStages)) | (1u << ((res0 + 2) % kStages)) |
                 (1u << ((res0 + 3) % kStages));
        if (work) tmem_ld_wait();
        if (rows_ok)
          s.part[(long long)((t.n0 / BN) * 2 + half) * m_pad + mrow0 + lane] = make_float2(mean, m2);
        __threadfence();  // the partials are ordered before this warp's "rows published" increment below
      } else {
        // ================= C stage: out = act(fa * acc + fb * g_n + c_n) as 16 bits (fa = 1, fb = 0 without the fold)
        if (rows_ok) {
          uint32_t rbuf[32];
          tmem_ld_32x32b_x32(t_row, rbuf);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int nc = ncol0 + c * 32;
            const bool live = nc < s.N;
            tmem_ld_wait();
            float v[32];
            if (s.apart != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sv0 + c * 32 + j);
                const float4 g4 = *reinterpret_cast<const float4*>(sv1 + c * 32 + j);
                v[j] = fmaf(fa, __uint_as_float(rbuf[j]), fmaf(fb, g4.x, b4.x));
                v[j + 1] = fmaf(fa, __uint_as_float(rbuf[j + 1]), fmaf(fb, g4.y, b4.y));
                v[j + 2] = fmaf(fa, __uint_as_float(rbuf[j + 2]), fmaf(fb, g4.z, b4.z));
                v[j + 3] = fmaf(fa, __uint_as_float(rbuf[j + 3]), fmaf(fb, g4.w, b4.w));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sv0 + c * 32 + j);
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
            if (c + 1 < 4) tmem_ld_32x32b_x32(t_row + (c + 1) * 32, rbuf);
            if (live) {
              if (issued > 0) {  // staging alternates: only the previous chunk's store may still be reading
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
              }
              uint8_t* brow = padA + (c & 1) * Cfg::kPad + lane * 64;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 u;
                u.x = Cvt<T16>::pack2(v[8 * j + 0], v[8 * j + 1]);
                u.y = Cvt<T16>::pack2(v[8 * j + 2], v[8 * j + 3]);
                u.z = Cvt<T16>::pack2(v[8 * j + 4], v[8 * j + 5]);
                u.w = Cvt<T16>::pack2(v[8 * j + 6], v[8 * j + 7]);
                *reinterpret_cast<uint4*>(brow + ((j ^ ((lane >> 1) & 3)) * 16)) = u;
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(mo, padA_u32 + (c & 1) * Cfg::kPad, nc, mrow0);
                tma_store_commit();
              }
              ++issued;
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
        if (s.done != nullptr) {
          tma_store_wait<0>();
          fence_proxy_async_global();
          flag_add_release(s.done + mt, 1u);
        } else {
          tma_store_wait_read<0>();
        }
        if (tr) mark(i, 7);
      }
      __syncwarp();
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace cptk
