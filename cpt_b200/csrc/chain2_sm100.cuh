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
//   * stores are TMA bulk stores out of double-buffered staging blocks (cp.async.bulk.wait_group.read 1);
//   * the per-column vectors (bias, gamma-sum or gamma, beta) and the per-row statistics (rstd, -mean * rstd) of a tile
//     are fetched by a dedicated FETCH warp one task ahead, into double-buffered shared memory, and handed over through
//     mbarriers.  (Second event log: with four operand stages in flight the SM's path to L2 holds ~1000 TMA line
//     requests, so ANY load an epilogue warp issues waits ~1.2 us behind them — two dependent ones per tile, vectors
//     then statistics, cost 2.5 us of the 6.3 us tile period of the FFN-up stage.  prefetch.global.L1 did not help.)
#pragma once
#include "chain_sm100.cuh"

namespace cptk {

struct Chain2Cfg {
  static constexpr int kStages = 4;
  static constexpr int kStageBytes = ChainCfg::kStageBytes;     // 32 KB: A 128 x 64 + B 128 x 64 (16-bit) = a residual slot
  static constexpr int kPad = 4096;                             // two of these per warp: staging blocks
  static constexpr int kPad16 = 2048;
  static constexpr int kThreads = kGemmThreads + 32;            // + the fetch warp
  static constexpr int kVecBytes = 3 * kChainBN * 4;            // bias | gamma(-sum) | beta of the tile's 256 columns
  static constexpr int kStatBytes = kGemmBM * 8;                // (rstd, -mean * rstd) of this CTA's 128 rows
  static constexpr int kRawPlanes = 6;                          // landing zone of the (mean, M2) planes, 1 KB each
  static constexpr int kXchBytes = kGemmBM * 8;                 // F stage: column half 1 hands (mean, M2) to half 0
  static constexpr int kFetchBytes = 2 * (kVecBytes + kStatBytes) + kRawPlanes * kStatBytes + kXchBytes;
  static constexpr int kEpiBytes = kGemmEpiWarps * (2 * kPad + kPad16) + kFetchBytes;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiBytes + 512 + kChainMaxTasks * 4;
  static_assert(kSmemBytes <= kSmemLimit, "shared memory");
};

template <typename T16>
__global__ void __launch_bounds__(Chain2Cfg::kThreads, 1)
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
  auto vfull_bar = [&](int b) { return bars + 8u * (3 * kStages + 6) + 4u * (kStages + 2) + 8u * b; };
  auto vfree_bar = [&](int b) { return bars + 8u * (3 * kStages + 6) + 4u * (kStages + 2) + 16u + 8u * b; };
  const uint32_t fbar = bars + 8u * (3 * kStages + 6) + 4u * (kStages + 2) + 32u;  // the fetch warp's own copies
  // the statistics half of a buffer is free again as soon as the epilogue warps have read their two numbers (at the top
  // of a tile), a whole tile earlier than the vectors: the slow part of a fetch (dependency + planes) starts that early
  auto sfree_bar = [&](int b) { return bars + 8u * (3 * kStages + 6) + 4u * (kStages + 2) + 56u + 8u * b; };
  // index of the first task whose dependency wait this CTA's producer has NOT passed yet (release.cta by the producer
  // thread, acquire.cta by the fetch warp: the rows a task's statistics describe are published once it is passed)
  const uint32_t ready_u32 = bars + 8u * (3 * kStages + 6) + 4u * kStages;
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
    if (threadIdx.x <= kStages) rel_cnt[threadIdx.x] = 0;  // [kStages] = the word at ready_u32
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
        mbar_init(vfull_bar(a), 1);
        mbar_init(vfree_bar(a), kGemmEpiWarps);
        mbar_init(sfree_bar(a), kGemmEpiWarps);
      }
      for (int a = 0; a < 3; ++a) mbar_init(fbar + 8u * a, 1);   // landing zones 0 / 1, vectors
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
  if (!p.no_wait) pdl_wait();
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
        asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(ready_u32), "r"(i + 1) : "memory");
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
  } else if (warp == 2 + kGemmEpiWarps) {
    // ------------------------------------------------------------------ fetch warp: vectors + row statistics, ahead of the epilogue
    // Everything arrives by BULK COPY (the TMA path).  Measured here: with the operand ring full on every SM about
    // 19 MB of requests are in flight chip-wide, so any request — TMA or LSU — waits ~3 us (Little's law), and LSU loads
    // additionally trickle at ~0.1 us per line.  So the fetch must be a whole task AHEAD: the (mean, M2) planes of this
    // CTA's 128 rows (one 1 KB plane per 256 columns of the normalised width — the F stage merges its two column
    // halves before writing) land in one of two zones; task i + 1's copy is in flight while task i's is merged.
    constexpr int kZonePlanes = Cfg::kRawPlanes / 2;
    const uint32_t fetch_u32 = epi_base + kGemmEpiWarps * (2 * Cfg::kPad + Cfg::kPad16);
    uint8_t* fetch_gen = epi_gen + kGemmEpiWarps * (2 * Cfg::kPad + Cfg::kPad16);
    const uint32_t zone_u32 = fetch_u32 + 2 * (Cfg::kVecBytes + Cfg::kStatBytes);
    const float2* zone = reinterpret_cast<const float2*>(fetch_gen + 2 * (Cfg::kVecBytes + Cfg::kStatBytes));
    auto stats_of = [&](const ChainStage& st) { return st.ln == 2 ? st.rpart : st.apart; };
    auto parts_of = [&](const ChainStage& st, const TileAt& tt) {
      if (stats_of(st) == nullptr || tt.m0 >= st.M) return 0;
      const int w = st.ln == 2 ? st.N : st.K;
      return (w + BN - 1) / BN;
    };
    uint32_t zphase[2] = {0u, 0u};
    uint32_t vphase = 0;
    int held[2] = {-1, -1};   // (stage, N tile) whose vectors each buffer holds
    // copy planes [p0, p0 + nb) of task j's rows into zone z (lane 0 issues; the zone's barrier completes when landed)
    auto issue_planes = [&](const ChainStage& st, const TileAt& tt, int p0, int nb, int z) {
      if (lane == 0) {
        const uint32_t zb = fbar + 8u * z;
        const long long m_pad = chain_rows_padded(st.M);
        fence_proxy_async_global();
        mbar_expect_tx(zb, (uint32_t)nb * Cfg::kStatBytes);
        for (int j = 0; j < nb; ++j)
          bulk_load_1d(zone_u32 + (z * kZonePlanes + j) * Cfg::kStatBytes,
                       stats_of(st) + (long long)(p0 + j) * m_pad + tt.m0, Cfg::kStatBytes, zb);
      }
      __syncwarp();
    };
    auto rows_published = [&](int j, bool block) {   // has this CTA's producer passed task j's dependency wait?
      unsigned rdy = 0, spins = 0;
      for (;;) {
        if (lane == 0) asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(rdy) : "r"(ready_u32) : "memory");
        rdy = __shfl_sync(0xffffffffu, rdy, 0);
        if ((int)rdy > j) return true;
        if (!block) return false;
        if (++spins > (1u << 24)) __trap();
        __nanosleep(32);
      }
    };
    // first batch of task j -> zone j & 1, if it has planes and (block or already) its rows are published
    auto prefetch = [&](int j, bool block) {
      const int task = task_s[j];
      if (task < 0) return true;
      const ChainStage& st = p.st[task >> 24];
      const TileAt tt = decode(st, task & 0xFFFFFF);
      const int np = parts_of(st, tt);
      if (np == 0) return true;
      if (st.dep != nullptr && !rows_published(j, block)) return false;
      issue_planes(st, tt, 0, min(np, kZonePlanes), j & 1);
      return true;
    };
    bool next_issued = prefetch(0, true);
    for (int i = 0;; ++i) {
      const int task = task_s[i];
      if (task < 0) break;
      const ChainStage& s = p.st[task >> 24];
      const TileAt t = decode(s, task & 0xFFFFFF);
      const int buf = i & 1;
      if (lane == 0) mark(i, 11);
      float* vec = reinterpret_cast<float*>(fetch_gen + buf * (Cfg::kVecBytes + Cfg::kStatBytes));
      float2* stat = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(vec) + Cfg::kVecBytes);
      const uint32_t vec_u32 = fetch_u32 + buf * (Cfg::kVecBytes + Cfg::kStatBytes);
      const bool fold = s.apart != nullptr, resid_ln = s.ln == 2 && s.rpart != nullptr;
      const int n_parts = parts_of(s, t);
      const int stats_width = s.ln == 2 ? s.N : s.K;
      const bool whole = t.n0 + BN <= s.N;   // the tile's 256 columns exist: 1 KB per vector, 16-byte aligned
      // task i's first batch is in flight (or landed) in zone i & 1; start task i + 1's into the other zone if its rows
      // are already known to be published — otherwise after this task has been handed over
      next_issued = prefetch(i + 1, false);
      float cnt[kGemmBM / 32], mean[kGemmBM / 32], m2[kGemmBM / 32];
#pragma unroll
      for (int k = 0; k < kGemmBM / 32; ++k) cnt[k] = mean[k] = m2[k] = 0.f;
      for (int p0 = 0; p0 < n_parts; p0 += kZonePlanes) {
        const int nb = min(kZonePlanes, n_parts - p0);
        if (p0 > 0) issue_planes(s, t, p0, nb, buf);   // hidden sizes above 768: further batches, serially
        mbar_wait(fbar + 8u * buf, zphase[buf]);
        zphase[buf] ^= 1u;
        const float2* z = zone + buf * kZonePlanes * kGemmBM;
        for (int j = 0; j < nb; ++j) {
          const float nbv = (float)min(BN, stats_width - (p0 + j) * BN);
#pragma unroll
          for (int k = 0; k < kGemmBM / 32; ++k) {
            const float2 pmv = z[j * kGemmBM + k * 32 + lane];
            const float tot = cnt[k] + nbv, delta = pmv.x - mean[k];
            mean[k] += delta * (nbv / tot);
            m2[k] += pmv.y + delta * delta * (cnt[k] * nbv / tot);
            cnt[k] = tot;
          }
        }
        __syncwarp();   // the zone may be overwritten
      }
      if (lane == 0) mark(i, 12);
      // the epilogue warps have read what task i - 2 left here.  Waited for on EVERY task: it is what keeps this warp
      // less than two tasks ahead (the "full" barriers below must never run two phases ahead of their waiters)
      mbar_wait(sfree_bar(buf), ((i >> 1) & 1u) ^ 1u);
      if (n_parts > 0) {
#pragma unroll
        for (int k = 0; k < kGemmBM / 32; ++k) {
          const float rstd = 1.0f / sqrtf(m2[k] / (float)stats_width + s.eps);
          stat[k * 32 + lane] = make_float2(rstd, -mean[k] * rstd);
        }
      }
      // ---- the vectors: reloaded only when the (stage, N tile) changed; their buffer is free once the epilogue warps
      // have finished task i - 2 (a parity wait stays correct when earlier phases were not waited for: the barrier can
      // be at most one phase ahead, because the epilogue of task i needs this task's "full" signal)
      if (lane == 0) mark(i, 14);
      const int vkey = ((task >> 24) << 20) | (t.n0 / BN);
      const bool reload = held[buf] != vkey;
      held[buf] = vkey;
      if (reload) mbar_wait(vfree_bar(buf), ((i >> 1) & 1u) ^ 1u);
      if (lane == 0) mark(i, 15);
      if (reload && (!whole || s.bias == nullptr)) {   // edge tiles (N < 256), bias-less stages: through the LSU, zero-filled
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int nb = t.n0 + h2 * (BN / 2) + lane * 4;
          float vv[3][4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool in = nb + e < s.N;
            vv[0][e] = (in && s.bias != nullptr) ? __ldg(s.bias + nb + e) : 0.f;
            vv[1][e] = !in ? 0.f : fold ? __ldg(s.gvec + nb + e) : resid_ln ? __ldg(s.gamma + nb + e) : 0.f;
            vv[2][e] = (in && resid_ln) ? __ldg(s.beta + nb + e) : 0.f;
          }
          *reinterpret_cast<float4*>(vec + h2 * (BN / 2) + lane * 4) = make_float4(vv[0][0], vv[0][1], vv[0][2], vv[0][3]);
          if (!whole) {
            *reinterpret_cast<float4*>(vec + BN + h2 * (BN / 2) + lane * 4) = make_float4(vv[1][0], vv[1][1], vv[1][2], vv[1][3]);
            *reinterpret_cast<float4*>(vec + 2 * BN + h2 * (BN / 2) + lane * 4) = make_float4(vv[2][0], vv[2][1], vv[2][2], vv[2][3]);
          }
        }
      }
      if (reload && whole && (s.bias != nullptr || fold || resid_ln)) {
        const uint32_t vb = fbar + 16u;
        if (lane == 0) {
          const uint32_t bytes = (s.bias != nullptr ? 1024u : 0u) + (fold ? 1024u : 0u) + (resid_ln ? 2048u : 0u);
          mbar_expect_tx(vb, bytes);
          if (s.bias != nullptr) bulk_load_1d(vec_u32, s.bias + t.n0, 1024u, vb);
          if (fold) bulk_load_1d(vec_u32 + BN * 4, s.gvec + t.n0, 1024u, vb);
          if (resid_ln) {
            bulk_load_1d(vec_u32 + BN * 4, s.gamma + t.n0, 1024u, vb);
            bulk_load_1d(vec_u32 + 2 * BN * 4, s.beta + t.n0, 1024u, vb);
          }
        }
        mbar_wait(vb, vphase);
        vphase ^= 1u;
      }
      __syncwarp();
      if (lane == 0) mark(i, 13);
      if (lane == 0) mbar_arrive(vfull_bar(buf));   // release.cta: the shared-memory writes above are visible to waiters
      if (!next_issued) next_issued = prefetch(i + 1, true);
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
    uint8_t* fetch_gen = epi_gen + kGemmEpiWarps * (2 * Cfg::kPad + Cfg::kPad16);
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
      // vectors and row statistics of this tile: left in shared memory by the fetch warp (normally long before)
      const int vbuf = i & 1;
      mbar_wait(vfull_bar(vbuf), (i >> 1) & 1u);
      const float* sv0 = reinterpret_cast<const float*>(fetch_gen + vbuf * (Cfg::kVecBytes + Cfg::kStatBytes)) +
                         half * kColsPerWarp;
      const float* sv1 = sv0 + BN;
      const float* sv2 = sv1 + BN;
      float fa = 1.f, fb = 0.f;
      if ((s.ln == 2 ? s.rpart : s.apart) != nullptr && rows_ok) {
        const float2 st = reinterpret_cast<const float2*>(fetch_gen + vbuf * (Cfg::kVecBytes + Cfg::kStatBytes) +
                                                          Cfg::kVecBytes)[q * 32 + lane];
        fa = st.x;
        fb = st.y;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sfree_bar(vbuf));
      if (tr) mark(i, 10);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (tr) mark(i, 6);
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
        {  // one (mean, M2) plane per 256-column tile: the two column halves (warps q and q + 4) merge before writing
          float2* xch = reinterpret_cast<float2*>(fetch_gen + 2 * (Cfg::kVecBytes + Cfg::kStatBytes) +
                                                  Cfg::kRawPlanes * Cfg::kStatBytes) + q * 32 + lane;
          if (half == 1) *xch = make_float2(mean, m2);
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
          if (half == 0) {
            const float2 o = *xch;
            const float na = work ? (float)n_live : 0.f;
            const float nb = rows_ok ? (float)max(0, min(kColsPerWarp, s.N - (t.n0 + kColsPerWarp))) : 0.f;
            if (nb > 0.f) {
              const float tot = na + nb, delta = o.x - mean;
              mean += delta * (nb / tot);
              m2 += o.y + delta * delta * (na * nb / tot);
            }
            if (rows_ok) s.part[(long long)(t.n0 / BN) * m_pad + mrow0 + lane] = make_float2(mean, m2);
          }
          asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // the slot may be rewritten
        }
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
        mbar_arrive(vfree_bar(vbuf));
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
