// Host side of the dataflow chain kernel (chain_sm100.cuh): the list scheduler that turns a sequence of dependent
// stages into one ordered task list per CTA pair, the launcher, and the layer chain of the encoder forward.
// Included by cpt_b200.cu after the GEMM helpers.

// One stage as the host describes it (mirrors cpt_chain_stage of the C ABI).
struct ChainStageHost {
  int kind = CHAIN_GEMM;
  int M = 0, N = 0, K = 0, gelu = 0, out_fp32 = 0, ksplit = 1;
  const void* A = nullptr;
  long long lda = 0;
  const void* W = nullptr;
  long long ldw = 0;
  const float* bias = nullptr;
  void* out = nullptr;
  long long ldo = 0;
  const float *ln_in = nullptr, *gamma = nullptr, *beta = nullptr;
  float eps = 0.f;
  float* out32 = nullptr;
  void* out16 = nullptr;
  int dep_stage = -1;  // the stage whose output this one reads (-1: an earlier launch produced it)
  bool publish = false;  // count this stage's finished rows even without an in-launch reader (a later launch polls them)
  const unsigned* ext_dep = nullptr;  // stage 0 only: per-128-row-tile counters a still-running EARLIER launch raises
  unsigned ext_target = 0;
  // GEMM with the LayerNorm in its epilogue: out32 / out16 = LayerNorm(A W^T + bias + resid) gamma + beta
  int ln = 0;  // 1: LayerNorm finished in the epilogue; 2: deferred to the consumers (see ChainStage)
  const float* resid = nullptr;
  long long ldr = 0;
  float2* part = nullptr;         // ln == 2: where this stage's row-statistics partials go
  const float2* rpart = nullptr;  // ln == 2: partials of the residual's rows (NULL: residual is final)
  const float2* apart = nullptr;  // consumer: partials of the A operand's rows; bias = c vector, gvec = g vector
  const float* gvec = nullptr;
};

static int chain_n_tasks(const ChainStageHost& s) {
  if (s.kind == CHAIN_LN) return (s.M + 2 * kChainLnRows - 1) / (2 * kChainLnRows);
  const int m_tiles = (s.M + kGemmBM - 1) / kGemmBM, n_tiles = (s.N + kChainBN - 1) / kChainBN;
  return ((m_tiles + 1) / 2) * n_tiles * std::max(1, s.ksplit);
}

// List scheduling: tasks in stage-major order (M fastest changing last: a stage's tiles complete M pair by M pair), each
// appended to the pair that becomes free first under a crude cost model (k-blocks of the main loop + a fixed epilogue
// share; a LayerNorm task costs about two k-blocks of epilogue-warp time).  Every list is a subsequence of ONE global
// order in which a task's producers precede it, so with all pairs resident no wait can be circular.
static int chain_schedule(cpt_handle* h, const ChainStageHost* st, int n, cudaStream_t stream, int* pairs_out,
                          const int** dev_out, int* pitch_out) {
  std::vector<int> key;
  for (int i = 0; i < n; ++i) {
    key.push_back(st[i].kind);
    key.push_back(st[i].M);
    key.push_back(st[i].N);
    key.push_back(st[i].K);
    key.push_back(st[i].ksplit);
    key.push_back(st[i].ln);
  }
  key.push_back(h->chain_groups);
  auto it = h->chain_scheds.find(key);
  if (it == h->chain_scheds.end()) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CK(cudaStreamIsCapturing(stream, &cap));
    if (cap != cudaStreamCaptureStatusNone)
      return fail("chain schedule for this shape is not built yet: run the call once outside CUDA-graph capture");
    long long total = 0;
    for (int i = 0; i < n; ++i) total += chain_n_tasks(st[i]);
    int pairs = h->num_sms / 2;
    if (total < pairs) pairs = (int)std::max<long long>(1, total);
    std::vector<std::vector<int>> lists(pairs);
    std::vector<double> avail(pairs, 0.0);
    std::vector<int> last_key(pairs, -2);
    // Global task order.  groups == 1: stage-major.  groups > 1: the M pairs are cut into `groups` contiguous groups and
    // the (stage, group) blocks are issued along anti-diagonals (stage + group = const), the blocks of one diagonal
    // interleaved proportionally — a software pipeline over row groups: while group 0 is in a LayerNorm stage (epilogue
    // warps only) group 1 is in the GEMM stage before it (tensor cores), and a narrow stage of one group (90 tiles on 74
    // pairs) shares the machine with a wide stage of the other.  A task's producers ((stage - 1, same group)) sit on an
    // earlier diagonal, so every list is still a subsequence of one order in which producers come first.
    const int groups = std::max(1, std::min(h->chain_groups, 8));
    struct Item { double key; int stage, idx; };
    std::vector<Item> order;
    {
      const int m_pairs = (((st[0].M + kGemmBM - 1) / kGemmBM) + 1) / 2;
      for (int i = 0; i < n; ++i) {
        const int nt = chain_n_tasks(st[i]);
        const int n_tiles = st[i].kind == CHAIN_GEMM ? (st[i].N + kChainBN - 1) / kChainBN : 1;
        const int mn = m_pairs * n_tiles;
        std::vector<int> cnt(groups, 0), seen(groups, 0);
        auto group_of = [&](int t) {
          const int mp = st[i].kind == CHAIN_GEMM ? (t % mn) / n_tiles : std::min(m_pairs - 1, t / 4);
          return std::min(groups - 1, mp * groups / std::max(1, m_pairs));
        };
        for (int t = 0; t < nt; ++t) cnt[group_of(t)]++;
        for (int t = 0; t < nt; ++t) {
          const int g = group_of(t);
          const double frac = (seen[g]++ + 0.5) / cnt[g];
          order.push_back(Item{(double)(i + g) + frac * 0.999, i, t});
        }
      }
      if (groups > 1)
        std::stable_sort(order.begin(), order.end(), [](const Item& a, const Item& b) { return a.key < b.key; });
    }
    std::vector<std::vector<int>> group_used(n);  // pairs holding a tile of the current fused-LayerNorm group, per stage
    std::vector<int> group_id(n, -1);
    for (const Item& it : order) {
      const int i = it.stage, t = it.idx;
      double cost = 2.0;
      if (st[i].kind == CHAIN_GEMM)
        cost = (double)((st[i].K + kGemmBK - 1) / kGemmBK) / std::max(1, st[i].ksplit) + 4.0;
      // the N tiles of one M pair of a fused-LayerNorm stage wait for each other's row statistics inside their
      // epilogues: they must sit on DISTINCT pairs (a pair runs its list in order)
      const int group = (st[i].kind == CHAIN_GEMM && st[i].ln) ? (st[i].N + kChainBN - 1) / kChainBN : 1;
      if (group > pairs) return fail("chain: a fused LayerNorm stage needs at least %d CTA pairs", group);
      if (group > 1 && t / group != group_id[i]) {
        group_id[i] = t / group;
        group_used[i].clear();
      }
      // Affinity: a pair that just ran a tile of the same stage and the same N tile keeps its per-column vectors (the
      // kernel's fetch warp skips reloading them) and finds the weight tile warm: it wins ties and near-ties.
      const int n_tiles_i = st[i].kind == CHAIN_GEMM ? (st[i].N + kChainBN - 1) / kChainBN : 1;
      const int my_key = st[i].kind == CHAIN_GEMM ? (i << 16) | (t % n_tiles_i) : -1;
      int best = -1;
      double best_score = 0.0;
      for (int p = 0; p < pairs; ++p) {
        if (group > 1 && std::find(group_used[i].begin(), group_used[i].end(), p) != group_used[i].end()) continue;
        const double score = avail[p] - ((my_key >= 0 && last_key[p] == my_key) ? 2.0 : 0.0);
        if (best < 0 || score < best_score - 1e-9) {
          best = p;
          best_score = score;
        }
      }
      last_key[best] = my_key;
      if (group > 1) group_used[i].push_back(best);
      lists[best].push_back((i << 24) | t);
      avail[best] += cost + (group > 1 ? 6.0 : 0.0);
    }
    size_t longest = 0;
    for (auto& l : lists) longest = std::max(longest, l.size());
    const int pitch = (int)longest + 1;
    if (pitch > kChainMaxTasks) return fail("chain: %d tasks per CTA pair exceed the %d the kernel stages", pitch, kChainMaxTasks);
    std::vector<int> flat((size_t)pairs * pitch, -1);
    for (int p = 0; p < pairs; ++p) std::copy(lists[p].begin(), lists[p].end(), flat.begin() + (size_t)p * pitch);
    cpt_handle::ChainSched s;
    s.pairs = pairs;
    s.pitch = pitch;
    CK(cudaMalloc((void**)&s.dev, flat.size() * sizeof(int)));
    CK(cudaMemcpy(s.dev, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice));
    h->owned_chain.push_back(s.dev);
    it = h->chain_scheds.emplace(key, s).first;
  }
  *pairs_out = it->second.pairs;
  *dev_out = it->second.dev;
  *pitch_out = it->second.pitch;
  return 0;
}

// Runs the stages as one launch.  `counters`: chain_counter_bytes(M, n) of ZEROED device memory (the caller's: the
// encoder forward clears its whole flag area once per call).
template <typename T16>
static int run_chain(cpt_handle* h, cudaStream_t st, const ChainStageHost* hs, int n, unsigned* counters,
                     float2* part) {
  if (n <= 0) return 0;
  if (n > kChainMaxStages) return fail("chain: at most %d stages", kChainMaxStages);
  const int dt = Cvt<T16>::kFmt;
  const int M = hs[0].M;
  const int per = chain_m_tiles2(M);
  ChainMaps maps;
  memset(&maps, 0, sizeof maps);
  ChainParams p;
  memset(&p, 0, sizeof p);
  p.n_stages = n;
  int n_maps = 0, n_fused = 0;
  // the production kernel (chain2_sm100.cuh) handles launches made of deferred-LayerNorm stages and 16-bit-output GEMM
  // stages only; its residual tiles travel through the operand ring as 128-row boxes
  bool lean = h->chain_lean != 0;
  for (int i = 0; i < n; ++i)
    lean = lean && hs[i].kind == CHAIN_GEMM && (hs[i].ln == 2 || (hs[i].ln == 0 && !hs[i].out_fp32)) && hs[i].ksplit <= 1;
  for (int i = 0; i < n; ++i) {
    const ChainStageHost& s = hs[i];
    ChainStage& d = p.st[i];
    if (s.M != M) return fail("chain: every stage must have the same row count");
    d.kind = s.kind;
    d.M = s.M;
    d.N = s.N;
    d.K = s.K;
    d.gelu = s.gelu;
    d.out_fp32 = s.out_fp32;
    d.ksplit = std::max(1, s.ksplit);
    d.bias = s.bias;
    d.done = nullptr;  // set below when a later stage of this launch reads this one
    d.map2 = -1;
    if (s.ext_dep != nullptr) {
      d.dep = s.ext_dep;
      d.dep_target = s.ext_target;
    }
    if (s.dep_stage >= 0) {
      if (s.dep_stage >= i) return fail("chain: stage %d depends on a later stage", i);
      const ChainStageHost& ps = hs[s.dep_stage];
      d.dep = counters + (size_t)(2 * s.dep_stage) * per;
      d.dep_target = ps.kind == CHAIN_LN
                         ? 0u
                         : (unsigned)(((ps.N + kChainBN - 1) / kChainBN) * std::max(1, ps.ksplit) * kGemmEpiWarps);
    }
    if (s.kind == CHAIN_LN) {
      if (!(s.N == 128 || s.N == 256 || s.N == 512 || s.N == 768 || s.N == 1024))
        return fail("chain LayerNorm: unsupported row width %d (128, 256, 512, 768 or 1024)", s.N);
      if (!s.ln_in || !s.gamma || !s.beta) return fail("chain LayerNorm: NULL argument");
      d.ln_in = s.ln_in;
      d.gamma = s.gamma;
      d.beta = s.beta;
      d.eps = s.eps;
      d.out32 = s.out32;
      d.out16 = s.out16;
      continue;
    }
    if (n_maps >= kChainMaxMaps) return fail("chain: at most %d GEMM stages", kChainMaxMaps);
    if (s.ln == 2) {
      if (s.N % 128 || s.N <= 0) return fail("chain deferred LayerNorm: row width %d is not a multiple of 128", s.N);
      if (!s.A || !s.W || !s.resid || !s.part || (!s.out32 && !s.out16) || s.M <= 0 || s.K <= 0)
        return fail("chain deferred LayerNorm: bad argument");
      if (s.rpart && (!s.gamma || !s.beta)) return fail("chain deferred LayerNorm: residual LayerNorm needs gamma / beta");
      d.ln = 2;
      d.ksplit = 1;
      d.resid = s.resid;
      d.ldr = s.ldr;
      d.rpart = s.rpart;
      d.gamma = s.gamma;
      d.beta = s.beta;
      d.eps = s.eps;
      d.out32 = s.out32;
      d.out16 = s.out16;
      d.part = s.part;
      d.map = d.map_r = n_maps;
      TRY(make_tmap(&maps.a[n_maps], s.A, dt, s.M, s.K, s.lda, kGemmBM));
      TRY(make_tmap(&maps.b[n_maps], s.W, dt, s.N, s.K, s.ldw, kChainBN / 2));
      TRY(make_tmap_ex(&maps.r[n_maps], s.resid, 2, s.M, s.N, s.ldr, 32, lean ? kGemmBM : 32, CU_TENSOR_MAP_SWIZZLE_128B));
      if (s.out32) TRY(make_tmap_ex(&maps.o[n_maps], s.out32, 2, s.M, s.N, s.N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
      if (s.out16) {
        TRY(make_tmap_ex(&maps.o2[n_maps], s.out16, dt, s.M, s.N, s.N, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
        d.map2 = n_maps;
      }
      ++n_maps;
      continue;
    }
    if (s.ln) {
      if (!(s.N == 128 || s.N == 256 || s.N == 512 || s.N == 768 || s.N == 1024))
        return fail("chain fused LayerNorm: unsupported row width %d", s.N);
      if (!s.A || !s.W || !s.resid || !s.gamma || !s.beta || (!s.out32 && !s.out16) || s.M <= 0 || s.K <= 0)
        return fail("chain fused LayerNorm: bad argument");
      if ((s.ldr & 3) || (reinterpret_cast<uintptr_t>(s.resid) & 15)) return fail("chain fused LayerNorm: residual alignment");
      if (n_fused >= 2 || !part) return fail("chain: at most two fused LayerNorm stages per launch");
      d.ln = 1;
      d.ksplit = 1;
      d.resid = s.resid;
      d.ldr = s.ldr;
      d.gamma = s.gamma;
      d.beta = s.beta;
      d.eps = s.eps;
      d.out32 = s.out32;
      d.out16 = s.out16;
      d.sflag = counters + (size_t)(2 * i + 1) * per;
      d.part = part + (size_t)n_fused * (2 * ((s.N + kChainBN - 1) / kChainBN)) * chain_rows_padded(s.M);
      ++n_fused;
      d.map = n_maps;
      TRY(make_tmap(&maps.a[n_maps], s.A, dt, s.M, s.K, s.lda, kGemmBM));
      TRY(make_tmap(&maps.b[n_maps], s.W, dt, s.N, s.K, s.ldw, kChainBN / 2));
      if (s.out32) TRY(make_tmap_ex(&maps.o[n_maps], s.out32, 2, s.M, s.N, s.N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
      if (s.out16) {
        TRY(make_tmap_ex(&maps.o2[n_maps], s.out16, dt, s.M, s.N, s.N, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
        d.map2 = n_maps;
      }
      ++n_maps;
      continue;
    }
    if (!s.A || !s.W || !s.out || s.M <= 0 || s.N <= 0 || s.K <= 0) return fail("chain GEMM: bad argument");
    if (s.apart) {
      if (!s.gvec || !s.bias || s.K % 128) return fail("chain GEMM: folded LayerNorm needs g / c vectors and K %% 128 == 0");
      d.apart = s.apart;
      d.gvec = s.gvec;
      d.eps = s.eps;
    }
    if (d.ksplit > 1 && !s.out_fp32) return fail("chain GEMM: split-K needs the accumulate-into-fp32 output");
    if (d.ksplit > 1) d.ksplit = std::max(1, std::min(d.ksplit, ((s.K + kGemmBK - 1) / kGemmBK) / 4));
    d.map = n_maps;
    TRY(make_tmap(&maps.a[n_maps], s.A, dt, s.M, s.K, s.lda, kGemmBM));
    TRY(make_tmap(&maps.b[n_maps], s.W, dt, s.N, s.K, s.ldw, kChainBN / 2));
    if (s.out_fp32)
      TRY(make_tmap_ex(&maps.o[n_maps], s.out, 2, s.M, s.N, s.ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
    else
      TRY(make_tmap_ex(&maps.o[n_maps], s.out, dt, s.M, s.N, s.ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
    ++n_maps;
  }
  for (int i = 0; i < n; ++i)
    if (hs[i].dep_stage >= 0) p.st[hs[i].dep_stage].done = counters + (size_t)(2 * hs[i].dep_stage) * per;
  for (int i = 0; i < n; ++i)   // a stage whose rows a LATER LAUNCH reads through the counters (attention after QKV)
    if (hs[i].publish) {
      p.st[i].done = counters + (size_t)(2 * i) * per;
      h->pub_ready = p.st[i].done;
      h->pub_target = (unsigned)(((p.st[i].N + kChainBN - 1) / kChainBN) * std::max(1, p.st[i].ksplit) * kGemmEpiWarps);
    }
  // the schedule depends on the clamped ksplit: key it on what the kernel will decode
  std::vector<ChainStageHost> eff(hs, hs + n);
  for (int i = 0; i < n; ++i) eff[i].ksplit = p.st[i].ksplit;
  for (int i = 0; i < n; ++i)  // a GEMM producer's target counts its (clamped) K pieces
    if (hs[i].dep_stage >= 0 && eff[hs[i].dep_stage].kind == CHAIN_GEMM)
      p.st[i].dep_target = (unsigned)(((eff[hs[i].dep_stage].N + kChainBN - 1) / kChainBN) * eff[hs[i].dep_stage].ksplit *
                                      kGemmEpiWarps);
  int pairs = 0;
  TRY(chain_schedule(h, eff.data(), n, st, &pairs, &p.tasks, &p.pitch));
  if (h->chain_trace_on) {  // event log of this launch (debug): sized for the largest launch seen
    const size_t need = (size_t)pairs * p.pitch * 16 * sizeof(long long), hdr = (size_t)pairs * 2 * sizeof(long long);
    if (need + hdr > h->chain_trace_bytes) {
      CK(cudaStreamSynchronize(st));
      if (h->chain_trace) cudaFree(h->chain_trace);
      h->chain_trace = nullptr;
      CK(cudaMalloc((void**)&h->chain_trace, need + hdr));
      h->chain_trace_bytes = need + hdr;
    }
    CK(cudaMemsetAsync(h->chain_trace, 0, need + hdr, st));
    p.trace_hdr = h->chain_trace;
    p.trace = h->chain_trace + (size_t)pairs * 2;
    h->chain_trace_pairs = pairs;
    h->chain_trace_pitch = p.pitch;
  }
  ProfScope ps(h, st, CPT_K_CHAIN);
  p.no_wait = (lean && hs[0].ext_dep != nullptr) ? 1 : 0;
  if (lean) {
    auto* fn = chain2_kernel<T16>;
    static bool attr_set2[64] = {};
    if (!attr_set2[h->device & 63]) {
      TRY(set_smem_attr(fn, Chain2Cfg::kSmemBytes));
      attr_set2[h->device & 63] = true;
    }
    CK(launch_k(fn, dim3(pairs * 2), dim3(Chain2Cfg::kThreads), Chain2Cfg::kSmemBytes, st, 2, maps, p));
    return 0;
  }
  auto* fn = chain_kernel<T16>;
  static bool attr_set[64] = {};
  if (!attr_set[h->device & 63]) {
    TRY(set_smem_attr(fn, ChainCfg::kSmemBytes));
    attr_set[h->device & 63] = true;
  }
  CK(launch_k(fn, dim3(pairs * 2), dim3(kGemmThreads), ChainCfg::kSmemBytes, st, 2, maps, p));
  return 0;
}

// One encoder layer after its attention kernel, plus the next layer's QKV projection, as ONE launch
// (modeling_bert.py:85 BertSelfOutput, :144 BertIntermediate, :145 BertOutput, then :38-40 of layer l+1).
//   fused LayerNorm (default): 4 stages
//     ctx16 -> [attention.output.dense + bias + h32 -> LayerNorm] -> a32 / a16 -> [intermediate.dense, GELU] -> inter16
//     -> [output.dense + bias + a32 -> LayerNorm] -> out32 (h32 or seq_out) / h16 -> [QKV of layer l+1] -> qkv16
//   CPT_B200_CHAIN_FUSE_LN=0: 6 stages, the dense outputs are ADDED into the fp32 stream by TMA reduce stores and the
//     LayerNorms run as row tasks between the GEMM stages
template <typename T16>
static int chain_layer(cpt_handle* h, cudaStream_t st, const Workspace& w, int l, int M, float* out32, bool last,
                       unsigned* counters) {
  const cpt_config& c = h->cfg;
  const int H = c.hidden_size, I = c.intermediate_size;
  const LayerDev& d = h->layers[l];
  const bool fuse = h->chain_fuse_ln != 0;
  ChainStageHost s[6];
  int n = 0;
  if (h->chain_fuse_ln == 2) {
    // LayerNorms deferred to their consumers: 4 GEMM stages, no LayerNorm pass and no wait inside any epilogue.
    //   x1 = ctx Wo^T + b + LN2'(x2 of layer l-1)      -> a32 (fp32, pre-LN1), a16 (raw), statistics P1
    //   inter = gelu(LN1(x1) W1^T + b1)                   LN1 folded: A = a16, W = gamma1 .* W1, epilogue correction from P1
    //   x2 = inter W2^T + b2 + LN1(x1)                  -> h32 (pre-LN2), h16 (raw), statistics P2
    //   qkv' = LN2(x2) Wqkv'^T + b                        LN2 folded into the next layer's QKV weights, correction from P2
    // (layer 0 adds the embedding output as it is; the last layer's LN2 runs as one row kernel after the loop)
    float2* P1 = w.part;
    float2* P2 = w.part + (size_t)(2 * ((H + kChainBN - 1) / kChainBN)) * chain_rows_padded(M);
    {
      ChainStageHost& g = s[n++];
      g.M = M; g.N = H; g.K = H; g.A = w.ctx16; g.lda = H; g.W = d.w_ao; g.ldw = H; g.bias = d.b_ao;
      g.ln = 2; g.resid = w.h32; g.ldr = H; g.eps = c.layer_norm_eps; g.out32 = w.a32; g.out16 = w.a16; g.part = P1;
      if (l > 0) { g.rpart = P2; g.gamma = h->layers[l - 1].o_g; g.beta = h->layers[l - 1].o_b; }
      if (h->ctx_published != nullptr) {   // the attention launch just before counts its context rows per 128-row tile:
        g.ext_dep = h->ctx_published;      // this launch starts inside ITS tail and waits tile by tile
        g.ext_target = 128u * (unsigned)c.num_attention_heads;
        h->ctx_published = nullptr;
      }
    }
    {
      ChainStageHost& g = s[n++];
      g.M = M; g.N = I; g.K = H; g.gelu = 1; g.A = w.a16; g.lda = H; g.W = d.w_i_f; g.ldw = H; g.bias = d.c_i;
      g.gvec = d.g_i; g.apart = P1; g.eps = c.layer_norm_eps; g.out = w.inter16; g.ldo = I; g.dep_stage = 0;
    }
    {
      ChainStageHost& g = s[n++];
      g.M = M; g.N = H; g.K = I; g.A = w.inter16; g.lda = I; g.W = d.w_o; g.ldw = I; g.bias = d.b_o; g.dep_stage = 1;
      g.ln = 2; g.resid = w.a32; g.ldr = H; g.rpart = P1; g.gamma = d.ao_g; g.beta = d.ao_b; g.eps = c.layer_norm_eps;
      g.out32 = w.h32; g.out16 = last ? nullptr : w.h16; g.part = P2;
    }
    if (!last) {
      const LayerDev& nx = h->layers[l + 1];
      ChainStageHost& g = s[n++];
      g.M = M; g.N = 3 * H; g.K = H; g.A = w.h16; g.lda = H; g.W = nx.w_qkv_f; g.ldw = H; g.bias = nx.c_qkv;
      g.gvec = nx.g_qkv; g.apart = P2; g.eps = c.layer_norm_eps; g.out = w.qkv16; g.ldo = 3 * H; g.dep_stage = 2;
      g.publish = h->attn_early != 0;   // the next attention launch starts on rows as they are published
    }
    return run_chain<T16>(h, st, s, n, counters, w.part);
  }
  {  // attention.output.dense (+bias) + residual (+ BertSelfOutput.LayerNorm)
    ChainStageHost& g = s[n++];
    g.M = M; g.N = H; g.K = H; g.A = w.ctx16; g.lda = H; g.W = d.w_ao; g.ldw = H; g.bias = d.b_ao;
    if (fuse) {
      g.ln = 1; g.resid = w.h32; g.ldr = H; g.gamma = d.ao_g; g.beta = d.ao_b; g.eps = c.layer_norm_eps;
      g.out32 = w.a32; g.out16 = w.a16;
    } else {
      g.out_fp32 = 1; g.out = w.h32; g.ldo = H;
    }
  }
  if (!fuse) {  // BertSelfOutput.LayerNorm
    ChainStageHost& g = s[n++];
    g.kind = CHAIN_LN; g.M = M; g.N = H; g.ln_in = w.h32; g.gamma = d.ao_g; g.beta = d.ao_b; g.eps = c.layer_norm_eps;
    g.out32 = w.a32; g.out16 = w.a16; g.dep_stage = n - 2;
  }
  {  // intermediate.dense + bias + erf-GELU
    ChainStageHost& g = s[n++];
    g.M = M; g.N = I; g.K = H; g.gelu = 1; g.A = w.a16; g.lda = H; g.W = d.w_i; g.ldw = H; g.bias = d.b_i;
    g.out = w.inter16; g.ldo = I; g.dep_stage = n - 2;
  }
  {  // output.dense (+bias) + residual (+ BertOutput.LayerNorm)
    ChainStageHost& g = s[n++];
    g.M = M; g.N = H; g.K = I; g.A = w.inter16; g.lda = I; g.W = d.w_o; g.ldw = I; g.bias = d.b_o; g.dep_stage = n - 2;
    if (fuse) {
      g.ln = 1; g.resid = w.a32; g.ldr = H; g.gamma = d.o_g; g.beta = d.o_b; g.eps = c.layer_norm_eps;
      g.out32 = out32; g.out16 = last ? nullptr : w.h16;
    } else {
      g.out_fp32 = 1; g.ksplit = h->chain_down_ksplit; g.out = w.a32; g.ldo = H;
    }
  }
  if (!fuse) {  // BertOutput.LayerNorm
    ChainStageHost& g = s[n++];
    g.kind = CHAIN_LN; g.M = M; g.N = H; g.ln_in = w.a32; g.gamma = d.o_g; g.beta = d.o_b; g.eps = c.layer_norm_eps;
    g.out32 = out32; g.out16 = last ? nullptr : w.h16; g.dep_stage = n - 2;
  }
  if (!last) {  // the next layer's query / key / value projection
    const LayerDev& nx = h->layers[l + 1];
    ChainStageHost& g = s[n++];
    g.M = M; g.N = 3 * H; g.K = H; g.A = w.h16; g.lda = H; g.W = nx.w_qkv; g.ldw = H; g.bias = nx.b_qkv;
    g.out = w.qkv16; g.ldo = 3 * H; g.dep_stage = n - 2;
  }
  return run_chain<T16>(h, st, s, n, counters, w.part);
}
