// Host orchestration of the training step (included by cpt_b200.cu; see include/cpt_b200.h "training step").
//
// Forward = the inference kernel sequence, except that every tensor the backward needs lands in the tape instead of a
// reused workspace buffer, and FFN-up keeps its pre-activation (GELU is a separate pass).  Backward walks the layers
// in reverse.  Every matrix product is the tcgen05 GEMM of gemm_sm100.cuh in its out = A . W^T form:
//   dgrad  dX[M,Kin]    = dY[M,Nout] . W[Nout,Kin]             W = the handle's 16-bit nn.Linear weight, read in place as
//                                                              the MN-major B operand (GemmParams::trans = 2)
//   wgrad  dW[Nout,Kin] = dY[M,Nout]^T . X[M,Kin]             both operands read where they lie, row-major, through
//                                                              MN-major UMMA descriptors (GemmParams::trans); the fp32
//                                                              result is ADDED into the caller's gradient tensor by the
//                                                              TMA reduce-add store
// The residual branches are summed the same way (dgrad results reduce-added into the fp32 gradient stream).

struct TapeLayer {
  char *h16, *qkv16, *ctx16, *a16, *preup16, *inter16;
  float *x1, *x2;
};
struct Tape {
  float *ext_mask, *imgpre32, *h32, *a32, *seq32;
  char* img16;
  std::vector<TapeLayer> layers;
  // head (n labelled rows)
  char *hx16, *ht16;
  float *htd32, *htg32, *logits, *lse;
  long long ldl;
  // NSP head (n labelled samples): [CLS] rows, pooled, logits; backward scratch
  float *nx32, *npool, *nlog, *ndlog, *ndpre, *ndx;
  // backward scratch
  float *dH, *dx32, *hd32a, *hd32b, *hdx32, *dimg32, *dwimg;
  char *dx16, *big16, *big16b, *dctx16;
  char *dlog16, *hd16, *dimg16;
  int Vp;
  size_t total;
};

static Tape carve_tape(const cpt_handle* h, int B, int T, int R, int n, char* base) {
  const cpt_config& c = h->cfg;
  const size_t S = T + R, M = (size_t)B * S, H = c.hidden_size, I = c.intermediate_size, L = c.num_hidden_layers;
  const size_t Mi = (size_t)B * R, V = c.vocab_size, W = std::max(I, 3 * H);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += al(bytes);
    return p;
  };
  Tape t;
  t.Vp = (int)((V + 7) & ~size_t(7));
  t.ldl = (long long)((V + 3) & ~size_t(3));
  t.ext_mask = (float*)take(M * 4);
  t.img16 = take(Mi * h->Fp * 2);
  t.imgpre32 = (float*)take(Mi * H * 4);
  t.h32 = (float*)take(M * H * 4);
  t.a32 = (float*)take(M * H * 4);
  t.seq32 = (float*)take(M * H * 4);
  t.layers.resize(L);
  for (auto& l : t.layers) {
    l.h16 = take(M * H * 2);
    l.qkv16 = take(M * 3 * H * 2);
    l.ctx16 = take(M * H * 2);
    l.a16 = take(M * H * 2);
    l.preup16 = take(M * I * 2);
    l.inter16 = take(M * I * 2);
    l.x1 = (float*)take(M * H * 4);
    l.x2 = (float*)take(M * H * 4);
  }
  t.hx16 = take((size_t)n * H * 2);
  t.ht16 = take((size_t)n * H * 2);
  t.htd32 = (float*)take((size_t)n * H * 4);
  t.htg32 = (float*)take((size_t)n * H * 4);
  t.logits = (float*)take((size_t)n * t.ldl * 4);
  t.lse = (float*)take((size_t)n * 4);
  const size_t Cn = std::max(1, c.num_contrast_classes);
  t.nx32 = (float*)take((size_t)n * H * 4);
  t.npool = (float*)take((size_t)n * H * 4);
  t.nlog = (float*)take((size_t)n * Cn * 4);
  t.ndlog = (float*)take((size_t)n * Cn * 4);
  t.ndpre = (float*)take((size_t)n * H * 4);
  t.ndx = (float*)take((size_t)n * H * 4);
  // backward scratch
  t.dH = (float*)take(M * H * 4);
  t.dx32 = (float*)take(M * H * 4);
  t.dx16 = take(M * H * 2);
  t.big16 = take(M * W * 2);
  t.big16b = take(M * W * 2);
  t.dctx16 = take(M * H * 2);
  t.dlog16 = take((size_t)n * t.Vp * 2);
  t.hd16 = take((size_t)n * H * 2);
  t.hd32a = (float*)take((size_t)n * H * 4);
  t.hd32b = (float*)take((size_t)n * H * 4);
  t.hdx32 = (float*)take((size_t)n * H * 4);
  t.dimg32 = (float*)take(Mi * H * 4);
  t.dimg16 = take(Mi * H * 2);
  t.dwimg = (float*)take(H * (size_t)h->Fp * 4);
  t.total = off + 256;
  return t;
}

// Drop of one site (train.cuh); p <= 0 -> off
static Drop make_drop(const cpt_dropout* d, float p, unsigned site) {
  Drop r{0u, 0u, site, 0u, 1.f};
  if (!d || !(p > 0.f)) return r;
  r.seed_lo = (unsigned)(d->seed & 0xffffffffull);
  r.seed_hi = (unsigned)(d->seed >> 32);
  r.seed_dev = reinterpret_cast<const unsigned*>(d->seed_dev);
  const double t = (double)p * 4294967296.0;
  r.thresh = t >= 4294967295.0 ? 4294967295u : (unsigned)t;
  if (r.thresh == 0u) r.thresh = 1u;
  r.scale = 1.0f / (1.0f - p);
  return r;
}
enum { SITE_ATTN = 0, SITE_AO = 1, SITE_DOWN = 2, SITE_EMB_TEXT = 0xFFFF0, SITE_EMB_IMG = 0xFFFF1 };

static int ew_grid(const cpt_handle* h, long long n) {
  const long long g = (n + 255) / 256;
  return (int)std::max<long long>(1, std::min<long long>(g, 8ll * h->num_sms));
}

template <typename T>
static int colsum(cpt_handle* h, cudaStream_t st, const T* in, int M, int N, long long ld, float* out,
                  float* out1 = nullptr, float* out2 = nullptr, int seg = 0) {
  if (M <= 0 || !out) return 0;
  if (seg <= 0) seg = N;
  ProfScope ps(h, st, CPT_K_COLSUM);
  const int vec = ((ld * sizeof(T)) % (4 * sizeof(T)) == 0) && (reinterpret_cast<uintptr_t>(in) % (4 * sizeof(T)) == 0);
  colsum_kernel<T><<<dim3((N + 255) / 256, (M + kColsumRows - 1) / kColsumRows), 256, 0, st>>>(in, M, N, ld, out, out1, out2, seg, vec);
  CKL("colsum_kernel");
  return 0;
}

template <typename T16>
static int ln_bwd(cpt_handle* h, cudaStream_t st, const float* dy, const float* x, int M, int H, const float* gamma,
                  float eps, bool do_ln, float* dx32, void* dx16, float* dgamma, float* dbeta, float* dbias,
                  int rin = 0, int rout = 0, int roff = 0, Drop drop_dy = Drop{0, 0, 0, 0, 1.f},
                  Drop drop16 = Drop{0, 0, 0, 0, 1.f}) {
  if (M <= 0) return 0;
  ProfScope ps(h, st, CPT_K_LN_BWD);
  const int grid = std::min((M + 7) / 8, 2 * h->num_sms);
#define CPT_LNB_CASE(NV_)                                                                                         \
  case NV_:                                                                                                       \
    ln_bwd_kernel<T16, NV_><<<grid, 256, 0, st>>>(dy, x, M, H, gamma, eps, do_ln ? 1 : 0, dx32,                   \
                                                  reinterpret_cast<T16*>(dx16), dgamma, dbeta, dbias, rin, rout,  \
                                                  roff, drop_dy, drop16);                                         \
    break;
  switch (H / 128) {
    CPT_LNB_CASE(1) CPT_LNB_CASE(2) CPT_LNB_CASE(3) CPT_LNB_CASE(4) CPT_LNB_CASE(5) CPT_LNB_CASE(6) CPT_LNB_CASE(7)
    CPT_LNB_CASE(8)
    default: return fail("layernorm backward: unsupported hidden size %d", H);
  }
#undef CPT_LNB_CASE
  CKL("ln_bwd_kernel");
  return 0;
}

// out(+)= A[M,K] . W[N,K]^T, no bias.  accumulate -> fp32 reduce-add into `out`.
// d(qkv) from d(ctx): impl -1 = pick (tensor cores for S <= 128 without probability dropout, else CUDA cores),
// 0 = tensor cores, 1 = CUDA cores
template <typename T16>
static int attention_backward(cpt_handle* h, cudaStream_t st, const void* qkv, const void* dctx, const float* ext_mask,
                              int B, int S, void* dqkv, const cpt_dropout* dropout, float p_a, unsigned site,
                              int impl) {
  const cpt_config& c = h->cfg;
  const int H = c.hidden_size, nH = c.num_attention_heads;
  if (S < 1 || S > 256) return fail("attention backward supports 1 <= S <= 256 (got %d)", S);
  if (impl < 0) impl = h->attn_bwd_simt ? 1 : 0;
  ProfScope ps(h, st, CPT_K_ATTN_BWD);
  if (impl == 0 && S > 128) {  // two query tiles x two key tiles per (head, sample)
    CUtensorMap tq, td;
    TRY(make_tmap(&tq, qkv, Cvt<T16>::kFmt, (unsigned long long)B * S, 3ull * H, 3ull * H, 64));
    TRY(make_tmap(&td, dctx, Cvt<T16>::kFmt, (unsigned long long)B * S, (unsigned long long)H, (unsigned long long)H, 64));
    auto* fn = attn_bwd_tc2_kernel<T16>;
    static bool attr_set[64] = {};
    if (!attr_set[h->device & 63]) {
      TRY(set_smem_attr(fn, kAttnBwd2Smem));
      attr_set[h->device & 63] = true;
    }
    AttnBwdParams p{B, S, H, nH, ext_mask, dqkv, 0.125f, make_drop(dropout, p_a, site)};
    fn<<<dim3(nH, B), kAttnBwdThreads, kAttnBwd2Smem, st>>>(tq, td, p);
    CKL("attn_bwd_tc2_kernel");
    return 0;
  }
  if (impl == 0) {
    CUtensorMap tq, td;
    TRY(make_tmap(&tq, qkv, Cvt<T16>::kFmt, (unsigned long long)B * S, 3ull * H, 3ull * H, 64));
    TRY(make_tmap(&td, dctx, Cvt<T16>::kFmt, (unsigned long long)B * S, (unsigned long long)H, (unsigned long long)H, 64));
    auto* fn = attn_bwd_tc_kernel<T16>;
    static bool attr_set[64] = {};
    if (!attr_set[h->device & 63]) {
      TRY(set_smem_attr(fn, kAttnBwdSmem));
      attr_set[h->device & 63] = true;
    }
    AttnBwdParams p{B, S, H, nH, ext_mask, dqkv, 0.125f, make_drop(dropout, p_a, site)};
    fn<<<dim3(nH, B), kAttnBwdThreads, kAttnBwdSmem, st>>>(tq, td, p);
    CKL("attn_bwd_tc_kernel");
    return 0;
  }
  auto* fn = attn_bwd_simt_kernel<T16>;
  const size_t smem = (size_t)S * kAttnDH * 2 * 4 + (size_t)S * 4 * 4;
  TRY(set_smem_attr(fn, smem));
  fn<<<dim3(nH, B), 128, smem, st>>>(reinterpret_cast<const T16*>(qkv), reinterpret_cast<const T16*>(dctx), ext_mask, S,
                                     H, 0.125f, reinterpret_cast<T16*>(dqkv), make_drop(dropout, p_a, site));
  CKL("attn_bwd_simt_kernel");
  return 0;
}

template <typename T16>
static int gemm_plain(cpt_handle* h, cudaStream_t st, int tag, const void* A, long long lda, const void* W,
                      long long ldw, int M, int N, int K, void* out, long long ldo, bool out_fp32, bool accumulate,
                      int trans = 0) {
  // out(+)= A[M,K] . W[N,K]^T, no bias.  accumulate -> fp32 reduce-add into `out`.  trans (GemmParams::trans): 3 = A is
  // [K,M] and W is [K,N]; 2 = W is [K,N]
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.out = out; p.ldo = ldo; p.bias = nullptr;
  p.tma_reduce = accumulate ? 1 : 0;
  p.trans = trans;
  if (accumulate) {
    // accumulate-into-output products can cut K: weight gradients (few output tiles, K = all rows) until the work items
    // cover the SMs about twice; data gradients of small batches (fewer tiles than SMs) until they cover them once
    const int bn = N >= 2048 ? 256 : 192;
    const int tiles = ((M + kGemmBM - 1) / kGemmBM) * ((N + bn - 1) / bn);
    if (trans == 3) p.ksplit = std::max(1, (2 * h->num_sms + tiles - 1) / tiles);
    else if (tiles < h->num_sms) p.ksplit = std::max(1, h->num_sms / tiles);
  }
  return gemm<T16>(h, st, tag, A, lda, W, ldw, p, EPI_BIAS, out_fp32);
}

// dX[rows,Kin] (+)= dY[rows,Nout] . W[Nout,Kin]  (W read in place through an MN-major descriptor)
template <typename T16>
static int dgrad(cpt_handle* h, cudaStream_t st, const void* dY, long long ldy, const void* W, long long ldw, int rows,
                 int Nout, int Kin, void* dX, long long ldx, bool out_fp32, bool accumulate) {
  return gemm_plain<T16>(h, st, CPT_K_GEMM_DGRAD, dY, ldy, W, ldw, rows, Kin, Nout, dX, ldx, out_fp32, accumulate, 2);
}

// dW[Nout,Kin] += dY[rows,Nout]^T . X[rows,Kin]  (both operands read in place through MN-major descriptors)
template <typename T16>
static int wgrad(cpt_handle* h, cudaStream_t st, const void* dY, long long ldy, const void* X, long long ldx, int rows,
                 int Nout, int Kin, float* dW, long long ldw, bool accumulate = true) {
  return gemm_plain<T16>(h, st, CPT_K_GEMM_WGRAD, dY, ldy, X, ldx, Nout, Kin, rows, dW, ldw, true, accumulate, 3);
}

static int small_matmul(cpt_handle* h, cudaStream_t st, const float* A, long long sa0, long long sa1, const float* B,
                        long long sb0, long long sb1, int M, int N, int K, float* Cm, long long sc0, bool accumulate,
                        const float* tanh_out = nullptr) {
  ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
  const long long total = (long long)M * N;
  small_matmul_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A, sa0, sa1, B, sb0, sb1, M, N, K, Cm, sc0,
                                                                        accumulate ? 1 : 0, tanh_out);
  CKL("small_matmul_kernel");
  return 0;
}

template <typename T16>
static int train_forward_impl(cpt_handle* h, int head, cudaStream_t st, const int64_t* ids, const int64_t* seg,
                              const int64_t* mask, const int64_t* pos_ids, const float* img, int B, int T, int R,
                              const int64_t* rows, const int64_t* targets, int n, const cpt_dropout* dropout,
                              void* tape_ptr, size_t tape_bytes, float* loss) {
  const cpt_config& c = h->cfg;
  const int H = c.hidden_size, I = c.intermediate_size, L = c.num_hidden_layers, S = T + R, M = B * S, V = c.vocab_size;
  if (!h->train) return fail("cpt_train_forward_mlm: call cpt_train_enable(h, 1) before cpt_set_weights");
  if (!h->has_weights) return fail("cpt_train_forward called before cpt_set_weights");
  if (head == CPT_HEAD_MLM && !h->has_mlm) return fail("cpt_train_forward_mlm needs the cls.predictions.* weights");
  if (head == CPT_HEAD_NSP && (!h->has_pooler || !h->has_nsp))
    return fail("cpt_train_forward_nsp needs the bert.pooler and cls.seq_relationship weights");
  if (B <= 0 || T <= 0 || R < 0 || n <= 0) return fail("bad shape B=%d T=%d R=%d n_rows=%d", B, T, R, n);
  if (T > c.max_position_embeddings) return fail("T=%d exceeds max_position_embeddings=%d", T, c.max_position_embeddings);
  if (R > 0 && (!img || !h->w_img)) return fail("img_feats given but no img_embedding weights (or NULL img_feats)");
  if (S > 256) return fail("sequence length T+R=%d exceeds the 256 this build's attention kernel supports", S);
  if (!h->tma_store || !h->reduce_resid) return fail("the training path needs the TMA reduce-add stores (unset CPT_B200_TMA_STORE / CPT_B200_REDUCE_RESID)");
  if (!ids || !rows || !targets || !loss) return fail("input_ids, rows, targets and loss must be non-NULL");
  Tape t = carve_tape(h, B, T, R, n, (char*)(((uintptr_t)tape_ptr + 255) & ~uintptr_t(255)));
  if (!tape_ptr || tape_bytes < t.total) return fail("tape too small: need %zu bytes, got %zu", t.total, tape_bytes);
  const size_t act = (size_t)M * H * 4;
  const float p_h = dropout ? dropout->p_hidden : 0.f, p_a = dropout ? dropout->p_attn : 0.f;
  if (p_h < 0.f || p_h >= 1.f || p_a < 0.f || p_a >= 1.f) return fail("dropout probabilities must be in [0, 1)");

  if (mask) {
    ProfScope ps(h, st, CPT_K_EXTMASK);
    CK(launch_k(ext_mask_kernel, dim3((M + 255) / 256), dim3(256), 0, st, 1, (const long long*)mask, M, t.ext_mask));
  } else {
    CK(cudaMemsetAsync(t.ext_mask, 0, (size_t)M * 4, st));
  }
  char* h16_0 = L > 0 ? t.layers[0].h16 : t.dx16;
  // Without hidden-state dropout the fp32 residual stream is written straight into the tape slot the next dense layer
  // accumulates into (x1 / x2 hold "residual + dense output"), so no copy of the residual is ever made; with dropout
  // the residual stays in a scratch buffer and one elementwise pass forms residual + dropout(dense output).
  const bool in_place = !(p_h > 0.f) && L > 0;
  float* const emb32 = in_place ? t.layers[0].x1 : t.h32;
  {
    ProfScope ps(h, st, CPT_K_EMBED);
#define CPT_EMB_CASE(NV_)                                                                                            \
  case NV_:                                                                                                          \
    CK(launch_k(embed_text_ln_kernel<T16, NV_>, dim3((B * T + 7) / 8), dim3(256), 0, st, 1, (const long long*)ids,    \
                (const long long*)seg, (const long long*)pos_ids, h->word, h->pos, h->type, (const float*)h->emb_g,  \
                (const float*)h->emb_b, c.layer_norm_eps, B, T, S, H, c.vocab_size, c.max_position_embeddings,       \
                c.type_vocab_size, emb32, reinterpret_cast<T16*>(h16_0), h->err_flag));                              \
    break;
    switch (H / 128) {
      CPT_EMB_CASE(1) CPT_EMB_CASE(2) CPT_EMB_CASE(3) CPT_EMB_CASE(4) CPT_EMB_CASE(5) CPT_EMB_CASE(6) CPT_EMB_CASE(7)
      CPT_EMB_CASE(8)
      default: return fail("unsupported hidden size %d", H);
    }
#undef CPT_EMB_CASE
  }
  if (R > 0) {
    const int F = c.img_feature_dim, Mi = B * R;
    const int grid = (Mi + 7) / 8 < 8 * h->num_sms ? (Mi + 7) / 8 : 8 * h->num_sms;
    {
      ProfScope ps(h, st, CPT_K_CAST);
      CK(launch_k(cast_pad_kernel<T16>, dim3(grid), dim3(256), 0, st, 1, img, Mi, F, h->Fp,
                  reinterpret_cast<T16*>(t.img16)));
    }
    GemmParams p{};
    p.M = Mi; p.N = H; p.K = F; p.out = t.imgpre32; p.ldo = H; p.bias = h->b_img;
    TRY(gemm<T16>(h, st, CPT_K_GEMM_IMG, t.img16, h->Fp, h->w_img, h->Fp, p, EPI_BIAS, true));
    TRY(layernorm<T16>(h, st, t.imgpre32, H, Mi, H, h->img_g, h->img_b, c.img_layer_norm_eps,
                       c.use_img_layernorm != 0, emb32, h16_0, R, S, T));
  }
  if (p_h > 0.f) {  // dropout on the embedding outputs (BertEmbeddings.dropout; modeling_bert.py:266 for regions)
    ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
    dropout_rows_kernel<T16><<<ew_grid(h, (long long)B * T * H), 256, 0, st>>>(
        t.h32, reinterpret_cast<T16*>(h16_0), B * T, H, T, S, 0, make_drop(dropout, p_h, SITE_EMB_TEXT));
    CKL("dropout_rows_kernel");
    if (R > 0) {
      dropout_rows_kernel<T16><<<ew_grid(h, (long long)B * R * H), 256, 0, st>>>(
          t.h32, reinterpret_cast<T16*>(h16_0), B * R, H, R, S, T, make_drop(dropout, p_h, SITE_EMB_IMG));
      CKL("dropout_rows_kernel");
    }
  }
  // x = resid + dropout(A W^T + b): without dropout the GEMM's TMA stores add into a copy of the residual; with it the
  // dense output goes to scratch first and one elementwise pass masks, scales and adds
  auto dense_residual = [&](int tag, const void* A, int K, const void* W, const float* bias, const float* resid,
                            float* x, unsigned site) -> int {
    GemmParams p{};
    p.M = M; p.N = H; p.K = K; p.ldo = H; p.bias = bias;
    if (p_h > 0.f) {
      p.out = t.dx32;
      TRY(gemm<T16>(h, st, tag, A, K, W, K, p, EPI_BIAS, true));
      ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
      dropout_add_kernel<<<ew_grid(h, (long long)M * H), 256, 0, st>>>(resid, t.dx32, (long long)M * H, x,
                                                                      make_drop(dropout, p_h, site));
      CKL("dropout_add_kernel");
      return 0;
    }
    if (resid != x) CK(cudaMemcpyAsync(x, resid, act, cudaMemcpyDeviceToDevice, st));
    p.out = x; p.tma_reduce = 1;
    return gemm<T16>(h, st, tag, A, K, W, K, p, EPI_BIAS, true);
  };
  for (int l = 0; l < L; ++l) {
    const LayerDev& d = h->layers[l];
    TapeLayer& tl = t.layers[l];
    {
      GemmParams p{};
      p.M = M; p.N = 3 * H; p.K = H; p.out = tl.qkv16; p.ldo = 3 * H; p.bias = d.b_qkv;
      TRY(gemm<T16>(h, st, CPT_K_GEMM_QKV, tl.h16, H, d.w_qkv, H, p, EPI_BIAS, false));
    }
    if (p_a > 0.f) {
      // probability dropout: the single-tile tcgen05 kernel masks P on its way to shared memory (CPT_B200_ATTN_BWD=simt
      // selects the CUDA-core pair of kernels instead)
      ProfScope ps(h, st, CPT_K_ATTN);
      const Drop dr = make_drop(dropout, p_a, l * 4 + SITE_ATTN);
      if (h->attn_bwd_simt) {
        auto* fn = attn_fwd_drop_kernel<T16>;
        const size_t smem = (size_t)S * kAttnDH * 2 * 2 + (size_t)S * 4;
        TRY(set_smem_attr(fn, smem));
        fn<<<dim3(c.num_attention_heads, B), 128, smem, st>>>(reinterpret_cast<const T16*>(tl.qkv16), t.ext_mask, S, H,
                                                              0.125f, reinterpret_cast<T16*>(tl.ctx16), dr);
        CKL("attn_fwd_drop_kernel");
      } else {
        CUtensorMap tq;
        TRY(make_tmap(&tq, tl.qkv16, Cvt<T16>::kFmt, (unsigned long long)B * S, 3ull * H, 3ull * H, 64));
        AttnParams ap{B, S, H, c.num_attention_heads, t.ext_mask, tl.ctx16, 0.125f, nullptr, dr};
        auto* fn = attn_tc_kernel<T16>;
        const size_t smem = attn_smem_bytes(S);
        TRY(set_smem_attr(fn, smem));
        CK(launch_k(fn, dim3(c.num_attention_heads, (S + 127) / 128, B), dim3(kAttnThreads), smem, st, 1, tq, ap));
      }
    } else {
      TRY(attention<T16>(h, st, tl.qkv16, t.ext_mask, B, S, tl.ctx16, h->attn_impl));
    }
    TRY(dense_residual(CPT_K_GEMM_AO, tl.ctx16, H, d.w_ao, d.b_ao, in_place ? tl.x1 : t.h32, tl.x1, l * 4 + SITE_AO));
    float* const a32 = in_place ? tl.x2 : t.a32;
    TRY(layernorm<T16>(h, st, tl.x1, H, M, H, d.ao_g, d.ao_b, c.layer_norm_eps, true, a32, tl.a16));
    {
      GemmParams p{};
      p.M = M; p.N = I; p.K = H; p.out = tl.preup16; p.ldo = I; p.bias = d.b_i;
      TRY(gemm<T16>(h, st, CPT_K_GEMM_UP, tl.a16, H, d.w_i, H, p, EPI_BIAS, false));
      ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
      const long long ne = (long long)M * I;
      gelu_fwd_kernel<T16><<<ew_grid(h, ne / 8), 256, 0, st>>>(reinterpret_cast<const T16*>(tl.preup16), ne,
                                                          reinterpret_cast<T16*>(tl.inter16));
      CKL("gelu_fwd_kernel");
    }
    TRY(dense_residual(CPT_K_GEMM_DOWN, tl.inter16, I, d.w_o, d.b_o, a32, tl.x2, l * 4 + SITE_DOWN));
    TRY(layernorm<T16>(h, st, tl.x2, H, M, H, d.o_g, d.o_b, c.layer_norm_eps, true,
                       (l == L - 1) ? t.seq32 : (in_place ? t.layers[l + 1].x1 : t.h32),
                       (l == L - 1) ? nullptr : t.layers[l + 1].h16));
  }
  if (L == 0) CK(cudaMemcpyAsync(t.seq32, t.h32, act, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemsetAsync(loss, 0, 4, st));
  if (head == CPT_HEAD_NSP) {
    // BertPooler on the [CLS] rows of the labelled samples, then cls.seq_relationship; fp32 throughout
    const int Cn = c.num_contrast_classes;
    {
      ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
      gather_rows_kernel<T16><<<n, 256, 0, st>>>(t.seq32, (const long long*)rows, n, H, (T16*)nullptr, t.nx32);
      CKL("gather_rows_kernel");
    }
    TRY(head_matvec(h, st, t.nx32, H, 1, nullptr, nullptr, nullptr, 0.f, h->pool_w, H, h->pool_b, nullptr, H, n, H, H,
                    ACT_TANH, t.npool, H));
    TRY(head_matvec(h, st, t.npool, H, 1, nullptr, nullptr, nullptr, 0.f, h->nsp_w, H, h->nsp_b, nullptr, Cn, n, H,
                    Cn, ACT_NONE, t.nlog, Cn));
    ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
    ce_fwd_kernel<<<n, 256, 0, st>>>(t.nlog, Cn, n, Cn, (const long long*)targets, t.lse, loss);
    CKL("ce_fwd_kernel");
    return 0;
  }
  // head at the labelled rows: cls.predictions.transform (dense, GELU, LayerNorm) and the tied decoder
  {
    ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
    gather_rows_kernel<T16><<<n, 256, 0, st>>>(t.seq32, (const long long*)rows, n, H, reinterpret_cast<T16*>(t.hx16),
                                               (float*)nullptr);
    CKL("gather_rows_kernel");
  }
  {
    GemmParams p{};
    p.M = n; p.N = H; p.K = H; p.out = t.htd32; p.ldo = H; p.bias = h->mlm_b;
    TRY(gemm<T16>(h, st, CPT_K_GEMM_HEAD, t.hx16, H, h->mlm_w16, H, p, EPI_BIAS, true));
    {
      ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
      gelu_fwd32_kernel<<<ew_grid(h, (long long)n * H), 256, 0, st>>>(t.htd32, (long long)n * H, t.htg32);
      CKL("gelu_fwd32_kernel");
    }
    TRY(layernorm<T16>(h, st, t.htg32, H, n, H, h->mlm_g, h->mlm_beta, c.layer_norm_eps, true, nullptr, t.ht16));
    GemmParams q{};
    q.M = n; q.N = V; q.K = H; q.out = t.logits; q.ldo = t.ldl; q.bias = h->mlm_bias;
    TRY(gemm<T16>(h, st, CPT_K_GEMM_HEAD, t.ht16, H, h->word16, H, q, EPI_BIAS, true));
  }
  {
    ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
    ce_fwd_kernel<<<n, 256, 0, st>>>(t.logits, t.ldl, n, V, (const long long*)targets, t.lse, loss);
    CKL("ce_fwd_kernel");
  }
  return 0;
}

template <typename T16>
static int train_backward_impl(cpt_handle* h, int head, cudaStream_t st, const int64_t* ids, const int64_t* seg,
                               const int64_t* pos_ids, int B, int T, int R, const int64_t* rows,
                               const int64_t* targets, int n, const cpt_dropout* dropout, const float* grad_loss,
                               void* tape_ptr, size_t tape_bytes, const cpt_grads* g) {
  const cpt_config& c = h->cfg;
  const int H = c.hidden_size, I = c.intermediate_size, L = c.num_hidden_layers, S = T + R, M = B * S, V = c.vocab_size;
  if (!h->train || !h->has_weights) return fail("cpt_train_backward: handle is not set up for training");
  if (B <= 0 || T <= 0 || R < 0 || n <= 0 || S > 256) return fail("bad shape B=%d T=%d R=%d n_rows=%d", B, T, R, n);
  if (!g || !g->word_emb || !g->pos_emb || !g->type_emb || !g->emb_ln_g || !g->emb_ln_b || (L > 0 && !g->layers))
    return fail("cpt_train_backward: NULL gradient tensor");
  if (head == CPT_HEAD_MLM &&
      (!h->has_mlm || !g->mlm_dense_w || !g->mlm_dense_b || !g->mlm_ln_g || !g->mlm_ln_b || !g->mlm_bias))
    return fail("cpt_train_backward_mlm: NULL cls.predictions.* gradient tensor");
  if (head == CPT_HEAD_NSP && (!h->has_pooler || !h->has_nsp || !g->pooler_w || !g->pooler_b || !g->nsp_w || !g->nsp_b))
    return fail("cpt_train_backward_nsp: NULL pooler / seq_relationship gradient tensor");
  if (R > 0 && (!g->img_w || !g->img_b || (c.use_img_layernorm && (!g->img_ln_g || !g->img_ln_b))))
    return fail("cpt_train_backward_mlm: NULL img_* gradient tensor");
  if (!ids || !rows || !targets || !grad_loss) return fail("NULL argument");
  Tape t = carve_tape(h, B, T, R, n, (char*)(((uintptr_t)tape_ptr + 255) & ~uintptr_t(255)));
  if (!tape_ptr || tape_bytes < t.total) return fail("tape too small: need %zu bytes, got %zu", t.total, tape_bytes);
  const int Vp = t.Vp;
  const float p_h = dropout ? dropout->p_hidden : 0.f, p_a = dropout ? dropout->p_attn : 0.f;
  const Drop no_drop{0u, 0u, 0u, 0u, 1.f};

  CK(cudaMemsetAsync(t.dH, 0, (size_t)M * H * 4, st));
  if (head == CPT_HEAD_NSP) {
    const int Cn = c.num_contrast_classes;
    {
      ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
      ce_bwd32_kernel<<<(n * Cn + 127) / 128, 128, 0, st>>>(t.nlog, n, Cn, (const long long*)targets, t.lse, grad_loss,
                                                           t.ndlog);
      CKL("ce_bwd32_kernel");
    }
    TRY(colsum<float>(h, st, t.ndlog, n, Cn, Cn, g->nsp_b));
    // dW_nsp[C,H] += dlogits^T pooled ; dpre[n,H] = (dlogits W_nsp) * (1 - pooled^2)
    TRY(small_matmul(h, st, t.ndlog, 1, Cn, t.npool, H, 1, Cn, H, n, g->nsp_w, H, true));
    TRY(small_matmul(h, st, t.ndlog, Cn, 1, h->nsp_w, H, 1, n, H, Cn, t.ndpre, H, false, t.npool));
    TRY(colsum<float>(h, st, t.ndpre, n, H, H, g->pooler_b));
    // dW_pool[H,H] += dpre^T x ; dx[n,H] = dpre W_pool
    TRY(small_matmul(h, st, t.ndpre, 1, H, t.nx32, H, 1, H, H, n, g->pooler_w, H, true));
    TRY(small_matmul(h, st, t.ndpre, H, 1, h->pool_w, H, 1, n, H, H, t.ndx, H, false));
    ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
    scatter_rows_add_kernel<<<n, 256, 0, st>>>(t.ndx, (const long long*)rows, n, H, t.dH);
    CKL("scatter_rows_add_kernel");
  } else {
  // ---- MLM head
  {
    ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
    ce_bwd_kernel<T16><<<dim3(n, std::max(1, std::min(32, 2 * h->num_sms / n))), 256, 0, st>>>(t.logits, t.ldl, n, V, (const long long*)targets, t.lse, grad_loss,
                                          reinterpret_cast<T16*>(t.dlog16), Vp);
    CKL("ce_bwd_kernel");
  }
  TRY(colsum<T16>(h, st, reinterpret_cast<const T16*>(t.dlog16), n, V, Vp, g->mlm_bias));
  TRY(wgrad<T16>(h, st, t.dlog16, Vp, t.ht16, H, n, V, H, g->word_emb, H));
  TRY(dgrad<T16>(h, st, t.dlog16, Vp, h->word16, H, n, V, H, t.hd32a, H, true, false));
  TRY(ln_bwd<T16>(h, st, t.hd32a, t.htg32, n, H, h->mlm_g, c.layer_norm_eps, true, t.hd32b, nullptr, g->mlm_ln_g,
                  g->mlm_ln_b, nullptr));
  {
    ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
    const long long ne = (long long)n * H;
    gelu_bwd32_kernel<<<ew_grid(h, ne), 256, 0, st>>>(t.hd32b, t.htd32, ne, t.hd32a);
    CKL("gelu_bwd32_kernel");
  }
  {
    ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
    const long long ne = (long long)n * H;
    cast32to16_kernel<<<ew_grid(h, ne), 256, 0, st>>>(
        t.hd32a, ne, c.dtype == 0 ? reinterpret_cast<__half*>(t.hd16) : nullptr,
        c.dtype == 0 ? nullptr : reinterpret_cast<__nv_bfloat16*>(t.hd16));
    CKL("cast32to16_kernel");
  }
  TRY(colsum<float>(h, st, t.hd32a, n, H, H, g->mlm_dense_b));
  TRY(wgrad<T16>(h, st, t.hd16, H, t.hx16, H, n, H, H, g->mlm_dense_w, H));
  TRY(dgrad<T16>(h, st, t.hd16, H, h->mlm_w16, H, n, H, H, t.hdx32, H, true, false));
  {
    ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
    scatter_rows_add_kernel<<<n, 256, 0, st>>>(t.hdx32, (const long long*)rows, n, H, t.dH);
    CKL("scatter_rows_add_kernel");
  }
  }

  if (h->progress_cb) h->progress_cb(h->progress_user, 0);
  // ---- encoder layers, last to first.  t.dH = gradient of the layer's output
  for (int l = L - 1; l >= 0; --l) {
    const LayerDev& d = h->layers[l];
    const TapeLayer& tl = t.layers[l];
    const cpt_layer_grads& gl = g->layers[l];
    // output.LayerNorm
    // output.dense: x2 = a + dropout(inter W2^T + b2); t.dx16 carries the masked gradient of the dense output, whose
    // column sums (the bias gradient) the same kernel accumulates
    TRY(ln_bwd<T16>(h, st, t.dH, tl.x2, M, H, d.o_g, c.layer_norm_eps, true, t.dx32, t.dx16, gl.o_ln_g, gl.o_ln_b,
                    gl.o_b, 0, 0, 0, no_drop, make_drop(dropout, p_h, l * 4 + SITE_DOWN)));
    TRY(wgrad<T16>(h, st, t.dx16, H, tl.inter16, I, M, H, I, gl.o_w, I));
    TRY(dgrad<T16>(h, st, t.dx16, H, d.w_o, I, M, H, I, t.big16, I, false, false));
    {  // GELU backward + the bias gradient of intermediate.dense
      ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
      const int gx = (I / 8 + 127) / 128;
      const int rpc = std::max(4, std::min(64, (M * gx + 4 * h->num_sms - 1) / (4 * h->num_sms)));  // ~4 CTAs per SM
      gelu_bwd_kernel<T16><<<dim3(gx, (M + rpc - 1) / rpc), 256, 0, st>>>(
          reinterpret_cast<const T16*>(t.big16), reinterpret_cast<const T16*>(tl.preup16), M, I,
          reinterpret_cast<T16*>(t.big16b), gl.i_b, rpc);
      CKL("gelu_bwd_kernel");
    }
    // intermediate.dense
    TRY(wgrad<T16>(h, st, t.big16b, I, tl.a16, H, M, I, H, gl.i_w, H));
    TRY(dgrad<T16>(h, st, t.big16b, I, d.w_i, H, M, I, H, t.dx32, H, true, true));  // += residual branch
    // attention.output.LayerNorm  (dx1 -> t.dH)
    TRY(ln_bwd<T16>(h, st, t.dx32, tl.x1, M, H, d.ao_g, c.layer_norm_eps, true, t.dH, t.dx16, gl.ao_ln_g,
                    gl.ao_ln_b, gl.ao_b, 0, 0, 0, no_drop, make_drop(dropout, p_h, l * 4 + SITE_AO)));
    // attention.output.dense
    TRY(wgrad<T16>(h, st, t.dx16, H, tl.ctx16, H, M, H, H, gl.ao_w, H));
    TRY(dgrad<T16>(h, st, t.dx16, H, d.w_ao, H, M, H, H, t.dctx16, H, false, false));
    TRY(attention_backward<T16>(h, st, tl.qkv16, t.dctx16, t.ext_mask, B, S, t.big16, dropout, p_a,
                                (unsigned)(l * 4 + SITE_ATTN), -1));
    // query / key / value
    float* qkv_b[3] = {gl.q_b, gl.k_b, gl.v_b};
    float* qkv_w[3] = {gl.q_w, gl.k_w, gl.v_w};
    TRY(colsum<T16>(h, st, reinterpret_cast<const T16*>(t.big16), M, 3 * H, 3 * H, qkv_b[0], qkv_b[1], qkv_b[2], H));
    if (qkv_w[1] == qkv_w[0] + (size_t)H * H && qkv_w[2] == qkv_w[1] + (size_t)H * H) {
      // the three gradients are adjacent in the caller's slab: one [3H, H] product instead of three
      TRY(wgrad<T16>(h, st, t.big16, 3 * H, tl.h16, H, M, 3 * H, H, qkv_w[0], H));
    } else {
      for (int j = 0; j < 3; ++j)
        TRY(wgrad<T16>(h, st, t.big16 + (size_t)j * H * 2, 3 * H, tl.h16, H, M, H, H, qkv_w[j], H));
    }
    TRY(dgrad<T16>(h, st, t.big16, 3 * H, d.w_qkv, H, M, 3 * H, H, t.dH, H, true, true));  // += residual
    if (h->progress_cb) h->progress_cb(h->progress_user, L - l);
  }

  // ---- embeddings
  {
    ProfScope ps(h, st, CPT_K_EMBED_BWD);
#define CPT_EMBB_CASE(NV_)                                                                                           \
  case NV_:                                                                                                          \
    embed_bwd_kernel<NV_><<<std::min((B * T + 7) / 8, 2 * h->num_sms), 256, 0, st>>>(                                                          \
        (const long long*)ids, (const long long*)seg, (const long long*)pos_ids, h->word, h->pos, h->type, h->emb_g, \
        c.layer_norm_eps, t.dH, B, T, S, H, c.vocab_size, c.max_position_embeddings, c.type_vocab_size, g->word_emb, \
        g->pos_emb, g->type_emb, g->emb_ln_g, g->emb_ln_b, make_drop(dropout, p_h, SITE_EMB_TEXT));                  \
    break;
    switch (H / 128) {
      CPT_EMBB_CASE(1) CPT_EMBB_CASE(2) CPT_EMBB_CASE(3) CPT_EMBB_CASE(4) CPT_EMBB_CASE(5) CPT_EMBB_CASE(6)
      CPT_EMBB_CASE(7) CPT_EMBB_CASE(8)
      default: return fail("unsupported hidden size %d", H);
    }
#undef CPT_EMBB_CASE
    CKL("embed_bwd_kernel");
  }
  if (R > 0) {
    const int F = c.img_feature_dim, Mi = B * R;
    TRY(ln_bwd<T16>(h, st, t.dH, t.imgpre32, Mi, H, h->img_g, c.img_layer_norm_eps, c.use_img_layernorm != 0,
                    t.dimg32, t.dimg16, g->img_ln_g, g->img_ln_b, g->img_b, R, S, T,
                    make_drop(dropout, p_h, SITE_EMB_IMG)));
    TRY(wgrad<T16>(h, st, t.dimg16, H, t.img16, h->Fp, Mi, H, h->Fp, t.dwimg, h->Fp, false));
    ProfScope ps(h, st, CPT_K_TRAIN_ROWWISE);
    add_rows_kernel<<<ew_grid(h, (long long)H * F), 256, 0, st>>>(t.dwimg, h->Fp, H, F, g->img_w, F);
    CKL("add_rows_kernel");
  }
  if (h->progress_cb) h->progress_cb(h->progress_user, L + 1);
  return 0;
}
