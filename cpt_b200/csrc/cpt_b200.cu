// C ABI + host orchestration of the CPT cross-modal BERT path on one B200 (see include/cpt_b200.h).
// Kernel sequence of one encoder forward (K-numbers: SURVEY.md 2.4b):
//   ext_mask (K4) | embed_text_ln (K1) | cast_pad + GEMM(bias) + LN -> region rows (K2,K3)
//   per layer: GEMM QKV(bias) (K5) | attention (K6-K9) | GEMM out(bias+resid) + LN (K10)
//              | GEMM up(bias+GELU) (K11) | GEMM down(bias+resid) + LN (K12)
//   heads: pooler/NSP mat-vec (K13,K17) | MLM transform + decoder at the [MASK] rows (K14,K15)
#include <cuda_runtime.h>
#include <cuda.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cpt_b200.h"
#include "attention_sm100.cuh"
#include "attention_bwd_sm100.cuh"
#include "gemm_sm100.cuh"
#include "chain_sm100.cuh"
#include "chain2_sm100.cuh"
#include "optim.cuh"
#include "rowwise.cuh"
#include "scoring.cuh"
#include "assemble.cuh"
#include "train.cuh"

using namespace cptk;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKL(what)                                                                                      \
  do {                                                                                                 \
    cudaError_t e_ = cudaGetLastError();                                                               \
    if (e_ != cudaSuccess) return fail("launch of %s failed: %s (%s:%d)", what, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define TRY(expr)              \
  do {                         \
    int rc_ = (expr);          \
    if (rc_ != 0) return rc_;  \
  } while (0)

// ------------------------------------------------------------------------------------------------ TMA maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;
static int get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  return g_encode ? 0 : fail("cuTensorMapEncodeTiled is not available from this driver");
}
// 2-D row-major 16-bit tensor [rows, cols] with pitch ld (elements); box = 64 cols x box_rows, 128B swizzle,
// out-of-bounds elements read as zero.
// dtype: 0 = fp16, 1 = bf16, 2 = fp32
static int make_tmap_ex(CUtensorMap* m, const void* ptr, int dtype, unsigned long long rows, unsigned long long cols,
                        unsigned long long ld, unsigned box_cols, unsigned box_rows, CUtensorMapSwizzle swz) {
  TRY(get_encode());
  const unsigned esz = dtype == 2 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * esz) & 15))
    return fail("TMA tensor must be 16-byte aligned with a 16-byte-multiple pitch (ptr=%p ld=%llu)", ptr, ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                            : (dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = g_encode(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}
// 16-bit [B][S][H] view of a [B*S, H] matrix, box = 64 cols x box_rows x 1, 128B swizzle: a store whose rows run past
// S is clipped at the sample boundary
static int make_tmap_bsh(CUtensorMap* m, const void* ptr, int dtype, unsigned long long B, unsigned long long S,
                         unsigned long long H, unsigned box_rows) {
  TRY(get_encode());
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((H * 2) & 15)) return fail("TMA tensor must be 16-byte aligned");
  cuuint64_t dims[3] = {H, S, B};
  cuuint64_t strides[2] = {H * 2, S * H * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                        const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}
static int make_tmap(CUtensorMap* m, const void* ptr, int dtype, unsigned long long rows, unsigned long long cols,
                     unsigned long long ld, unsigned box_rows) {
  return make_tmap_ex(m, ptr, dtype, rows, cols, ld, 64, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

// ------------------------------------------------------------------------------------------------ handle
struct LayerDev {
  void *w_qkv = nullptr, *w_ao = nullptr, *w_i = nullptr, *w_o = nullptr;  // 16-bit [3H,H] [H,H] [I,H] [H,I]
  float *b_qkv = nullptr, *b_ao = nullptr, *ao_g = nullptr, *ao_b = nullptr, *b_i = nullptr, *b_o = nullptr,
        *o_g = nullptr, *o_b = nullptr;
  // LayerNorm-folded copies (DESIGN.md "LayerNorm folding"): W .* gamma of the LayerNorm that feeds the GEMM
  void *w_qkv_f = nullptr, *w_i_f = nullptr;
  float *g_qkv = nullptr, *c_qkv = nullptr, *g_i = nullptr, *c_i = nullptr;
};
struct cpt_handle {
  cpt_config cfg;
  int device = 0, num_sms = 148;
  int Fp = 0;  // padded region-feature pitch
  bool has_weights = false, has_mlm = false, has_nsp = false, has_pooler = false;
  std::vector<void*> owned;
  const float *word = nullptr, *pos = nullptr, *type = nullptr;
  float *emb_g = nullptr, *emb_b = nullptr;
  void* w_img = nullptr;
  float *b_img = nullptr, *img_g = nullptr, *img_b = nullptr;
  float *pool_w = nullptr, *pool_b = nullptr;
  float *mlm_w = nullptr, *mlm_b = nullptr, *mlm_g = nullptr, *mlm_beta = nullptr, *mlm_bias = nullptr;
  void *mlm_w16 = nullptr, *word16 = nullptr;
  cpt_progress_fn progress_cb = nullptr;           // cpt_train_set_progress_callback
  void* progress_user = nullptr;
  int ln_ctas_per_sm = 2;                          // CPT_B200_LN_CTAS: persistent LayerNorm CTAs per SM (A/B experiment)
  int small_m_tiles = 1;                           // CPT_B200_SMALL_M=0: no narrower tiles for small row counts
  int down_ksplit = 1;                             // CPT_B200_DOWN_KSPLIT: split-K pieces of the FFN-down GEMM (A/B experiment)
  int attn_bwd_simt = 0;                           // CPT_B200_ATTN_BWD=simt: CUDA-core attention backward everywhere
  int train = 0;                                   // cpt_train_enable: cpt_set_weights refreshes in place, no LN-folded copies
  // the copy list of the last cpt_set_weights (host mirror + device tables) — replayed as one launch
  std::vector<CopyEntry> copy_list;
  std::vector<CopyEntry> copy_list_dev_mirror;
  void *copy_entries_dev = nullptr, *copy_chunks_dev = nullptr;
  int copy_n_chunks = 0;
  unsigned weights_sig = 0;                        // which optional tensors the current allocations cover
  float *nsp_w = nullptr, *nsp_b = nullptr;
  std::vector<LayerDev> layers;
  // dataflow chain kernel (chain_sm100.cuh): one launch per layer for everything between two attention kernels
  int chain = 1;                 // CPT_B200_CHAIN=0: the round-1 launch sequence (one kernel per GEMM / LayerNorm)
  int chain_min_rows = 1024;     // below this many rows the narrow-tile GEMMs of the unfused path spread better
  int chain_down_ksplit = 1;     // CPT_B200_CHAIN_KSPLIT: K pieces of the FFN-down tiles (unfused LayerNorm only)
  int chain_fuse_ln = 2;         // CPT_B200_CHAIN_FUSE_LN: 0 LayerNorm as row tasks between the GEMM stages, 1 finished
                                 // in the dense epilogues (row statistics exchanged through L2), 2 deferred to the consumers
  int chain_groups = 1;          // CPT_B200_CHAIN_GROUPS: row groups software-pipelined across the stages
  int attn_early = 1;            // CPT_B200_ATTN_EARLY=0: attention waits for the whole chain launch before it (plain PDL)
  const unsigned* pub_ready = nullptr;   // set by a chain launch whose last stage publishes its rows for the next launch
  unsigned pub_target = 0;
  int chain_early = 0;           // CPT_B200_CHAIN_EARLY=1: the chain launch too starts inside the previous launch's tail
                                 // (attention counts its context rows per tile).  Measured: no gain (1.818 vs 1.810 ms at
                                 // B=64; the counting costs attention what the overlap saves), so off by default.
  unsigned* ctx_counters = nullptr;      // set by the encoder loop: where the next attention launch may count its rows
  const unsigned* ctx_published = nullptr;  // set by that attention launch if it does; read by the chain launch after it
  int chain_lean = 1;            // CPT_B200_CHAIN_LEAN=0: run deferred-LayerNorm launches on the general kernel
  float2* chain_part = nullptr;  // cpt_chain_run (tests): scratch of the fused LayerNorm epilogues
  size_t chain_part_bytes = 0;
  struct ChainSched { int pairs = 0, pitch = 0; int* dev = nullptr; };
  std::map<std::vector<int>, ChainSched> chain_scheds;  // per chain shape: the task lists of the CTA pairs
  std::vector<void*> owned_chain;
  int chain_trace_on = 0;              // CPT_B200_CHAIN_TRACE=1: per-task event log of the last chain launch
  long long* chain_trace = nullptr;
  size_t chain_trace_bytes = 0;
  int chain_trace_pairs = 0, chain_trace_pitch = 0;
  unsigned* chain_counters = nullptr;  // cpt_chain_run (tests): readiness counters
  size_t chain_counters_bytes = 0;
  int* err_flag = nullptr;
  int attn_impl = 0;
  int fold_ln = 0;       // CPT_B200_FOLD_LN=1: LayerNorm folded into the neighbouring GEMM epilogues (slower, kept for study)
  int resid_in_ln = 1;   // residual added by the (streaming) LayerNorm kernel instead of the GEMM epilogue
  int reduce_resid = 1;  // ... or, better, by the GEMM's TMA store itself (cp.reduce .add at L2): CPT_B200_REDUCE_RESID=0 off
  int split = 1;         // CPT_B200_SPLIT=2: run the two halves of the batch as concurrent branches (tail filling)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int tma_store = 1;     // CPT_B200_TMA_STORE=0: LSU stores in the GEMM epilogue (A/B experiments)
  int delta16 = 0;       // CPT_B200_DELTA16=1: hand the dense+bias delta to it as 16 bits (2% faster, ~1.8x the logit error)
  struct { int bn, pair; } gemm_choice[16] = {};  // per kernel class, bn 0 = default (CPT_B200_GEMM overrides)
  // launch accounting / optional per-kernel-class CUDA-event timing (cpt_profile_*)
  long long launches = 0;
  long long* trace = nullptr;  // [num_sms][8] GEMM cycle counters when CPT_B200_TRACE=1
  bool profiling = false;
  struct ProfRec { int tag; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[CPT_K_COUNT] = {};
  long long prof_n[CPT_K_COUNT] = {};
};

// Brackets one kernel launch: counts it, and when profiling is on records CUDA events on the launch stream.
struct ProfScope {
  cpt_handle* h; cudaStream_t st; cudaEvent_t b = nullptr;
  ProfScope(cpt_handle* h_, cudaStream_t st_, int tag) : h(h_), st(st_) {
    h->launches++;
    h->prof_n[tag]++;
    if (!h->profiling) return;
    auto get = [&]() { cudaEvent_t e; if (!h->ev_pool.empty()) { e = h->ev_pool.back(); h->ev_pool.pop_back(); } else cudaEventCreate(&e); return e; };
    cudaEvent_t a = get(); b = get();
    cudaEventRecord(a, st);
    h->prof.push_back({tag, a, b});
  }
  ~ProfScope() { if (b) cudaEventRecord(b, st); }
};

static int dev_alloc(cpt_handle* h, void** p, size_t bytes) {
  CK(cudaMalloc(p, bytes ? bytes : 16));
  h->owned.push_back(*p);
  return 0;
}
static void free_owned(cpt_handle* h) {
  for (void* p : h->owned) cudaFree(p);
  h->owned.clear();
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int d) {
    cudaGetDevice(&prev);
    if (prev != d) cudaSetDevice(d);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ------------------------------------------------------------------------------------------------ launchers
// Launch on `st` with programmatic stream serialization (PDL) and, optionally, a thread-block cluster.
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*fn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[n].val.programmaticStreamSerializationAllowed = 1;
  ++n;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, fn, static_cast<KArgs>(args)...);
}

template <typename F>
static int set_smem_attr(F* fn, size_t bytes) {
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

struct GemmChoice { int bn, pair; };

template <int BN, bool PAIR, int EPI, typename OutT, typename T16>
static int launch_gemm_t(cpt_handle* h, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb,
                         const CUtensorMap& to, const GemmParams& p) {
  using Cfg = GemmCfg<BN, (int)sizeof(OutT), PAIR, (EPI == EPI_BIAS_RESID) ? 3 : 2>;
  auto* fn = gemm_kernel<BN, PAIR, EPI, OutT, T16>;
  static bool attr_set[64] = {};
  if (!attr_set[h->device & 63]) {
    TRY(set_smem_attr(fn, Cfg::kSmemBytes));
    attr_set[h->device & 63] = true;
  }
  constexpr int kCluster = PAIR ? 2 : 1;
  const int m_tiles = (p.M + kGemmBM - 1) / kGemmBM, n_tiles = (p.N + BN - 1) / BN;
  const int ctiles = ((m_tiles + kCluster - 1) / kCluster) * n_tiles * (p.ksplit > 1 ? p.ksplit : 1);
  int clusters = h->num_sms / kCluster;
  if (ctiles < clusters) clusters = ctiles;
  CK(launch_k(fn, dim3(clusters * kCluster), dim3(kGemmThreads), Cfg::kSmemBytes, st, kCluster, ta, tb, to, p));
  return 0;
}

template <int EPI, typename OutT, typename T16>
static int launch_gemm_bn(cpt_handle* h, cudaStream_t st, GemmChoice c, const CUtensorMap& ta, const CUtensorMap& tb,
                          const CUtensorMap& to, const GemmParams& p) {
#define CPT_BN_CASE(BN_)                                                                \
  case BN_:                                                                             \
    return c.pair ? launch_gemm_t<BN_, true, EPI, OutT, T16>(h, st, ta, tb, to, p)     \
                  : launch_gemm_t<BN_, false, EPI, OutT, T16>(h, st, ta, tb, to, p);
  switch (c.bn) {
    CPT_BN_CASE(64)
    CPT_BN_CASE(128)
    CPT_BN_CASE(192)
    CPT_BN_CASE(256)
  }
#undef CPT_BN_CASE
  return fail("unsupported GEMM block_n %d", c.bn);
}

// tile choice per GEMM class; cfg = block_n + 1000 * (1 + pair), 0 = default for the class
static GemmChoice pick_gemm(const cpt_handle* h, int tag, int M, int N, int K, int cfg, int ksplit = 1) {
  GemmChoice c{h->gemm_choice[tag].bn, h->gemm_choice[tag].pair};
  // few-shot / single-query batches: when the default tiling yields fewer work items than a quarter of the SMs,
  // narrower single-CTA tiles spread the same work over more SMs (measured: B=1 latency 0.87 -> 0.70 ms, B=4 training
  // step 4.27 -> 4.09 ms; at M = 1920 the default tiling is still the better one).  CPT_B200_SMALL_M=0 turns it off.
  if (c.bn == 0 && cfg == 0 && h->small_m_tiles && ksplit <= 1) {  // split-K already multiplies the work items
    const int bn0 = N >= 2048 ? 256 : 192;
    const int m_tiles = (M + kGemmBM - 1) / kGemmBM;
    if (m_tiles * ((N + bn0 - 1) / bn0) * 4 <= h->num_sms) {
      const int bn = m_tiles * ((N + 127) / 128) * 2 <= h->num_sms ? 64 : 128;
      return GemmChoice{bn, 0};
    }
  }
  // A/B-measured in situ on B200 at M = 7680 (scripts in tools/, results in profiles/): wide N -> 256-wide CTA-pair
  // tiles; N = 768 -> 192-wide tiles (4 per row, better wave balance), paired only when K is long
  if (c.bn == 0) c = N >= 2048 ? GemmChoice{256, 1} : GemmChoice{192, K >= 2048 ? 1 : 0};
  if (cfg > 0) {
    if (cfg % 1000) c.bn = cfg % 1000;
    if (cfg / 1000) c.pair = (cfg / 1000) - 1;
  }
  return c;
}

// A [M,K] lda, W [N,K] ldw (16-bit) -> out.  p.M/N/K and epilogue fields must be filled in.
template <typename T16>
static int gemm(cpt_handle* h, cudaStream_t st, int tag, const void* A, long long lda, const void* W, long long ldw,
                GemmParams p, int epi, bool out_fp32, int cfg = 0) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return 0;
  ProfScope ps(h, st, tag);
  GemmChoice c = pick_gemm(h, tag, p.M, p.N, p.K, cfg, p.ksplit);
  if (p.trans) c.pair = 0;
  p.trace = h->trace;
  if (h->trace) CK(cudaMemsetAsync(h->trace, 0, (size_t)h->num_sms * 128, st));
  const int dt = Cvt<T16>::kFmt;
  CUtensorMap ta, tb;
  // a transposed operand ([K, M] / [K, N] in memory) is fetched as 64 x 64 boxes, k along the rows
  if (p.trans & 1) TRY(make_tmap(&ta, A, dt, p.K, p.M, lda, 64));
  else TRY(make_tmap(&ta, A, dt, p.M, p.K, lda, kGemmBM));
  if (p.trans & 2) TRY(make_tmap(&tb, W, dt, p.K, p.N, ldw, 64));
  else TRY(make_tmap(&tb, W, dt, p.N, p.K, ldw, c.pair ? c.bn / 2 : c.bn));
  // outputs leave through TMA bulk stores (32x32 blocks) whenever the destination is 16-byte aligned and pitched;
  // otherwise (e.g. the [rows, 30522] fp32 score matrix) through the LSU path
  const unsigned osz = out_fp32 ? 4 : 2;
  CUtensorMap to = ta;
  const int want_reduce = p.tma_reduce;
  p.tma_store = 0;
  p.tma_reduce = 0;
  if (epi != EPI_BIAS_RESID && h->tma_store && !(reinterpret_cast<uintptr_t>(p.out) & 15) && !((p.ldo * osz) & 15)) {
    TRY(make_tmap_ex(&to, p.out, out_fp32 ? 2 : dt, p.M, p.N, p.ldo, 32, 32,
                     out_fp32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B));
    p.tma_store = 1;
    p.tma_reduce = want_reduce && out_fp32;
  }
  if (want_reduce && !p.tma_reduce) return fail("GEMM: accumulate-into-output needs an aligned fp32 destination");
  if (p.ksplit > 1) {
    // split-K only where partial products can be added at the destination; pieces of at least 4 k-blocks
    const int num_kb = (p.K + kGemmBK - 1) / kGemmBK;
    if (!p.tma_reduce) p.ksplit = 1;
    else p.ksplit = std::max(1, std::min(p.ksplit, num_kb / 4));
  }
  if (epi == EPI_BIAS && !out_fp32) return launch_gemm_bn<EPI_BIAS, T16, T16>(h, st, c, ta, tb, to, p);
  if (epi == EPI_BIAS && out_fp32) return launch_gemm_bn<EPI_BIAS, float, T16>(h, st, c, ta, tb, to, p);
  if (epi == EPI_BIAS_GELU && !out_fp32) return launch_gemm_bn<EPI_BIAS_GELU, T16, T16>(h, st, c, ta, tb, to, p);
  if (epi == EPI_BIAS_GELU && out_fp32) return launch_gemm_bn<EPI_BIAS_GELU, float, T16>(h, st, c, ta, tb, to, p);
  if (epi == EPI_BIAS_RESID && out_fp32) return launch_gemm_bn<EPI_BIAS_RESID, float, T16>(h, st, c, ta, tb, to, p);
  return fail("unsupported GEMM epilogue/output combination (epi=%d out_fp32=%d)", epi, (int)out_fp32);
}

template <typename T16>
static int attention(cpt_handle* h, cudaStream_t st, const void* qkv, const float* ext_mask, int B, int S, void* ctx,
                     int impl) {
  const int H = h->cfg.hidden_size, nH = h->cfg.num_attention_heads;
  if (H != nH * kAttnDH) return fail("attention kernel requires head size 64 (hidden %d, heads %d)", H, nH);
  if (S < 1 || S > 256) return fail("attention kernel supports 1 <= S <= 256 (got %d)", S);
  AttnParams p{B, S, H, nH, ext_mask, ctx, 0.125f, h->trace, Drop{0u, 0u, 0u, 0u, 1.f}};
  if (impl == 0 && h->pub_ready != nullptr && !h->trace) {   // the chain launch just before published per-tile counters
    p.qkv_ready = h->pub_ready;
    p.qkv_target = h->pub_target;
  }
  h->pub_ready = nullptr;
  h->ctx_published = nullptr;
  if (impl == 0 && h->ctx_counters != nullptr && !h->trace) {
    p.ctx_done = h->ctx_counters;
    h->ctx_published = h->ctx_counters;
  }
  h->ctx_counters = nullptr;
  if (h->trace) CK(cudaMemsetAsync(h->trace, 0, (size_t)h->num_sms * 128, st));
  ProfScope ps(h, st, CPT_K_ATTN);
  if (impl == 1) {
    auto* fn = attn_simt_kernel<T16>;
    const size_t smem = (size_t)S * kAttnDH * 2 * 2 + (size_t)S * 4;
    TRY(set_smem_attr(fn, smem));
    CK(launch_k(fn, dim3(nH, B), dim3(128), smem, st, 1, reinterpret_cast<const T16*>(qkv), p));
    return 0;
  }
  CUtensorMap tq;
  TRY(make_tmap(&tq, qkv, Cvt<T16>::kFmt, (unsigned long long)B * S, 3ull * H, 3ull * H, 64));
  if (impl == 0) {  // production: ping-pong softmax groups, one thread per query row
    const int items = B * nH * ((S + 127) / 128);
    const int grid = items < h->num_sms ? items : h->num_sms;
    CUtensorMap tc;
    TRY(make_tmap_bsh(&tc, ctx, Cvt<T16>::kFmt, B, S, H, 32));
#define CPT_ATTN_PP_CASE(N_)                                                                       \
  case N_: {                                                                                       \
    auto* fn = attn_pp_kernel<T16, N_>;                                                            \
    static bool attr_set[64] = {};                                                                 \
    if (!attr_set[h->device & 63]) {                                                               \
      TRY(set_smem_attr(fn, AttnPPCfg<N_>::kSmemBytes));                                           \
      attr_set[h->device & 63] = true;                                                             \
    }                                                                                              \
    CK(launch_k(fn, dim3(grid), dim3(kAttn2Threads), AttnPPCfg<N_>::kSmemBytes, st, 1, tq, tc, p)); \
    break;                                                                                         \
  }
    switch ((S + 63) / 64) {
      CPT_ATTN_PP_CASE(1)
      CPT_ATTN_PP_CASE(2)
      CPT_ATTN_PP_CASE(3)
      CPT_ATTN_PP_CASE(4)
      default: return fail("attention: unsupported S=%d", S);
    }
#undef CPT_ATTN_PP_CASE
    return 0;
  }
  if (impl == 3) {  // earlier design kept as a cross-check: persistent pipelined kernel, two threads per query row
    const int items = B * nH * ((S + 127) / 128);
    const int grid = items < h->num_sms ? items : h->num_sms;
    const int nch = (S + 63) / 64;
#define CPT_ATTN_CASE(N_)                                                              \
  case N_: {                                                                           \
    auto* fn = attn_pipe_kernel<T16, N_>;                                              \
    static bool attr_set[64] = {};                                                     \
    if (!attr_set[h->device & 63]) {                                                   \
      TRY(set_smem_attr(fn, Attn2Cfg<N_>::kSmemBytes));                                \
      attr_set[h->device & 63] = true;                                                 \
    }                                                                                  \
    CK(launch_k(fn, dim3(grid), dim3(kAttn2Threads), Attn2Cfg<N_>::kSmemBytes, st, 1, tq, p)); \
    break;                                                                             \
  }
    switch (nch) {
      CPT_ATTN_CASE(1)
      CPT_ATTN_CASE(2)
      CPT_ATTN_CASE(3)
      CPT_ATTN_CASE(4)
      default: return fail("attention: unsupported S=%d", S);
    }
#undef CPT_ATTN_CASE
    CKL("attn_pipe_kernel");
    return 0;
  }
  auto* fn = attn_tc_kernel<T16>;
  const size_t smem = attn_smem_bytes(S);
  TRY(set_smem_attr(fn, smem));
  CK(launch_k(fn, dim3(nH, (S + 127) / 128, B), dim3(kAttnThreads), smem, st, 1, tq, p));
  return 0;
}

template <typename T16>
static int layernorm(cpt_handle* h, cudaStream_t st, const float* x, long long ldx, int M, int H, const float* g, const float* b,
                     float eps, bool do_ln, float* o32, void* o16, int rin = 0, int rout = 0, int roff = 0,
                     const float* resid = nullptr, const void* x16 = nullptr) {
  if (M <= 0) return 0;
  ProfScope ps(h, st, CPT_K_LN);
  // persistent: 2 CTAs of 8 warps per SM, each warp walks rows with the next row's loads in flight
  const int ln_grid = (M + 7) / 8 < h->ln_ctas_per_sm * h->num_sms ? (M + 7) / 8 : h->ln_ctas_per_sm * h->num_sms;
#define CPT_LN_CASE(NV_)                                                                                             \
  case NV_:                                                                                                          \
    CK(launch_k(ln_rows_kernel<T16, NV_>, dim3(ln_grid), dim3(256), 0, st, 1, x, reinterpret_cast<const T16*>(x16), ldx, \
                resid, M, H, g, b, eps, do_ln ? 1 : 0, o32, reinterpret_cast<T16*>(o16), rin, rout, roff));          \
    break;
  switch (H / 128) {
    CPT_LN_CASE(1) CPT_LN_CASE(2) CPT_LN_CASE(3) CPT_LN_CASE(4) CPT_LN_CASE(5) CPT_LN_CASE(6) CPT_LN_CASE(7) CPT_LN_CASE(8)
    default: return fail("layernorm: unsupported hidden size %d", H);
  }
#undef CPT_LN_CASE
  return 0;
}

static int head_matvec(cpt_handle* h, cudaStream_t st, const float* X, long long ldx, int x_rows_per_b,
                       const long long* x_pos, const float* ln_g, const float* ln_b, float eps, const float* W,
                       long long ldw, const float* bias, const long long* w_ids, int w_rows, int B, int H, int O,
                       int act, float* Y, long long ldy, const HeadPeers& peers = HeadPeers()) {
  if (B <= 0 || O <= 0) return 0;
  const size_t smem = (size_t)kHeadRows * H * sizeof(float);
  dim3 grid((B + kHeadRows - 1) / kHeadRows, (O + kHeadOuts - 1) / kHeadOuts);
  ProfScope ps(h, st, CPT_K_HEAD);
  CK(launch_k(head_matvec_kernel, grid, dim3(256), smem, st, 1, X, ldx, x_rows_per_b, x_pos, ln_g, ln_b, eps, W, ldw,
              bias, w_ids, w_rows, B, H, O, act, Y, ldy, h->err_flag, peers));
  return 0;
}

// ------------------------------------------------------------------------------------------------ workspace
static size_t al(size_t x) { return (x + 255) & ~size_t(255); }
struct Workspace {
  float *ext_mask, *h32, *a32, *pre32, *head_t, *stats;
  char* pre16;
  size_t stats_bytes;
  char *h16, *a16, *ctx16, *qkv16, *inter16, *img16;
  unsigned* flags;     // readiness counters of the chain kernel: [L][6 stages][2][M tiles]
  size_t flags_bytes;
  float2* part;        // row-statistics scratch of the chain kernel's fused LayerNorm epilogues
  size_t total;
};
static Workspace carve(const cpt_handle* h, int B, int T, int R, char* base) {
  const cpt_config& c = h->cfg;
  const size_t M = (size_t)B * (T + R), H = c.hidden_size, I = c.intermediate_size;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += al(bytes);
    return p;
  };
  Workspace w;
  w.ext_mask = (float*)take(M * 4);
  w.h32 = (float*)take(M * H * 4);
  w.a32 = (float*)take(M * H * 4);
  w.pre32 = (float*)take(M * H * 4);
  w.head_t = (float*)take((size_t)B * H * 4);
  w.h16 = take(M * H * 2);
  w.a16 = take(M * H * 2);
  w.ctx16 = take(M * H * 2);
  w.pre16 = take(M * H * 2);
  w.qkv16 = take(M * 3 * H * 2);
  w.inter16 = take(M * I * 2);
  w.img16 = take((size_t)B * R * h->Fp * 2);
  w.stats_bytes = (size_t)(c.num_hidden_layers > 0 ? c.num_hidden_layers : 1) * 2 * M * 2 * 4;  // [L][2][M](sum, sumsq)
  w.stats = (float*)take(w.stats_bytes);
  w.flags_bytes = (size_t)(c.num_hidden_layers > 0 ? c.num_hidden_layers : 1) * chain_counter_bytes((int)M, 6);
  w.flags = (unsigned*)take(w.flags_bytes);
  w.part = (float2*)take(chain_part_bytes((int)M, (int)H));
  w.total = off + 256;
  return w;
}

#include "chain_host.inl"

// ------------------------------------------------------------------------------------------------ weights
template <typename T16>
static int cast_now(cudaStream_t st, const float* src, long long rows, int cols, int ldo, void* dst) {
  const long long total = rows * ldo;
  const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  cast_weight_kernel<T16><<<grid, 256, 0, st>>>(src, rows, cols, ldo, reinterpret_cast<T16*>(dst));
  CKL("cast_weight_kernel");
  return 0;
}

// weight copies are RECORDED while cpt_set_weights walks the state dict and issued as one launch at its end
static void rec_copy(cpt_handle* h, const float* src, void* dst, long long rows, int cols, int ldo, int is16) {
  h->copy_list.push_back(CopyEntry{src, dst, rows, cols, ldo, is16, 0});
}
template <typename T16>
static int cast_w(cpt_handle* h, const float* src, long long rows, int cols, int ldo, void* dst) {
  rec_copy(h, src, dst, rows, cols, ldo, 1);
  return 0;
}
// (re)allocation policy of cpt_set_weights: a training handle whose previous call covered the same tensors keeps
// its device buffers and only refreshes their contents (one optimizer step later the shapes cannot have changed)
static int walloc(cpt_handle* h, bool reuse, void** p, size_t bytes) {
  if (reuse && *p) return 0;
  return dev_alloc(h, p, bytes);
}
static int copy_vec(cpt_handle* h, bool reuse, cudaStream_t st, const float* src, size_t n, float** dst) {
  if (!src) {
    *dst = nullptr;
    return 0;
  }
  TRY(walloc(h, reuse, (void**)dst, n * 4));
  rec_copy(h, src, *dst, 1, (int)n, (int)n, 0);
  return 0;
}
template <typename T16>
static int transpose16(cudaStream_t st, const void* in, int R, int C, long long ld_in, void* out, long long ld_out) {
  dim3 grid((C + 31) / 32, (int)((std::max<long long>(R, ld_out) + 31) / 32));
  transpose16_kernel<T16><<<grid, 256, 0, st>>>(reinterpret_cast<const T16*>(in), R, C, ld_in,
                                                 reinterpret_cast<T16*>(out), ld_out);
  CKL("transpose16_kernel");
  return 0;
}

// issue the recorded copies: one launch; the device tables are reused while the (source, destination) list is unchanged
template <typename T16>
static int flush_copies(cpt_handle* h, cudaStream_t st) {
  if (h->copy_list.empty()) return 0;
  const size_t n = h->copy_list.size();
  const bool same = h->copy_entries_dev && h->copy_list_dev_mirror.size() == n &&
                    memcmp(h->copy_list_dev_mirror.data(), h->copy_list.data(), n * sizeof(CopyEntry)) == 0;
  if (!same) {
    std::vector<CopyChunk> chunks;
    for (size_t e = 0; e < n; ++e) {
      const long long total = h->copy_list[e].rows * h->copy_list[e].ldo;
      for (long long f = 0; f < total; f += kCopyChunk)
        chunks.push_back(CopyChunk{(int)e, (int)std::min<long long>(kCopyChunk, total - f), f});
    }
    CK(cudaStreamSynchronize(st));  // an earlier launch may still be reading the old tables
    if (h->copy_entries_dev) cudaFree(h->copy_entries_dev);
    if (h->copy_chunks_dev) cudaFree(h->copy_chunks_dev);
    h->copy_entries_dev = h->copy_chunks_dev = nullptr;
    CK(cudaMalloc(&h->copy_entries_dev, n * sizeof(CopyEntry)));
    CK(cudaMalloc(&h->copy_chunks_dev, std::max<size_t>(1, chunks.size()) * sizeof(CopyChunk)));
    CK(cudaMemcpy(h->copy_entries_dev, h->copy_list.data(), n * sizeof(CopyEntry), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->copy_chunks_dev, chunks.data(), chunks.size() * sizeof(CopyChunk), cudaMemcpyHostToDevice));
    h->copy_n_chunks = (int)chunks.size();
    h->copy_list_dev_mirror = h->copy_list;
  }
  if (h->copy_n_chunks > 0) {
    refresh_weights_kernel<T16><<<h->copy_n_chunks, 256, 0, st>>>(
        reinterpret_cast<const CopyEntry*>(h->copy_entries_dev), reinterpret_cast<const CopyChunk*>(h->copy_chunks_dev));
    CKL("refresh_weights_kernel");
  }
  return 0;
}

template <typename T16>
static int set_weights_impl(cpt_handle* h, const cpt_weights* w, cudaStream_t st) {
  const cpt_config& c = h->cfg;
  const int H = c.hidden_size, I = c.intermediate_size, L = c.num_hidden_layers, F = c.img_feature_dim;
  if (!w->word_emb || !w->pos_emb || !w->type_emb || !w->emb_ln_g || !w->emb_ln_b || !w->layers)
    return fail("cpt_set_weights: embedding tables / LayerNorm / layers must be non-NULL");
  const bool has_pooler = w->pooler_w && w->pooler_b;
  const bool has_mlm = w->mlm_dense_w && w->mlm_dense_b && w->mlm_ln_g && w->mlm_ln_b && w->mlm_bias;
  const bool has_nsp = w->nsp_w && w->nsp_b;
  const unsigned sig = 1u | (w->img_w ? 2u : 0u) | (w->img_ln_g ? 4u : 0u) | (has_pooler ? 8u : 0u) |
                       (has_mlm ? 16u : 0u) | (has_nsp ? 32u : 0u) | (h->train ? 64u : 0u);
  const bool reuse = h->train && h->has_weights && h->weights_sig == sig;
  if (!reuse) {
    CK(cudaDeviceSynchronize());  // no kernel may still be reading the buffers we are about to free
    int* flag = h->err_flag;
    h->owned.erase(std::remove(h->owned.begin(), h->owned.end(), (void*)flag), h->owned.end());
    h->owned.erase(std::remove(h->owned.begin(), h->owned.end(), (void*)h->trace), h->owned.end());
    free_owned(h);
    h->owned.push_back(flag);
    if (h->trace) h->owned.push_back(h->trace);
    h->emb_g = h->emb_b = h->b_img = h->img_g = h->img_b = h->pool_w = h->pool_b = nullptr;
    h->mlm_w = h->mlm_b = h->mlm_g = h->mlm_beta = h->mlm_bias = h->nsp_w = h->nsp_b = nullptr;
    h->w_img = h->mlm_w16 = h->word16 = nullptr;
    h->layers.assign(L, LayerDev{});
  }
  h->has_weights = false;
  h->copy_list.clear();
  h->word = w->word_emb;
  h->pos = w->pos_emb;
  h->type = w->type_emb;
  TRY(copy_vec(h, reuse, st, w->emb_ln_g, H, &h->emb_g));
  TRY(copy_vec(h, reuse, st, w->emb_ln_b, H, &h->emb_b));
  if (w->img_w) {
    if (!w->img_b) return fail("cpt_set_weights: img_w without img_b");
    if (c.use_img_layernorm && (!w->img_ln_g || !w->img_ln_b))
      return fail("cpt_set_weights: use_img_layernorm=1 needs bert.LayerNorm weights");
    TRY(walloc(h, reuse, &h->w_img, (size_t)H * h->Fp * 2));
    TRY(cast_w<T16>(h, w->img_w, H, F, h->Fp, h->w_img));
    TRY(copy_vec(h, reuse, st, w->img_b, H, &h->b_img));
    TRY(copy_vec(h, reuse, st, w->img_ln_g, H, &h->img_g));
    TRY(copy_vec(h, reuse, st, w->img_ln_b, H, &h->img_b));
  }
  for (int l = 0; l < L; ++l) {
    const cpt_layer_weights& s = w->layers[l];
    LayerDev& d = h->layers[l];
    const float* need[] = {s.q_w, s.q_b, s.k_w, s.k_b, s.v_w, s.v_b, s.ao_w, s.ao_b, s.ao_ln_g, s.ao_ln_b,
                           s.i_w, s.i_b, s.o_w, s.o_b, s.o_ln_g, s.o_ln_b};
    for (const float* q : need)
      if (!q) return fail("cpt_set_weights: layer %d has a NULL tensor", l);
    TRY(walloc(h, reuse, &d.w_qkv, (size_t)3 * H * H * 2));
    TRY(cast_w<T16>(h, s.q_w, H, H, H, d.w_qkv));
    TRY(cast_w<T16>(h, s.k_w, H, H, H, (char*)d.w_qkv + (size_t)H * H * 2));
    TRY(cast_w<T16>(h, s.v_w, H, H, H, (char*)d.w_qkv + (size_t)2 * H * H * 2));
    TRY(walloc(h, reuse, (void**)&d.b_qkv, (size_t)3 * H * 4));
    rec_copy(h, s.q_b, d.b_qkv, 1, H, H, 0);
    rec_copy(h, s.k_b, d.b_qkv + H, 1, H, H, 0);
    rec_copy(h, s.v_b, d.b_qkv + 2 * H, 1, H, H, 0);
    TRY(walloc(h, reuse, &d.w_ao, (size_t)H * H * 2));
    TRY(cast_w<T16>(h, s.ao_w, H, H, H, d.w_ao));
    TRY(copy_vec(h, reuse, st, s.ao_b, H, &d.b_ao));
    TRY(copy_vec(h, reuse, st, s.ao_ln_g, H, &d.ao_g));
    TRY(copy_vec(h, reuse, st, s.ao_ln_b, H, &d.ao_b));
    TRY(walloc(h, reuse, &d.w_i, (size_t)I * H * 2));
    TRY(cast_w<T16>(h, s.i_w, I, H, H, d.w_i));
    TRY(copy_vec(h, reuse, st, s.i_b, I, &d.b_i));
    TRY(walloc(h, reuse, &d.w_o, (size_t)H * I * 2));
    TRY(cast_w<T16>(h, s.o_w, H, I, I, d.w_o));
    TRY(copy_vec(h, reuse, st, s.o_b, H, &d.b_o));
    TRY(copy_vec(h, reuse, st, s.o_ln_g, H, &d.o_g));
    TRY(copy_vec(h, reuse, st, s.o_ln_b, H, &d.o_b));
    if (h->train) continue;  // the LayerNorm-folded copies below serve the (inference-only) folded path
    // folded copies: FFN-up reads pre-LN1 rows (gamma/beta of this layer's attention.output.LayerNorm); the QKV
    // projection of layer l >= 1 reads pre-LN2 rows of layer l-1 (gamma/beta of its output.LayerNorm)
    TRY(dev_alloc(h, &d.w_i_f, (size_t)I * H * 2));
    TRY(dev_alloc(h, (void**)&d.g_i, (size_t)I * 4));
    TRY(dev_alloc(h, (void**)&d.c_i, (size_t)I * 4));
    fold_weight_kernel<T16><<<(I + 7) / 8, 256, 0, st>>>(s.i_w, s.ao_ln_g, s.ao_ln_b, s.i_b, I, H,
                                                         reinterpret_cast<T16*>(d.w_i_f), d.g_i, d.c_i);
    CKL("fold_weight_kernel");
    if (l > 0) {
      const cpt_layer_weights& pv = w->layers[l - 1];
      TRY(dev_alloc(h, &d.w_qkv_f, (size_t)3 * H * H * 2));
      TRY(dev_alloc(h, (void**)&d.g_qkv, (size_t)3 * H * 4));
      TRY(dev_alloc(h, (void**)&d.c_qkv, (size_t)3 * H * 4));
      const float* ws[3] = {s.q_w, s.k_w, s.v_w};
      const float* bs[3] = {s.q_b, s.k_b, s.v_b};
      for (int j = 0; j < 3; ++j) {
        fold_weight_kernel<T16><<<(H + 7) / 8, 256, 0, st>>>(
            ws[j], pv.o_ln_g, pv.o_ln_b, bs[j], H, H, reinterpret_cast<T16*>(d.w_qkv_f) + (size_t)j * H * H,
            d.g_qkv + j * H, d.c_qkv + j * H);
        CKL("fold_weight_kernel");
      }
    }
  }
  h->has_pooler = has_pooler;
  if (h->has_pooler) {
    TRY(copy_vec(h, reuse, st, w->pooler_w, (size_t)H * H, &h->pool_w));
    TRY(copy_vec(h, reuse, st, w->pooler_b, H, &h->pool_b));
  }
  h->has_mlm = has_mlm;
  if (h->has_mlm) {
    TRY(copy_vec(h, reuse, st, w->mlm_dense_w, (size_t)H * H, &h->mlm_w));
    TRY(copy_vec(h, reuse, st, w->mlm_dense_b, H, &h->mlm_b));
    TRY(copy_vec(h, reuse, st, w->mlm_ln_g, H, &h->mlm_g));
    TRY(copy_vec(h, reuse, st, w->mlm_ln_b, H, &h->mlm_beta));
    TRY(copy_vec(h, reuse, st, w->mlm_bias, c.vocab_size, &h->mlm_bias));
    // 16-bit copies for the full-vocabulary scores path (tensor-core GEMM over all rows)
    TRY(walloc(h, reuse, &h->mlm_w16, (size_t)H * H * 2));
    TRY(cast_w<T16>(h, w->mlm_dense_w, H, H, H, h->mlm_w16));
    TRY(walloc(h, reuse, &h->word16, (size_t)c.vocab_size * H * 2));
    TRY(cast_w<T16>(h, w->word_emb, c.vocab_size, H, H, h->word16));
  }
  h->has_nsp = has_nsp;
  if (h->has_nsp) {
    TRY(copy_vec(h, reuse, st, w->nsp_w, (size_t)c.num_contrast_classes * H, &h->nsp_w));
    TRY(copy_vec(h, reuse, st, w->nsp_b, c.num_contrast_classes, &h->nsp_b));
  }
  TRY(flush_copies<T16>(h, st));
  h->has_weights = true;
  h->weights_sig = sig;
  return 0;
}

// ------------------------------------------------------------------------------------------------ forward
template <typename T16>
static int encoder_forward_impl(cpt_handle* h, cudaStream_t st, const int64_t* ids, const int64_t* seg,
                                const int64_t* mask, const int64_t* pos_ids, const float* img, int B, int T, int R,
                                void* ws_ptr, size_t ws_bytes, float* seq_out, float* pooled, float* hidden_states) {
  const cpt_config& c = h->cfg;
  const int H = c.hidden_size, I = c.intermediate_size, L = c.num_hidden_layers, S = T + R, M = B * S;
  if (!h->has_weights) return fail("cpt_encoder_forward called before cpt_set_weights");
  if (B <= 0 || T <= 0 || R < 0) return fail("bad shape B=%d T=%d R=%d", B, T, R);
  if (T > c.max_position_embeddings) return fail("T=%d exceeds max_position_embeddings=%d", T, c.max_position_embeddings);
  if (R > 0 && (!img || !h->w_img)) return fail("img_feats given but no img_embedding weights (or NULL img_feats)");
  if (S > 256) return fail("sequence length T+R=%d exceeds the 256 this build's attention kernel supports", S);
  Workspace w = carve(h, B, T, R, (char*)(((uintptr_t)ws_ptr + 255) & ~uintptr_t(255)));
  if (!ws_ptr || ws_bytes < w.total) return fail("workspace too small: need %zu bytes, got %zu", w.total, ws_bytes);
  if (!ids || !seq_out) return fail("input_ids and seq_out must be non-NULL");

  h->pub_ready = nullptr;
  h->ctx_counters = nullptr;
  h->ctx_published = nullptr;
  const bool chain_h_ok = H == 128 || H == 256 || H == 512 || H == 768 || H == 1024;
  const bool use_chain = h->chain && !h->train && !(h->fold_ln && !hidden_states) && L > 0 && h->tma_store &&
                         h->reduce_resid && M >= h->chain_min_rows && chain_h_ok;
  const size_t chain_ctr_per_layer = chain_counter_bytes(M, 6);
  if (use_chain) CK(cudaMemsetAsync(w.flags, 0, chain_ctr_per_layer * L, st));  // first: keeps the PDL chain unbroken
  // K4: additive mask
  if (mask) {
    ProfScope ps(h, st, CPT_K_EXTMASK);
    CK(launch_k(ext_mask_kernel, dim3((M + 255) / 256), dim3(256), 0, st, 1, (const long long*)mask, M, w.ext_mask));
  } else {
    CK(cudaMemsetAsync(w.ext_mask, 0, (size_t)M * 4, st));
  }
  {  // K1: text rows
    ProfScope ps(h, st, CPT_K_EMBED);
#define CPT_EMB_CASE(NV_)                                                                                            \
  case NV_:                                                                                                          \
    CK(launch_k(embed_text_ln_kernel<T16, NV_>, dim3((B * T + 7) / 8), dim3(256), 0, st, 1, (const long long*)ids,    \
                (const long long*)seg, (const long long*)pos_ids, h->word, h->pos, h->type, (const float*)h->emb_g,  \
                (const float*)h->emb_b, c.layer_norm_eps, B, T, S, H, c.vocab_size, c.max_position_embeddings,       \
                c.type_vocab_size, w.h32, reinterpret_cast<T16*>(w.h16), h->err_flag));                              \
    break;
    switch (H / 128) {
      CPT_EMB_CASE(1) CPT_EMB_CASE(2) CPT_EMB_CASE(3) CPT_EMB_CASE(4) CPT_EMB_CASE(5) CPT_EMB_CASE(6) CPT_EMB_CASE(7)
      CPT_EMB_CASE(8)
      default: return fail("unsupported hidden size %d", H);
    }
#undef CPT_EMB_CASE
  }
  // K2+K3: region rows
  if (R > 0) {
    const int F = c.img_feature_dim, Mi = B * R;
    const int grid = (Mi + 7) / 8 < 8 * h->num_sms ? (Mi + 7) / 8 : 8 * h->num_sms;
    {
      ProfScope ps(h, st, CPT_K_CAST);
      CK(launch_k(cast_pad_kernel<T16>, dim3(grid), dim3(256), 0, st, 1, img, Mi, F, h->Fp,
                  reinterpret_cast<T16*>(w.img16)));
    }
    GemmParams p{};
    p.M = Mi; p.N = H; p.K = F; p.out = w.pre32; p.ldo = H; p.bias = h->b_img;
    TRY(gemm<T16>(h, st, CPT_K_GEMM_IMG, w.img16, h->Fp, h->w_img, h->Fp, p, EPI_BIAS, true));
    TRY(layernorm<T16>(h, st, w.pre32, H, Mi, H, h->img_g, h->img_b, c.img_layer_norm_eps, c.use_img_layernorm != 0,
                       w.h32, w.h16, R, S, T));
  }
  if (hidden_states) CK(cudaMemcpyAsync(hidden_states, w.h32, (size_t)M * H * 4, cudaMemcpyDeviceToDevice, st));

  const bool fold = h->fold_ln && !h->train && !hidden_states && L > 0;
  if (use_chain) {
    // two launches per layer: the attention kernel, then ONE dataflow launch for attention-out + LayerNorm + FFN-up +
    // FFN-down + LayerNorm + the next layer's QKV projection (chain_sm100.cuh)
    {
      GemmParams p{};
      p.M = M; p.N = 3 * H; p.K = H; p.out = w.qkv16; p.ldo = 3 * H; p.bias = h->layers[0].b_qkv;
      TRY(gemm<T16>(h, st, CPT_K_GEMM_QKV, w.h16, H, h->layers[0].w_qkv, H, p, EPI_BIAS, false));
    }
    const int saved_mode = h->chain_fuse_ln;
    if (hidden_states && h->chain_fuse_ln == 2) h->chain_fuse_ln = 1;  // per-layer outputs need finished LayerNorms
    int rc = 0;
    for (int l = 0; l < L && !rc; ++l) {
      unsigned* ctr_l = w.flags + (size_t)l * (chain_ctr_per_layer / sizeof(unsigned));
      // deferred-LayerNorm chains use four of the six counter pairs: the attention launch counts its context rows in a
      // free one, and the chain launch starts inside its tail
      h->ctx_counters = (h->chain_early && h->chain_fuse_ln == 2 && h->chain_lean) ? ctr_l + (size_t)10 * chain_m_tiles2(M)
                                                                                    : nullptr;
      rc = attention<T16>(h, st, w.qkv16, w.ext_mask, B, S, w.ctx16, h->attn_impl);
      float* o32 = (l == L - 1 && h->chain_fuse_ln != 2) ? seq_out : w.h32;
      if (!rc)
        rc = chain_layer<T16>(h, st, w, l, M, o32, l == L - 1,
                              w.flags + (size_t)l * (chain_ctr_per_layer / sizeof(unsigned)));
      if (!rc && hidden_states &&
          cudaMemcpyAsync(hidden_states + (size_t)(l + 1) * M * H, o32, (size_t)M * H * 4, cudaMemcpyDeviceToDevice,
                          st) != cudaSuccess)
        rc = fail("cudaMemcpyAsync(hidden_states) failed");
    }
    if (!rc && h->chain_fuse_ln == 2) {  // the last layer's BertOutput.LayerNorm: h32 holds its pre-LayerNorm rows
      const LayerDev& last = h->layers[L - 1];
      rc = layernorm<T16>(h, st, w.h32, H, M, H, last.o_g, last.o_b, c.layer_norm_eps, true, seq_out, nullptr);
    }
    h->chain_fuse_ln = saved_mode;
    TRY(rc);
  } else if (fold) {
    // LayerNorm folded into the GEMMs on either side of it (DESIGN.md "LayerNorm folding"): the stream buffers hold
    // PRE-LayerNorm rows (fp32 + 16-bit) plus per-row (sum, sum of squares); no LayerNorm kernel runs between GEMMs.
    CK(cudaMemsetAsync(w.stats, 0, w.stats_bytes, st));
    auto stats = [&](int l, int which) { return w.stats + ((size_t)l * 2 + which) * (size_t)M * 2; };
    for (int l = 0; l < L; ++l) {
      const LayerDev& d = h->layers[l];
      {  // K5: QKV projection; for l >= 1 its A operand is the pre-LN2 stream of layer l-1
        GemmParams p{};
        p.M = M; p.N = 3 * H; p.K = H; p.out = w.qkv16; p.ldo = 3 * H;
        const void* W = d.w_qkv;
        p.bias = d.b_qkv;
        if (l > 0) {
          W = d.w_qkv_f; p.bias = d.c_qkv; p.gvec = d.g_qkv; p.nstats = stats(l - 1, 1);
          p.neps = c.layer_norm_eps; p.nH = H;
        }
        TRY(gemm<T16>(h, st, CPT_K_GEMM_QKV, w.h16, H, W, H, p, EPI_BIAS, false));
      }
      TRY(attention<T16>(h, st, w.qkv16, w.ext_mask, B, S, w.ctx16, h->attn_impl));  // K6-K9
      {  // K10: x1 = ctx Wo^T + b + LN2_{l-1}(x2) -> fp32 + 16-bit + row statistics
        GemmParams p{};
        p.M = M; p.N = H; p.K = H; p.out = w.a32; p.ldo = H; p.bias = d.b_ao; p.resid = w.h32; p.ldr = H;
        if (l > 0) {
          p.rstats = stats(l - 1, 1); p.rgamma = h->layers[l - 1].o_g; p.rbeta = h->layers[l - 1].o_b;
          p.reps = c.layer_norm_eps;
        }
        p.nH = H; p.out16 = w.a16; p.ldo16 = H; p.stats_out = stats(l, 0);
        TRY(gemm<T16>(h, st, CPT_K_GEMM_AO, w.ctx16, H, d.w_ao, H, p, EPI_BIAS_RESID, true));
      }
      {  // K11: gelu(LN1(x1) W1^T + b1) with LN1 folded
        GemmParams p{};
        p.M = M; p.N = I; p.K = H; p.out = w.inter16; p.ldo = I; p.bias = d.c_i; p.gvec = d.g_i;
        p.nstats = stats(l, 0); p.neps = c.layer_norm_eps; p.nH = H;
        TRY(gemm<T16>(h, st, CPT_K_GEMM_UP, w.a16, H, d.w_i_f, H, p, EPI_BIAS_GELU, false));
      }
      {  // K12: x2 = inter W2^T + b2 + LN1(x1) -> fp32 + 16-bit + row statistics
        GemmParams p{};
        p.M = M; p.N = H; p.K = I; p.out = w.h32; p.ldo = H; p.bias = d.b_o; p.resid = w.a32; p.ldr = H;
        p.rstats = stats(l, 0); p.rgamma = d.ao_g; p.rbeta = d.ao_b; p.reps = c.layer_norm_eps; p.nH = H;
        p.out16 = (l == L - 1) ? nullptr : w.h16; p.ldo16 = H;
        p.stats_out = (l == L - 1) ? nullptr : stats(l, 1);
        TRY(gemm<T16>(h, st, CPT_K_GEMM_DOWN, w.inter16, I, d.w_o, I, p, EPI_BIAS_RESID, true));
      }
    }
    const LayerDev& last = h->layers[L - 1];
    TRY(layernorm<T16>(h, st, w.h32, H, M, H, last.o_g, last.o_b, c.layer_norm_eps, true, seq_out, nullptr));
  } else {
  for (int l = 0; l < L; ++l) {
      const LayerDev& d = h->layers[l];
      {  // K5
        GemmParams p{};
        p.M = M; p.N = 3 * H; p.K = H; p.out = w.qkv16; p.ldo = 3 * H; p.bias = d.b_qkv;
        TRY(gemm<T16>(h, st, CPT_K_GEMM_QKV, w.h16, H, d.w_qkv, H, p, EPI_BIAS, false));
      }
      TRY(attention<T16>(h, st, w.qkv16, w.ext_mask, B, S, w.ctx16, h->attn_impl));  // K6-K9
      {  // K10
        GemmParams p{};
        p.M = M; p.N = H; p.K = H; p.out = w.pre32; p.ldo = H; p.bias = d.b_ao;
        if (h->reduce_resid && h->tma_store) {
          // x1 = h + (ctx Wo^T + b): the epilogue's bulk tensor stores ADD into the residual buffer at L2; LayerNorm
          // then reads one fp32 tensor instead of two
          p.out = w.h32; p.tma_reduce = 1;
          TRY(gemm<T16>(h, st, CPT_K_GEMM_AO, w.ctx16, H, d.w_ao, H, p, EPI_BIAS, true));
          TRY(layernorm<T16>(h, st, w.h32, H, M, H, d.ao_g, d.ao_b, c.layer_norm_eps, true, w.a32, w.a16));
        } else if (h->resid_in_ln && h->delta16) {
          p.out = w.pre16;
          TRY(gemm<T16>(h, st, CPT_K_GEMM_AO, w.ctx16, H, d.w_ao, H, p, EPI_BIAS, false));
          TRY(layernorm<T16>(h, st, nullptr, H, M, H, d.ao_g, d.ao_b, c.layer_norm_eps, true, w.a32, w.a16, 0, 0, 0,
                             w.h32, w.pre16));
        } else if (h->resid_in_ln) {
          TRY(gemm<T16>(h, st, CPT_K_GEMM_AO, w.ctx16, H, d.w_ao, H, p, EPI_BIAS, true));
          TRY(layernorm<T16>(h, st, w.pre32, H, M, H, d.ao_g, d.ao_b, c.layer_norm_eps, true, w.a32, w.a16, 0, 0, 0,
                             w.h32));
        } else {
          p.resid = w.h32; p.ldr = H;
          TRY(gemm<T16>(h, st, CPT_K_GEMM_AO, w.ctx16, H, d.w_ao, H, p, EPI_BIAS_RESID, true));
          TRY(layernorm<T16>(h, st, w.pre32, H, M, H, d.ao_g, d.ao_b, c.layer_norm_eps, true, w.a32, w.a16));
        }
      }
      {  // K11
        GemmParams p{};
        p.M = M; p.N = I; p.K = H; p.out = w.inter16; p.ldo = I; p.bias = d.b_i;
        TRY(gemm<T16>(h, st, CPT_K_GEMM_UP, w.a16, H, d.w_i, H, p, EPI_BIAS_GELU, false));
      }
      {  // K12
        GemmParams p{};
        p.M = M; p.N = H; p.K = I; p.out = w.pre32; p.ldo = H; p.bias = d.b_o;
        float* o32 = (l == L - 1) ? seq_out : w.h32;
        if (h->reduce_resid && h->tma_store) {
          p.out = w.a32; p.tma_reduce = 1; p.ksplit = h->down_ksplit;
          TRY(gemm<T16>(h, st, CPT_K_GEMM_DOWN, w.inter16, I, d.w_o, I, p, EPI_BIAS, true));
          TRY(layernorm<T16>(h, st, w.a32, H, M, H, d.o_g, d.o_b, c.layer_norm_eps, true, o32,
                             (l == L - 1) ? nullptr : w.h16));
        } else if (h->resid_in_ln && h->delta16) {
          p.out = w.pre16;
          TRY(gemm<T16>(h, st, CPT_K_GEMM_DOWN, w.inter16, I, d.w_o, I, p, EPI_BIAS, false));
          TRY(layernorm<T16>(h, st, nullptr, H, M, H, d.o_g, d.o_b, c.layer_norm_eps, true, o32,
                             (l == L - 1) ? nullptr : w.h16, 0, 0, 0, w.a32, w.pre16));
        } else if (h->resid_in_ln) {
          TRY(gemm<T16>(h, st, CPT_K_GEMM_DOWN, w.inter16, I, d.w_o, I, p, EPI_BIAS, true));
          TRY(layernorm<T16>(h, st, w.pre32, H, M, H, d.o_g, d.o_b, c.layer_norm_eps, true, o32,
                             (l == L - 1) ? nullptr : w.h16, 0, 0, 0, w.a32));
        } else {
          p.resid = w.a32; p.ldr = H;
          TRY(gemm<T16>(h, st, CPT_K_GEMM_DOWN, w.inter16, I, d.w_o, I, p, EPI_BIAS_RESID, true));
          TRY(layernorm<T16>(h, st, w.pre32, H, M, H, d.o_g, d.o_b, c.layer_norm_eps, true, o32,
                             (l == L - 1) ? nullptr : w.h16));
        }
        if (hidden_states)
          CK(cudaMemcpyAsync(hidden_states + (size_t)(l + 1) * M * H, o32, (size_t)M * H * 4,
                             cudaMemcpyDeviceToDevice, st));
      }
    }
  }
  if (L == 0) CK(cudaMemcpyAsync(seq_out, w.h32, (size_t)M * H * 4, cudaMemcpyDeviceToDevice, st));
  if (pooled) {  // K13
    if (!h->has_pooler) return fail("pooled output requested but bert.pooler weights were not provided");
    TRY(head_matvec(h, st, seq_out, H, S, nullptr, nullptr, nullptr, 0.f, h->pool_w, H, h->pool_b, nullptr, H, B, H,
                    H, ACT_TANH, pooled, H));
  }
  return 0;
}

template <typename T16>
static int mlm_scores_impl(cpt_handle* h, cudaStream_t st, const float* seq_out, long long rows, void* ws_ptr,
                           size_t ws_bytes, float* scores) {
  const cpt_config& c = h->cfg;
  const int H = c.hidden_size, V = c.vocab_size;
  const size_t need = cpt_mlm_scores_workspace_bytes(h, rows);
  if (!ws_ptr || ws_bytes < need) return fail("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
  char* base = (char*)(((uintptr_t)ws_ptr + 255) & ~uintptr_t(255));
  char* x16 = base;
  float* t32 = (float*)(base + al((size_t)rows * H * 2));
  char* t16 = (char*)t32 + al((size_t)rows * H * 4);
  TRY(layernorm<T16>(h, st, seq_out, H, (int)rows, H, nullptr, nullptr, 0.f, false, nullptr, x16));  // fp32 -> 16-bit
  GemmParams p{};
  p.M = (int)rows; p.N = H; p.K = H; p.out = t32; p.ldo = H; p.bias = h->mlm_b;
  TRY(gemm<T16>(h, st, CPT_K_GEMM_HEAD, x16, H, h->mlm_w16, H, p, EPI_BIAS_GELU, true));
  TRY(layernorm<T16>(h, st, t32, H, (int)rows, H, h->mlm_g, h->mlm_beta, c.layer_norm_eps, true, nullptr, t16));
  GemmParams q{};
  q.M = (int)rows; q.N = V; q.K = H; q.out = scores; q.ldo = V; q.bias = h->mlm_bias;
  TRY(gemm<T16>(h, st, CPT_K_GEMM_HEAD, t16, H, h->word16, H, q, EPI_BIAS, true));
  return 0;
}

#define DISPATCH_DTYPE(h, CALL)                                  \
  ((h)->cfg.dtype == 0 ? CALL(__half) : CALL(__nv_bfloat16))

#include "train_host.inl"

// ================================================================================================ C ABI
extern "C" {

const char* cpt_kernel_name(int tag);
const char* cpt_last_error(void) { return g_err.c_str(); }
int cpt_abi_version(void) { return CPT_B200_ABI_VERSION; }

int cpt_create(const cpt_config* cfg, int device, cpt_handle** out) {
  if (!cfg || !out) return fail("cpt_create: NULL argument");
  *out = nullptr;
  const cpt_config& c = *cfg;
  if (c.hidden_size <= 0 || c.hidden_size % 128 || c.hidden_size > 1024)
    return fail("hidden_size must be a multiple of 128 and <= 1024 (got %d)", c.hidden_size);
  if (c.num_attention_heads <= 0 || c.hidden_size != c.num_attention_heads * 64)
    return fail("head size must be 64 (hidden %d / heads %d)", c.hidden_size, c.num_attention_heads);
  if (c.intermediate_size <= 0 || c.intermediate_size % 8) return fail("intermediate_size must be a multiple of 8");
  if (c.dtype != 0 && c.dtype != 1) return fail("dtype must be 0 (fp16) or 1 (bf16)");
  if (c.num_hidden_layers < 0 || c.vocab_size <= 0 || c.max_position_embeddings <= 0 || c.type_vocab_size <= 0)
    return fail("bad config");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail("device %d out of range (%d CUDA devices visible)", device, ndev);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail("cpt_b200 is built for sm_100a (B200) only; device %d is sm_%d%d — there is no fallback path", device,
                prop.major, prop.minor);
  DeviceGuard g(device);
  cpt_handle* h = new cpt_handle();
  h->cfg = c;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->Fp = (c.img_feature_dim + 7) & ~7;
  if (cudaMalloc((void**)&h->err_flag, 16) != cudaSuccess) {
    delete h;
    return fail("cudaMalloc failed");
  }
  cudaMemset(h->err_flag, 0, 16);
  h->owned.push_back(h->err_flag);
  if (const char* e = getenv("CPT_B200_ATTN")) h->attn_impl = (strcmp(e, "simt") == 0) ? 1 : 0;
  if (const char* e = getenv("CPT_B200_DOWN_KSPLIT")) h->down_ksplit = atoi(e);
  if (const char* e = getenv("CPT_B200_LN_CTAS")) h->ln_ctas_per_sm = std::max(1, atoi(e));
  if (const char* e = getenv("CPT_B200_SMALL_M")) h->small_m_tiles = atoi(e) != 0;
  if (const char* e = getenv("CPT_B200_ATTN_BWD")) h->attn_bwd_simt = strcmp(e, "simt") == 0;
  if (const char* e = getenv("CPT_B200_FOLD_LN")) h->fold_ln = atoi(e) != 0;
  if (const char* e = getenv("CPT_B200_RESID_IN_LN")) h->resid_in_ln = atoi(e) != 0;
  if (const char* e = getenv("CPT_B200_DELTA16")) h->delta16 = atoi(e) != 0;
  if (const char* e = getenv("CPT_B200_TMA_STORE")) h->tma_store = atoi(e) != 0;
  if (const char* e = getenv("CPT_B200_SPLIT")) h->split = atoi(e);
  if (const char* e = getenv("CPT_B200_REDUCE_RESID")) h->reduce_resid = atoi(e) != 0;
  if (const char* e = getenv("CPT_B200_CHAIN")) h->chain = atoi(e) != 0;
  if (const char* e = getenv("CPT_B200_CHAIN_TRACE")) h->chain_trace_on = atoi(e) != 0;
  if (const char* e = getenv("CPT_B200_CHAIN_FUSE_LN")) h->chain_fuse_ln = std::max(0, std::min(2, atoi(e)));
  if (const char* e = getenv("CPT_B200_CHAIN_GROUPS")) h->chain_groups = std::max(1, atoi(e));
  if (const char* e = getenv("CPT_B200_CHAIN_LEAN")) h->chain_lean = atoi(e) != 0;
  if (const char* e = getenv("CPT_B200_CHAIN_MIN_ROWS")) h->chain_min_rows = atoi(e);
  if (const char* e = getenv("CPT_B200_ATTN_EARLY")) h->attn_early = atoi(e);
  if (const char* e = getenv("CPT_B200_CHAIN_EARLY")) h->chain_early = atoi(e);
  if (const char* e = getenv("CPT_B200_CHAIN_KSPLIT")) h->chain_down_ksplit = std::max(1, atoi(e));
  if (getenv("CPT_B200_TRACE")) {
    if (cudaMalloc((void**)&h->trace, (size_t)h->num_sms * 128) == cudaSuccess) {
      cudaMemset(h->trace, 0, (size_t)h->num_sms * 128);
      h->owned.push_back(h->trace);
    }
  }
  // CPT_B200_GEMM="gemm_qkv:256:2,gemm_ffn_down:128:1"  (class:block_n:CTAs per MMA; tuning / A-B experiments)
  if (const char* e = getenv("CPT_B200_GEMM")) {
    std::string str(e);
    size_t pos = 0;
    while (pos < str.size()) {
      size_t end = str.find(',', pos);
      if (end == std::string::npos) end = str.size();
      char name[64];
      int bn = 0, ncta = 0;
      if (sscanf(str.substr(pos, end - pos).c_str(), "%63[^:]:%d:%d", name, &bn, &ncta) == 3)
        for (int t = 0; t < CPT_K_COUNT; ++t)
          if (strcmp(name, cpt_kernel_name(t)) == 0) {
            h->gemm_choice[t].bn = bn;
            h->gemm_choice[t].pair = ncta == 2;
          }
      pos = end + 1;
    }
  }
  *out = h;
  return 0;
}

int cpt_destroy(cpt_handle* h) {
  if (!h) return 0;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  free_owned(h);
  if (h->copy_entries_dev) cudaFree(h->copy_entries_dev);
  if (h->copy_chunks_dev) cudaFree(h->copy_chunks_dev);
  for (void* q : h->owned_chain) cudaFree(q);
  if (h->chain_counters) cudaFree(h->chain_counters);
  if (h->chain_part) cudaFree(h->chain_part);
  if (h->chain_trace) cudaFree(h->chain_trace);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  delete h;
  return 0;
}

int cpt_set_weights(cpt_handle* h, const cpt_weights* w, void* stream) {
  if (!h || !w) return fail("cpt_set_weights: NULL argument");
  DeviceGuard g(h->device);
#define CALL(T) set_weights_impl<T>(h, w, (cudaStream_t)stream)
  return DISPATCH_DTYPE(h, CALL);
#undef CALL
}

size_t cpt_workspace_bytes(const cpt_handle* h, int B, int T, int R) {
  if (!h || B <= 0 || T <= 0 || R < 0) return 0;
  size_t n = carve(h, B, T, R, nullptr).total;
  if (B >= 2) {  // room for the split forward (two half-batch workspaces)
    const size_t two = carve(h, B / 2, T, R, nullptr).total + carve(h, B - B / 2, T, R, nullptr).total + 1024;
    if (two > n) n = two;
  }
  return n;
}

int cpt_encoder_forward(cpt_handle* h, void* stream, const int64_t* input_ids, const int64_t* token_type_ids,
                        const int64_t* attention_mask, const int64_t* position_ids, const float* img_feats, int B,
                        int T, int R, void* workspace, size_t workspace_bytes, float* seq_out, float* pooled,
                        float* hidden_states) {
  if (!h) return fail("NULL handle");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(T16, ST, IDS, SEG, MSK, POS, IMG, NB, WS, WSB, SEQ, POOL)                                            \
  encoder_forward_impl<T16>(h, ST, IDS, SEG, MSK, POS, IMG, NB, T, R, WS, WSB, SEQ, POOL, hidden_states)
  const int S = T + R, H = h->cfg.hidden_size;
  if (h->split == 2 && B >= 16 && !hidden_states && !h->profiling && workspace) {
    // Two half-batches as concurrent branches (fork / join on a side stream; inside a CUDA graph they become parallel
    // branches): every kernel is a persistent grid of <= 148 CTAs, so while one branch's GEMM drains its last tiles
    // the other branch's kernel takes the SMs that are already free.
    if (!h->side) {
      CK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    const int B0 = B / 2, B1 = B - B0;
    const size_t w0 = carve(h, B0, T, R, nullptr).total, w1 = carve(h, B1, T, R, nullptr).total;
    if (workspace_bytes < w0 + w1 + 512) return fail("workspace too small for the split forward");
    char* ws1 = (char*)workspace + ((w0 + 255) & ~size_t(255));
    CK(cudaEventRecord(h->ev_fork, st));
    CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    int rc;
    if (h->cfg.dtype == 0) {
      rc = CALL(__half, st, input_ids, token_type_ids, attention_mask, position_ids, img_feats, B0, workspace, w0,
                seq_out, pooled);
      if (!rc)
        rc = CALL(__half, h->side, input_ids + (size_t)B0 * T, token_type_ids ? token_type_ids + (size_t)B0 * T : nullptr,
                  attention_mask ? attention_mask + (size_t)B0 * S : nullptr,
                  position_ids ? position_ids + (size_t)B0 * T : nullptr,
                  img_feats ? img_feats + (size_t)B0 * R * h->cfg.img_feature_dim : nullptr, B1, ws1, w1,
                  seq_out + (size_t)B0 * S * H, pooled ? pooled + (size_t)B0 * H : nullptr);
    } else {
      rc = CALL(__nv_bfloat16, st, input_ids, token_type_ids, attention_mask, position_ids, img_feats, B0, workspace, w0,
                seq_out, pooled);
      if (!rc)
        rc = CALL(__nv_bfloat16, h->side, input_ids + (size_t)B0 * T,
                  token_type_ids ? token_type_ids + (size_t)B0 * T : nullptr,
                  attention_mask ? attention_mask + (size_t)B0 * S : nullptr,
                  position_ids ? position_ids + (size_t)B0 * T : nullptr,
                  img_feats ? img_feats + (size_t)B0 * R * h->cfg.img_feature_dim : nullptr, B1, ws1, w1,
                  seq_out + (size_t)B0 * S * H, pooled ? pooled + (size_t)B0 * H : nullptr);
    }
    CK(cudaEventRecord(h->ev_join, h->side));
    CK(cudaStreamWaitEvent(st, h->ev_join, 0));
    return rc;
  }
  if (h->cfg.dtype == 0)
    return CALL(__half, st, input_ids, token_type_ids, attention_mask, position_ids, img_feats, B, workspace,
                workspace_bytes, seq_out, pooled);
  return CALL(__nv_bfloat16, st, input_ids, token_type_ids, attention_mask, position_ids, img_feats, B, workspace,
              workspace_bytes, seq_out, pooled);
#undef CALL
}

int cpt_mlm_gather_forward(cpt_handle* h, void* stream, const float* seq_out, int B, int S, const int64_t* mask_pos,
                           const int64_t* vocab_ids, int K, void* workspace, size_t workspace_bytes, float* logits) {
  if (!h) return fail("NULL handle");
  if (!h->has_mlm) return fail("MLM head weights (cls.predictions.*) were not provided");
  const cpt_config& c = h->cfg;
  const int H = c.hidden_size;
  if (!seq_out || !mask_pos || !logits) return fail("NULL argument");
  if (!vocab_ids && K != c.vocab_size) return fail("vocab_ids NULL requires K == vocab_size");
  if ((size_t)B * H * 4 + 256 > workspace_bytes || !workspace) return fail("workspace too small for the MLM head");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  float* t = (float*)(((uintptr_t)workspace + 255) & ~uintptr_t(255));
  // transform: gelu(dense(x[mask_pos]))            (BertPredictionHeadTransform, LN folded into the next kernel)
  TRY(head_matvec(h, st, seq_out, H, S, (const long long*)mask_pos, nullptr, nullptr, 0.f, h->mlm_w, H, h->mlm_b,
                  nullptr, H, B, H, H, ACT_GELU, t, H));
  // decoder over the gathered vocabulary rows (tied to word embeddings) + bias
  TRY(head_matvec(h, st, t, H, 1, nullptr, h->mlm_g, h->mlm_beta, c.layer_norm_eps, h->word, H, h->mlm_bias,
                  (const long long*)vocab_ids, c.vocab_size, B, H, K, ACT_NONE, logits, K));
  return 0;
}

// ------------------------------------------------------------------------------------------------ peer-memory exchange
struct cpt_exchange {
  int device = 0, rank = 0, world = 1, rows = 0, K = 0;
  char* base = nullptr;                 // cudaMalloc: [2][world * rows * K] floats | flags[world] | ctr, ctr2, epoch
  void* peer_base[kMaxPeers] = {};      // cudaIpcOpenMemHandle mappings (peer_base[rank] = base)
  size_t data_bytes = 0;
  bool connected = false;
  HeadPeers peers;
  unsigned* my_flags = nullptr;
  unsigned *ctr = nullptr, *ctr2 = nullptr, *epoch = nullptr;
};

int cpt_exchange_create(int device, int rank, int world, int rows_per_rank, int K, cpt_exchange** out,
                        unsigned char* handle) {
  if (!out || !handle || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || rows_per_rank <= 0 || K <= 0)
    return fail("cpt_exchange_create: bad argument (world <= %d)", kMaxPeers);
  DeviceGuard g(device);
  cpt_exchange* ex = new cpt_exchange();
  ex->device = device;
  ex->rank = rank;
  ex->world = world;
  ex->rows = rows_per_rank;
  ex->K = K;
  ex->data_bytes = (2 * (size_t)world * rows_per_rank * K * sizeof(float) + 255) & ~size_t(255);
  const size_t total = ex->data_bytes + 256;
  cudaError_t e = cudaMalloc((void**)&ex->base, total);
  if (e != cudaSuccess) {
    delete ex;
    return fail("cpt_exchange_create: cudaMalloc(%zu): %s", total, cudaGetErrorString(e));
  }
  cudaMemset(ex->base, 0, total);
  cudaDeviceSynchronize();
  cudaIpcMemHandle_t hnd;
  e = cudaIpcGetMemHandle(&hnd, ex->base);
  if (e != cudaSuccess) {
    cudaFree(ex->base);
    delete ex;
    return fail("cpt_exchange_create: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  static_assert(sizeof(hnd) == CPT_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  memcpy(handle, &hnd, sizeof(hnd));
  *out = ex;
  return 0;
}

int cpt_exchange_connect(cpt_exchange* ex, const unsigned char* handles) {
  if (!ex || !handles) return fail("cpt_exchange_connect: NULL argument");
  if (ex->connected) return fail("cpt_exchange_connect: already connected");
  DeviceGuard g(ex->device);
  for (int r = 0; r < ex->world; ++r) {
    if (r == ex->rank) {
      ex->peer_base[r] = ex->base;
      continue;
    }
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handles + (size_t)r * CPT_IPC_HANDLE_BYTES, sizeof(hnd));
    cudaError_t e = cudaIpcOpenMemHandle(&ex->peer_base[r], hnd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail("cpt_exchange_connect: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
  }
  HeadPeers& p = ex->peers;
  p.world = ex->world;
  p.rank = ex->rank;
  p.half_elems = (long long)ex->world * ex->rows * ex->K;
  p.rank_off = (long long)ex->rank * ex->rows * ex->K;
  for (int r = 0; r < ex->world; ++r) {
    p.dst[r] = (float*)ex->peer_base[r];
    p.flags[r] = (unsigned*)((char*)ex->peer_base[r] + ex->data_bytes);
  }
  ex->my_flags = (unsigned*)(ex->base + ex->data_bytes);
  ex->ctr = ex->my_flags + 32;
  ex->ctr2 = ex->my_flags + 33;
  ex->epoch = ex->my_flags + 34;
  p.ctr = ex->ctr;
  p.epoch = ex->epoch;
  ex->connected = true;
  return 0;
}

int cpt_exchange_destroy(cpt_exchange* ex) {
  if (!ex) return 0;
  DeviceGuard g(ex->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < ex->world; ++r)
    if (r != ex->rank && ex->peer_base[r]) cudaIpcCloseMemHandle(ex->peer_base[r]);
  if (ex->base) cudaFree(ex->base);
  delete ex;
  return 0;
}

static int exchange_wait(cpt_handle* h, cpt_exchange* ex, cudaStream_t st, float* gathered) {
  const long long n = ex->peers.half_elems;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(64, (n + 4095) / 4096));
  ProfScope ps(h, st, CPT_K_HEAD);
  CK(launch_k(exchange_wait_kernel, dim3(blocks), dim3(256), 0, st, 1, ex->peers, (const unsigned*)ex->my_flags, ex->epoch,
              ex->ctr2, gathered, h->err_flag));
  return 0;
}

int cpt_exchange_rows(cpt_handle* h, cpt_exchange* ex, void* stream, const float* local, int rows, float* gathered) {
  if (!h || !ex || !local || !gathered) return fail("cpt_exchange_rows: NULL argument");
  if (!ex->connected) return fail("cpt_exchange_rows: cpt_exchange_connect has not run");
  if (rows != ex->rows) return fail("cpt_exchange_rows: %d rows, the exchange was created for %d per rank", rows, ex->rows);
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)rows * ex->K;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(64, (n + 4095) / 4096));
  {
    ProfScope ps(h, st, CPT_K_HEAD);
    CK(launch_k(exchange_push_kernel, dim3(blocks), dim3(256), 0, st, 1, local, n, ex->peers));
  }
  return exchange_wait(h, ex, st, gathered);
}

int cpt_mlm_gather_exchange(cpt_handle* h, cpt_exchange* ex, void* stream, const float* seq_out, int B, int S,
                            const int64_t* mask_pos, const int64_t* vocab_ids, int K, void* workspace,
                            size_t workspace_bytes, float* gathered) {
  if (!h || !ex) return fail("NULL handle");
  if (!h->has_mlm) return fail("MLM head weights (cls.predictions.*) were not provided");
  if (!ex->connected) return fail("cpt_mlm_gather_exchange: cpt_exchange_connect has not run");
  const cpt_config& c = h->cfg;
  const int H = c.hidden_size;
  if (!seq_out || !mask_pos || !gathered) return fail("NULL argument");
  if (B != ex->rows || K != ex->K) return fail("cpt_mlm_gather_exchange: [%d, %d] logits, the exchange was created for [%d, %d] per rank", B, K, ex->rows, ex->K);
  if (!vocab_ids && K != c.vocab_size) return fail("vocab_ids NULL requires K == vocab_size");
  if ((size_t)B * H * 4 + 256 > workspace_bytes || !workspace) return fail("workspace too small for the MLM head");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  float* t = (float*)(((uintptr_t)workspace + 255) & ~uintptr_t(255));
  TRY(head_matvec(h, st, seq_out, H, S, (const long long*)mask_pos, nullptr, nullptr, 0.f, h->mlm_w, H, h->mlm_b,
                  nullptr, H, B, H, H, ACT_GELU, t, H));
  // the decoder writes its logits into every rank's gather buffer and raises the flags: compute + exchange in one kernel
  TRY(head_matvec(h, st, t, H, 1, nullptr, h->mlm_g, h->mlm_beta, c.layer_norm_eps, h->word, H, h->mlm_bias,
                  (const long long*)vocab_ids, c.vocab_size, B, H, K, ACT_NONE, nullptr, K, ex->peers));
  return exchange_wait(h, ex, st, gathered);
}

size_t cpt_mlm_scores_workspace_bytes(const cpt_handle* h, long long rows) {
  if (!h || rows <= 0) return 0;
  const size_t H = h->cfg.hidden_size;
  return al(rows * H * 2) * 2 + al(rows * H * 4) + 512;
}

int cpt_mlm_scores_forward(cpt_handle* h, void* stream, const float* seq_out, long long rows, void* workspace,
                           size_t workspace_bytes, float* scores) {
  if (!h) return fail("NULL handle");
  if (!h->has_mlm) return fail("MLM head weights (cls.predictions.*) were not provided");
  if (!seq_out || !scores || rows <= 0 || rows > 0x7fffffff) return fail("bad argument");
  DeviceGuard g(h->device);
#define CALL(T16) mlm_scores_impl<T16>(h, (cudaStream_t)stream, seq_out, rows, workspace, workspace_bytes, scores)
  return DISPATCH_DTYPE(h, CALL);
#undef CALL
}

int cpt_nsp_forward(cpt_handle* h, void* stream, const float* pooled, int B, float* out) {
  if (!h) return fail("NULL handle");
  if (!h->has_nsp) return fail("NSP head weights (cls.seq_relationship.*) were not provided");
  if (!pooled || !out) return fail("NULL argument");
  DeviceGuard g(h->device);
  const int H = h->cfg.hidden_size, C = h->cfg.num_contrast_classes;
  return head_matvec(h, (cudaStream_t)stream, pooled, H, 1, nullptr, nullptr, nullptr, 0.f, h->nsp_w, H, h->nsp_b,
                     nullptr, C, B, H, C, ACT_NONE, out, C);
}

int cpt_head_linear(cpt_handle* h, void* stream, const float* x, int B, const float* W, const float* bias, int C,
                    float* out) {
  if (!h) return fail("NULL handle");
  if (!x || !W || !out || B <= 0 || C <= 0) return fail("cpt_head_linear: bad argument");
  DeviceGuard g(h->device);
  const int H = h->cfg.hidden_size;
  return head_matvec(h, (cudaStream_t)stream, x, H, 1, nullptr, nullptr, nullptr, 0.f, W, H, bias, nullptr, C, B, H, C,
                     ACT_NONE, out, C);
}

int cpt_train_enable(cpt_handle* h, int on) {
  if (!h) return fail("NULL handle");
  if ((h->train != 0) != (on != 0)) h->has_weights = false;  // the next cpt_set_weights rebuilds the copies
  h->train = on != 0;
  return 0;
}

int cpt_train_set_progress_callback(cpt_handle* h, cpt_progress_fn fn, void* user) {
  if (!h) return fail("NULL handle");
  h->progress_cb = fn;
  h->progress_user = user;
  return 0;
}

size_t cpt_train_tape_bytes(const cpt_handle* h, int B, int T, int R, int n_rows) {
  if (!h || B <= 0 || T <= 0 || R < 0 || n_rows <= 0) return 0;
  return carve_tape(h, B, T, R, n_rows, nullptr).total;
}

#define CPT_TRAIN_ENTRY(NAME, HEAD)                                                                                   \
  int cpt_train_forward_##NAME(cpt_handle* h, void* stream, const int64_t* input_ids, const int64_t* token_type_ids,  \
                               const int64_t* attention_mask, const int64_t* position_ids, const float* img_feats,    \
                               int B, int T, int R, const int64_t* rows, const int64_t* targets, int n_rows,          \
                               const cpt_dropout* dropout, void* tape, size_t tape_bytes, float* loss) {                                          \
    if (!h) return fail("NULL handle");                                                                               \
    DeviceGuard g(h->device);                                                                                         \
    if (h->cfg.dtype == 0)                                                                                            \
      return train_forward_impl<__half>(h, HEAD, (cudaStream_t)stream, input_ids, token_type_ids, attention_mask,     \
                                        position_ids, img_feats, B, T, R, rows, targets, n_rows, dropout, tape,       \
                                        tape_bytes, loss);                                                                        \
    return train_forward_impl<__nv_bfloat16>(h, HEAD, (cudaStream_t)stream, input_ids, token_type_ids,                \
                                             attention_mask, position_ids, img_feats, B, T, R, rows, targets, n_rows, \
                                             dropout, tape, tape_bytes, loss);                                                 \
  }                                                                                                                   \
  int cpt_train_backward_##NAME(cpt_handle* h, void* stream, const int64_t* input_ids,                                \
                                const int64_t* token_type_ids, const int64_t* position_ids, int B, int T, int R,      \
                                const int64_t* rows, const int64_t* targets, int n_rows, const cpt_dropout* dropout,  \
                                const float* grad_loss, void* tape, size_t tape_bytes, const cpt_grads* grads) {                              \
    if (!h) return fail("NULL handle");                                                                               \
    DeviceGuard g(h->device);                                                                                         \
    if (h->cfg.dtype == 0)                                                                                            \
      return train_backward_impl<__half>(h, HEAD, (cudaStream_t)stream, input_ids, token_type_ids, position_ids, B,   \
                                         T, R, rows, targets, n_rows, dropout, grad_loss, tape, tape_bytes, grads);   \
    return train_backward_impl<__nv_bfloat16>(h, HEAD, (cudaStream_t)stream, input_ids, token_type_ids,               \
                                              position_ids, B, T, R, rows, targets, n_rows, dropout, grad_loss, tape, \
                                              tape_bytes, grads);                                                     \
  }
CPT_TRAIN_ENTRY(mlm, CPT_HEAD_MLM)
CPT_TRAIN_ENTRY(nsp, CPT_HEAD_NSP)
#undef CPT_TRAIN_ENTRY

static_assert(sizeof(cpt_adam_tensor) == sizeof(AdamTensor) && sizeof(cpt_adam_chunk) == sizeof(AdamChunk),
              "optimizer table layouts");
int cpt_adamw_step(int device, void* stream, const cpt_adam_tensor* tensors, const cpt_adam_chunk* chunks, int n_chunks,
                   float beta1, float beta2, float eps, int mode, const float* grad_scale) {
  if (!tensors || !chunks || n_chunks < 0) return fail("cpt_adamw_step: bad argument");
  if (mode != 0 && mode != 1) return fail("cpt_adamw_step: mode must be 0 (torch) or 1 (pytorch-transformers 1.x)");
  if (n_chunks == 0) return 0;
  DeviceGuard g(device);
  adamw_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const AdamTensor*>(tensors),
                                                           reinterpret_cast<const AdamChunk*>(chunks), beta1, beta2, eps,
                                                           mode, grad_scale);
  CKL("adamw_kernel");
  return 0;
}

int cpt_grad_clip_scale(int device, void* stream, const cpt_adam_tensor* tensors, const cpt_adam_chunk* chunks,
                        int n_chunks, float max_norm, const float* grad_scale_in, void* scratch, float* norm_out,
                        float* scale_out) {
  if (!tensors || !chunks || n_chunks <= 0 || !scratch || !norm_out || !scale_out || !(max_norm > 0.f))
    return fail("cpt_grad_clip_scale: bad argument");
  DeviceGuard g(device);
  grad_clip_scale_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const AdamTensor*>(tensors), reinterpret_cast<const AdamChunk*>(chunks), max_norm, grad_scale_in,
      reinterpret_cast<double*>(scratch), norm_out, scale_out);
  CKL("grad_clip_scale_kernel");
  return 0;
}

int cpt_check_async_error(cpt_handle* h, void* stream) {
  if (!h) return fail("NULL handle");
  DeviceGuard g(h->device);
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  int flag = 0;
  CK(cudaMemcpy(&flag, h->err_flag, 4, cudaMemcpyDeviceToHost));
  if (flag) {
    cudaMemset(h->err_flag, 0, 4);
    static const char* what[] = {"", "token / segment / position id out of range", "mask position out of range",
                                 "vocabulary id out of range",
                                 "predicted rectangle with x2 <= x1 or y2 <= y1 (the reference asserts p[2] > p[0])",
                                 "a sample has more region boxes than max_img_seq_len",
                                 "a prompt without [MASK] token (the reference's input_ids.index(103) raises)",
                                 "peer-memory exchange: a rank did not arrive within 10 s (ranks must make the same calls)"};
    return fail("device-side check failed: %s", what[flag < 8 ? flag : 0]);
  }
  return 0;
}

static const char* kKernelNames[CPT_K_COUNT] = {"ext_mask", "embed_text_ln", "cast_pad", "gemm_img", "layernorm",
                                                 "gemm_qkv", "attention", "gemm_attn_out", "gemm_ffn_up",
                                                 "gemm_ffn_down", "head_matvec", "gemm_head", "gemm_other",
                                                 "gemm_dgrad", "gemm_wgrad", "attention_bwd", "train_rowwise",
                                                 "transpose16", "colsum", "layernorm_bwd", "embed_bwd", "chain"};
const char* cpt_kernel_name(int tag) { return (tag >= 0 && tag < CPT_K_COUNT) ? kKernelNames[tag] : ""; }
long long cpt_launch_count(const cpt_handle* h) { return h ? h->launches : 0; }

int cpt_profile_enable(cpt_handle* h, int on) {
  if (!h) return fail("NULL handle");
  DeviceGuard g(h->device);
  CK(cudaDeviceSynchronize());
  for (auto& r : h->prof) { h->ev_pool.push_back(r.a); h->ev_pool.push_back(r.b); }
  h->prof.clear();
  for (int i = 0; i < CPT_K_COUNT; ++i) { h->prof_ms[i] = 0; h->prof_n[i] = 0; }
  h->profiling = on != 0;
  return 0;
}

int cpt_profile_read(cpt_handle* h, double* ms, long long* launches) {
  if (!h || !ms || !launches) return fail("NULL argument");
  DeviceGuard g(h->device);
  CK(cudaDeviceSynchronize());
  for (auto& r : h->prof) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, r.a, r.b));
    h->prof_ms[r.tag] += t;
    h->ev_pool.push_back(r.a);
    h->ev_pool.push_back(r.b);
  }
  h->prof.clear();
  for (int i = 0; i < CPT_K_COUNT; ++i) { ms[i] = h->prof_ms[i]; launches[i] = h->prof_n[i]; h->prof_ms[i] = 0; h->prof_n[i] = 0; }
  return 0;
}

int cpt_gemm_trace(cpt_handle* h, long long* out, int max_ctas) {
  if (!h || !out) return fail("NULL argument");
  if (!h->trace) return fail("tracing is off (set CPT_B200_TRACE=1 before cpt_create)");
  DeviceGuard g(h->device);
  CK(cudaDeviceSynchronize());
  const int n = max_ctas < h->num_sms ? max_ctas : h->num_sms;
  CK(cudaMemcpy(out, h->trace, (size_t)n * 128, cudaMemcpyDeviceToHost));
  return 0;
}

int cpt_gemm(cpt_handle* h, void* stream, const void* A, long long lda, const void* W, long long ldw, int M, int N,
             int K, const float* bias, const float* resid, long long ldr, int epi, int out_fp32, void* out,
             long long ldo, int tile_cfg) {
  if (!h) return fail("NULL handle");
  DeviceGuard g(h->device);
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.out = out; p.ldo = ldo; p.bias = bias; p.resid = resid; p.ldr = ldr;
  p.trans = ((epi & 0x100) ? 3 : 0) | ((epi & 0x400) ? 2 : 0);
  p.tma_reduce = (epi & 0x200) ? 1 : 0;
  p.ksplit = (epi >> 12) & 0xff;
  epi &= 0xff;
#define CALL(T16) gemm<T16>(h, (cudaStream_t)stream, CPT_K_GEMM_OTHER, A, lda, W, ldw, p, epi, out_fp32 != 0, tile_cfg)
  return DISPATCH_DTYPE(h, CALL);
#undef CALL
}

int cpt_assemble_inputs(cpt_handle* h, void* stream, int B, int T, int R, const float* store, const int64_t* feat_row0,
                        const int32_t* n_boxes, const int32_t* tok_a, const int32_t* a_off, const int32_t* tok_b,
                        const int32_t* b_off, const int32_t* has_b, int cls_id, int sep_id, int pad_id, int mask_id,
                        int64_t* input_ids, int64_t* segment_ids, int64_t* input_mask, int64_t* mask_pos,
                        float* img_feats) {
  if (!h) return fail("NULL handle");
  if (B <= 0 || T < 3 || R < 0 || !feat_row0 || !n_boxes || !tok_a || !a_off || !tok_b || !b_off || !has_b || !input_ids ||
      !segment_ids || !input_mask || !mask_pos || (R > 0 && (!store || !img_feats)))
    return fail("cpt_assemble_inputs: bad argument");
  DeviceGuard g(h->device);
  AssembleParams p{B, T, R, h->cfg.img_feature_dim, store, (const long long*)feat_row0, n_boxes, tok_a, a_off, tok_b, b_off,
                   has_b, cls_id, sep_id, pad_id, mask_id, (long long*)input_ids, (long long*)segment_ids,
                   (long long*)input_mask, (long long*)mask_pos, img_feats, h->err_flag};
  ProfScope ps(h, (cudaStream_t)stream, CPT_K_CAST);
  assemble_inputs_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(p);
  CKL("assemble_inputs_kernel");
  return 0;
}

int cpt_score_queries(cpt_handle* h, void* stream, const float* logits, long long ld, int K, int Q,
                      const int32_t* row_start, const int32_t* col_start, const double* rects, const double* gt, int mode,
                      int32_t* pick, double* pick_rect, double* iou, int32_t* correct) {
  if (!h) return fail("NULL handle");
  if (!logits || !row_start || !pick || Q < 0 || K < 2 || ld < K) return fail("cpt_score_queries: bad argument");
  if (mode < 0 || mode > 2) return fail("cpt_score_queries: mode must be 0 (zsl), 1 (fsl) or 2 (vcr)");
  if (Q == 0) return 0;
  DeviceGuard g(h->device);
  ScoreParams p{logits, ld, K, Q, row_start, col_start, rects, gt, mode, pick, pick_rect, iou, correct, h->err_flag};
  ProfScope ps(h, (cudaStream_t)stream, CPT_K_HEAD);
  score_queries_kernel<<<(Q + 3) / 4, 128, 0, (cudaStream_t)stream>>>(p);
  CKL("score_queries_kernel");
  return 0;
}

int cpt_chain_run(cpt_handle* h, void* stream, const cpt_chain_stage* stages, int n_stages) {
  if (!h || !stages || n_stages <= 0 || n_stages > kChainMaxStages) return fail("cpt_chain_run: bad argument");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<ChainStageHost> hs(n_stages);
  for (int i = 0; i < n_stages; ++i) {
    const cpt_chain_stage& s = stages[i];
    ChainStageHost& d = hs[i];
    d.kind = s.kind == 1 ? CHAIN_LN : CHAIN_GEMM;
    d.M = s.M; d.N = s.N; d.K = s.K; d.gelu = s.gelu; d.out_fp32 = s.out_fp32; d.ksplit = s.ksplit;
    d.A = s.A; d.lda = s.lda; d.W = s.W; d.ldw = s.ldw; d.bias = s.bias; d.out = s.out; d.ldo = s.ldo;
    d.ln_in = s.ln_in; d.gamma = s.gamma; d.beta = s.beta; d.eps = s.eps; d.out32 = s.out32; d.out16 = s.out16;
    d.dep_stage = s.dep_stage;
    d.ln = s.ln; d.resid = s.resid; d.ldr = s.ldr;
    d.gvec = s.gvec;
    d.part = reinterpret_cast<float2*>(s.part);
    d.rpart = reinterpret_cast<const float2*>(s.rpart);
    d.apart = reinterpret_cast<const float2*>(s.apart);
  }
  int widest = 128;
  for (auto& d : hs) if (d.ln) widest = std::max(widest, d.N);
  const size_t pneed = chain_part_bytes(hs[0].M, widest);
  if (pneed > h->chain_part_bytes) {
    CK(cudaStreamSynchronize(st));
    if (h->chain_part) cudaFree(h->chain_part);
    h->chain_part = nullptr;
    CK(cudaMalloc((void**)&h->chain_part, pneed));
    h->chain_part_bytes = pneed;
  }
  const size_t need = chain_counter_bytes(hs[0].M, n_stages);
  if (need > h->chain_counters_bytes) {
    CK(cudaStreamSynchronize(st));
    if (h->chain_counters) cudaFree(h->chain_counters);
    h->chain_counters = nullptr;
    CK(cudaMalloc((void**)&h->chain_counters, need));
    h->chain_counters_bytes = need;
  }
  CK(cudaMemsetAsync(h->chain_counters, 0, need, st));
#define CALL(T16) run_chain<T16>(h, st, hs.data(), n_stages, h->chain_counters, h->chain_part)
  return DISPATCH_DTYPE(h, CALL);
#undef CALL
}

int cpt_chain_trace(cpt_handle* h, long long* out, long long max_words, int* pairs, int* pitch) {
  if (!h || !out || !pairs || !pitch) return fail("NULL argument");
  if (!h->chain_trace) return fail("chain tracing is off (set CPT_B200_CHAIN_TRACE=1 before cpt_create) or nothing ran");
  DeviceGuard g(h->device);
  CK(cudaDeviceSynchronize());
  *pairs = h->chain_trace_pairs;
  *pitch = h->chain_trace_pitch;
  const long long words = (long long)h->chain_trace_pairs * (2 + (long long)h->chain_trace_pitch * 16);
  if (words > max_words) return fail("cpt_chain_trace: buffer too small (%lld words needed)", words);
  CK(cudaMemcpy(out, h->chain_trace, (size_t)words * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

int cpt_attention(cpt_handle* h, void* stream, const void* qkv, const float* ext_mask, int B, int S, void* ctx,
                  int impl) {
  if (!h) return fail("NULL handle");
  DeviceGuard g(h->device);
#define CALL(T16) attention<T16>(h, (cudaStream_t)stream, qkv, ext_mask, B, S, ctx, impl)
  return DISPATCH_DTYPE(h, CALL);
#undef CALL
}

int cpt_attention_backward(cpt_handle* h, void* stream, const void* qkv, const void* dctx, const float* ext_mask, int B,
                           int S, void* dqkv, int impl) {
  if (!h || !qkv || !dctx || !ext_mask || !dqkv) return fail("NULL argument");
  DeviceGuard g(h->device);
#define CALL(T) attention_backward<T>(h, (cudaStream_t)stream, qkv, dctx, ext_mask, B, S, dqkv, nullptr, 0.f, 0u, impl)
  return DISPATCH_DTYPE(h, CALL);
#undef CALL
}

int cpt_layernorm(cpt_handle* h, void* stream, const float* x, int M, const float* gamma, const float* beta,
                  float eps, float* out32, void* out16) {
  if (!h) return fail("NULL handle");
  DeviceGuard g(h->device);
  const int H = h->cfg.hidden_size;
#define CALL(T16) layernorm<T16>(h, (cudaStream_t)stream, x, H, M, H, gamma, beta, eps, gamma != nullptr, out32, out16)
  return DISPATCH_DTYPE(h, CALL);
#undef CALL
}

int cpt_cast16(cpt_handle* h, void* stream, const float* x, long long rows, int cols, int ld_out, void* out16) {
  if (!h) return fail("NULL handle");
  DeviceGuard g(h->device);
#define CALL(T16) cast_now<T16>((cudaStream_t)stream, x, rows, cols, ld_out, out16)
  return DISPATCH_DTYPE(h, CALL);
#undef CALL
}

}  // extern "C"
