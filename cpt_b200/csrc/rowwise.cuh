// HBM/L2-bound row kernels of the CPT path: text-embedding gather + LayerNorm, (residual-)LayerNorm with a
// 16-bit shadow copy for the next GEMM, region-feature cast/pad, and the small fp32 head mat-vecs
// (MLM transform/decoder at the [MASK] rows only, pooler, NSP).  One warp owns one row of H floats
// (H % 128 == 0, H <= 1024), read and written as float4 / 8-byte packed 16-bit.
#pragma once
#include "ptx.cuh"

namespace cptk {

constexpr int kMaxVec = 8;  // H <= 1024

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// BertLayerNorm on a row held as nv float4 per lane: biased variance, eps inside the sqrt, two-pass.
template <typename T16, int NV>
__device__ __forceinline__ void ln_store_row(float4* x, int H, const float* __restrict__ gamma,
                                             const float* __restrict__ beta, float eps, float* __restrict__ out32,
                                             T16* __restrict__ out16, int lane, bool do_ln) {
  float mean = 0.f, rstd = 1.f;
  if (do_ln) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + eps);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    {
      const int col = (i * 32 + lane) * 4;
      float4 y = x[i];
      if (do_ln) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
        y.x = (y.x - mean) * rstd * g.x + b.x;
        y.y = (y.y - mean) * rstd * g.y + b.y;
        y.z = (y.z - mean) * rstd * g.z + b.z;
        y.w = (y.w - mean) * rstd * g.w + b.w;
      }
      if (out32) *reinterpret_cast<float4*>(out32 + col) = y;
      if (out16) {
        uint2 u;
        u.x = Cvt<T16>::pack2(y.x, y.y);
        u.y = Cvt<T16>::pack2(y.z, y.w);
        *reinterpret_cast<uint2*>(out16 + col) = u;
      }
    }
  }
}

// K1: LN(word[ids] + pos[t or position_ids] + type[seg]) -> rows [b*S + t] of the [B,S,H] stream.
template <typename T16, int NV>
__global__ void __launch_bounds__(256) embed_text_ln_kernel(
    const long long* __restrict__ ids, const long long* __restrict__ seg, const long long* __restrict__ pos_ids,
    const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int B, int T, int S, int H,
    int vocab, int max_pos, int n_type, float* __restrict__ out32, T16* __restrict__ out16, int* __restrict__ err) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= B * T) return;
  const int b = row / T, t = row % T;
  long long id = ids[row];
  long long sg = seg ? seg[row] : 0;
  long long ps = pos_ids ? pos_ids[row] : t;
  if (id < 0 || id >= vocab || sg < 0 || sg >= n_type || ps < 0 || ps >= max_pos) {
    if (lane == 0) atomicExch(err, 1);  // surfaced as an error by the host; clamp so we do not fault
    id = min(max(id, 0ll), (long long)vocab - 1);
    sg = min(max(sg, 0ll), (long long)n_type - 1);
    ps = min(max(ps, 0ll), (long long)max_pos - 1);
  }
  float4 x[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 w = __ldg(reinterpret_cast<const float4*>(word + id * H + col));
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos + ps * H + col));
    const float4 y = __ldg(reinterpret_cast<const float4*>(type + sg * H + col));
    x[i] = make_float4((w.x + p.x) + y.x, (w.y + p.y) + y.y, (w.z + p.z) + y.z, (w.w + p.w) + y.w);
  }
  const long long orow = (long long)b * S + t;
  ln_store_row<T16, NV>(x, H, gamma, beta, eps, out32 + orow * H, out16 + orow * H, lane, true);
}

// LayerNorm of fp32 (or 16-bit) rows + optional fp32 residual -> fp32 stream + 16-bit shadow.  Row remap as in
// the region-embedding path (drops the region rows behind the text rows: the `cat` of modeling_bert.py:269 is never
// materialised).  Persistent: a warp walks rows with stride gridDim * warps and has the NEXT row's loads in flight
// while it reduces the current one (ncu: the one-row-per-warp version sat at 34 % of DRAM bandwidth, latency bound).
template <typename T16, int NV>
__device__ __forceinline__ void ln_load_row(float4* x, float4* rs, const float* __restrict__ in,
                                            const T16* __restrict__ in16, long long ld_in,
                                            const float* __restrict__ resid, int H, int row, int lane) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 4;
    if (in16 != nullptr) {  // 16-bit GEMM output (the dense+bias delta); the fp32 residual is added by the caller
      const uint2 u = *reinterpret_cast<const uint2*>(in16 + (long long)row * ld_in + col);
      const T16* e = reinterpret_cast<const T16*>(&u);
      x[i] = make_float4(Cvt<T16>::to(e[0]), Cvt<T16>::to(e[1]), Cvt<T16>::to(e[2]), Cvt<T16>::to(e[3]));
    } else {
      x[i] = *reinterpret_cast<const float4*>(in + (long long)row * ld_in + col);
    }
    if (resid != nullptr) rs[i] = *reinterpret_cast<const float4*>(resid + (long long)row * H + col);
  }
}

// NV = H / 128 float4 per lane (compile-time so the double-buffered row fits the register file at 2 CTAs / SM)
template <typename T16, int NV>
__global__ void __launch_bounds__(256, 2) ln_rows_kernel(const float* __restrict__ in, const T16* __restrict__ in16,
                                                         long long ld_in, const float* __restrict__ resid, int M, int H,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, int do_ln,
                                                         float* __restrict__ out32, T16* __restrict__ out16, int rin,
                                                         int rout, int roff) {
  const int lane = threadIdx.x & 31;
  const int wstride = gridDim.x * (blockDim.x >> 5);
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_launch_dependents();
  pdl_wait();
  float4 x[NV], rs[NV], nx[NV], nrs[NV];
  if (row < M) ln_load_row<T16, NV>(x, rs, in, in16, ld_in, resid, H, row, lane);
  for (; row < M; row += wstride) {
    const int next = row + wstride;
    if (next < M) ln_load_row<T16, NV>(nx, nrs, in, in16, ld_in, resid, H, next, lane);
    if (resid != nullptr) {  // residual added here (a streaming kernel) rather than in the GEMM epilogue
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        x[i].x += rs[i].x; x[i].y += rs[i].y; x[i].z += rs[i].z; x[i].w += rs[i].w;
      }
    }
    long long orow = row;
    if (rin > 0) orow = (long long)(row / rin) * rout + roff + (row % rin);
    ln_store_row<T16, NV>(x, H, gamma, beta, eps, out32 ? out32 + orow * H : nullptr,
                          out16 ? out16 + orow * H : nullptr, lane, do_ln != 0);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      x[i] = nx[i];
      rs[i] = nrs[i];
    }
  }
}

// fp32 [rows, F] -> 16-bit [rows, Fp] (Fp = F rounded up to 8, zero padded) so the row pitch is a legal TMA stride.
template <typename T16>
__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ in, int rows, int F, int Fp,
                                                       T16* __restrict__ out) {
  // one warp per row, 8 independent 8-byte loads in flight per lane (the element-per-thread version was latency bound)
  const int lane = threadIdx.x & 31;
  const int wstride = gridDim.x * (blockDim.x >> 5);
  pdl_launch_dependents();
  pdl_wait();
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += wstride) {
    const float* src = in + (long long)r * F;
    T16* dst = out + (long long)r * Fp;
    const bool al8 = (reinterpret_cast<uintptr_t>(src) & 7) == 0;
    for (int c0 = lane * 2; c0 < Fp; c0 += 64 * 8) {
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * 64;
        v[u] = make_float2(0.f, 0.f);
        if (c + 1 < F) {
          if (al8) v[u] = __ldg(reinterpret_cast<const float2*>(src + c));
          else v[u] = make_float2(__ldg(src + c), __ldg(src + c + 1));
        } else if (c < F) {
          v[u].x = __ldg(src + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * 64;
        if (c < Fp) *reinterpret_cast<uint32_t*>(dst + c) = Cvt<T16>::pack2(v[u].x, v[u].y);
      }
    }
  }
}

// fp32 weight [rows, cols] -> 16-bit [rows, ldo] (zero padded); used once per weight (re)load.
template <typename T16>
__global__ void __launch_bounds__(256) cast_weight_kernel(const float* __restrict__ in, long long rows, int cols,
                                                          int ldo, T16* __restrict__ out) {
  const long long total = rows * ldo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ldo;
    const int c = int(i % ldo);
    out[i] = Cvt<T16>::from(c < cols ? in[r * cols + c] : 0.f);
  }
}

// All weight copies of one cpt_set_weights in ONE launch: entry e converts / copies src[rows, cols] fp32 into
// dst[rows, ldo] (16-bit zero-padded, or fp32 when is16 == 0); a chunk table maps blockIdx.x to (entry, first element).
// A training step refreshes ~180 tensors after every optimizer step; one launch each cost more host time than the
// copies took on the GPU.
struct CopyEntry {
  const float* src;
  void* dst;
  long long rows;
  int cols, ldo, is16, pad;
};
struct CopyChunk {
  int entry, count;
  long long first;
};
constexpr int kCopyChunk = 32768;
template <typename T16>
__global__ void __launch_bounds__(256) refresh_weights_kernel(const CopyEntry* __restrict__ entries,
                                                              const CopyChunk* __restrict__ chunks) {
  const CopyChunk c = chunks[blockIdx.x];
  const CopyEntry e = entries[c.entry];
  for (int i = threadIdx.x; i < c.count; i += 256) {
    const long long k = c.first + i;
    const long long r = k / e.ldo;
    const int col = (int)(k % e.ldo);
    const float v = col < e.cols ? e.src[r * e.cols + col] : 0.f;
    if (e.is16) reinterpret_cast<T16*>(e.dst)[k] = Cvt<T16>::from(v);
    else reinterpret_cast<float*>(e.dst)[k] = v;
  }
}

// LayerNorm folding, weight side (once per weight load): the consumer GEMM of a LayerNorm output
//   LN(x) W^T + b = rstd * (x (gamma .* W)^T - mu * g) + c,   g_n = sum_k fp16(gamma_k W_nk),  c_n = sum_k beta_k W_nk + b_n
// so the GEMM can read the PRE-LayerNorm rows x and finish the normalisation in its epilogue.  g is summed from the
// ROUNDED weights so that the mean term cancels exactly against what the tensor cores accumulated.
template <typename T16>
__global__ void __launch_bounds__(256) fold_weight_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ bias, int N, int K,
                                                          T16* __restrict__ W16, float* __restrict__ g,
                                                          float* __restrict__ c) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float gs = 0.f, cs = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[(long long)n * K + k];
    const T16 h = Cvt<T16>::from(gamma ? w * gamma[k] : w);
    W16[(long long)n * K + k] = h;
    gs += Cvt<T16>::to(h);
    if (beta) cs = fmaf(beta[k], w, cs);
  }
  gs = warp_sum(gs);
  cs = warp_sum(cs);
  if (lane == 0) {
    g[n] = gs;
    c[n] = cs + (bias ? bias[n] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Small fp32 head mat-vec:  Y[b, o] = act( W[wrow(o), :] . LN?(X[xrow(b), :]) + bias[wrow(o)] )
//   xrow(b) = b * x_rows_per_b + (x_pos ? x_pos[b] : 0)        (gather the [MASK] / [CLS] row of sample b)
//   wrow(o) = w_ids ? w_ids[o] : o                              (gather the colour / answer vocabulary rows)
// One CTA handles kHeadRows samples x kHeadOuts outputs (2 per warp); the weight row is read once for all
// kHeadRows samples.
// All fp32: the heads add no rounding beyond the encoder's (BertLMPredictionHead / BertPooler / seq_relationship).
constexpr int kHeadRows = 8;
constexpr int kHeadOuts = 16;
enum HeadAct { ACT_NONE = 0, ACT_GELU = 1, ACT_TANH = 2 };

// Peer-memory exchange of the per-rank logits: every rank's head kernel stores its [rows, K] block
// straight into EVERY rank's gather buffer over NVLink (P2P stores into cudaIpc-mapped memory), then raises one flag
// per peer; a small second kernel waits for all peers' flags and hands the gathered block over.  Epochs count
// exchanges on the device, so the pair of launches can sit in a replayed CUDA graph; two buffer halves alternate by
// epoch parity (a rank cannot get two exchanges ahead of a peer: it needs that peer's flag for the one in between).
constexpr int kMaxPeers = 8;
struct HeadPeers {
  int world = 1, rank = 0;
  long long half_elems = 0;      // floats in one parity half: world * rows_per_rank * K
  long long rank_off = 0;        // rank * rows_per_rank * K
  float* dst[kMaxPeers];         // every rank's buffer (dst[rank] = the local one)
  unsigned* flags[kMaxPeers];    // every rank's flag array [world]
  unsigned* ctr = nullptr;       // local: blocks finished / blocks that read the epoch
  const unsigned* epoch = nullptr;  // local: exchanges completed so far
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// after a block's last peer store: the last block of the grid raises this rank's flag on every peer
__device__ __forceinline__ void peers_signal(const HeadPeers& pr, unsigned n_blocks) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(pr.ctr, 1u) == n_blocks - 1) {
      *pr.ctr = 0;
      const unsigned e = *pr.epoch + 1;
      __threadfence_system();
      for (int r = 0; r < pr.world; ++r) st_release_sys(pr.flags[r] + pr.rank, e);
    }
  }
}

// rows of any producer -> every rank's buffer (NSP scores, or logits computed by another kernel)
__global__ void __launch_bounds__(256) exchange_push_kernel(const float* __restrict__ local, long long n, HeadPeers pr) {
  pdl_launch_dependents();
  pdl_wait();
  const long long off = (long long)((*pr.epoch + 1) & 1u) * pr.half_elems + pr.rank_off;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const float v = local[i];
    for (int r = 0; r < pr.world; ++r) pr.dst[r][off + i] = v;
  }
  peers_signal(pr, gridDim.x);
}

// wait for every peer's flag of this exchange, copy the gathered block out, count the exchange as done
__global__ void __launch_bounds__(256) exchange_wait_kernel(HeadPeers pr, const unsigned* __restrict__ my_flags,
                                                            unsigned* __restrict__ epoch_rw, unsigned* __restrict__ ctr2,
                                                            float* __restrict__ out, int* __restrict__ err) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned e = *pr.epoch + 1;
  if (threadIdx.x < pr.world) {
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(my_flags + threadIdx.x) < e) {
      long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ll) {   // 10 s: a peer never arrived — fail the launch instead of hanging the GPU
        atomicExch(err, 7);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  const float* src = pr.dst[pr.rank] + (long long)(e & 1u) * pr.half_elems;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < pr.half_elems; i += 256ll * gridDim.x) out[i] = __ldcg(src + i);
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(ctr2, 1u) == gridDim.x - 1) {
    *ctr2 = 0;
    *epoch_rw = e;   // every block has read the old value (it arrived at the counter after doing so)
  }
}

__global__ void __launch_bounds__(256) head_matvec_kernel(
    const float* __restrict__ X, long long ldx, int x_rows_per_b, const long long* __restrict__ x_pos,
    const float* __restrict__ ln_g, const float* __restrict__ ln_b, float ln_eps, const float* __restrict__ W,
    long long ldw, const float* __restrict__ bias, const long long* __restrict__ w_ids, int w_rows, int B, int H,
    int O, int act, float* __restrict__ Y, long long ldy, int* __restrict__ err, HeadPeers pr) {
  extern __shared__ float xs[];  // [kHeadRows][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  const int b0 = blockIdx.x * kHeadRows;
  const int nb = min(kHeadRows, B - b0);
  // stage (and optionally normalise) the input rows: warp w owns row w
  if (warp < nb) {
    const int b = b0 + warp;
    long long pos = x_pos ? x_pos[b] : 0;
    if (pos < 0 || pos >= x_rows_per_b) {
      if (lane == 0) atomicExch(err, 2);
      pos = 0;
    }
    const float* xr = X + ((long long)b * x_rows_per_b + pos) * ldx;
    float s = 0.f;
    for (int c = lane; c < H; c += 32) {
      const float v = xr[c];
      xs[warp * H + c] = v;
      s += v;
    }
    if (ln_g != nullptr) {
      const float mean = warp_sum(s) / (float)H;
      float q = 0.f;
      for (int c = lane; c < H; c += 32) {
        const float d = xs[warp * H + c] - mean;
        q += d * d;
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + ln_eps);
      for (int c = lane; c < H; c += 32) xs[warp * H + c] = (xs[warp * H + c] - mean) * rstd * ln_g[c] + ln_b[c];
    }
  } else if (warp < kHeadRows) {
    for (int c = lane; c < H; c += 32) xs[warp * H + c] = 0.f;
  }
  __syncthreads();
  for (int oo = warp; oo < kHeadOuts; oo += 8) {
    const int o = blockIdx.y * kHeadOuts + oo;
    if (o >= O) break;
    long long wr = w_ids ? w_ids[o] : o;
    if (wr < 0 || wr >= w_rows) {
      if (lane == 0) atomicExch(err, 3);
      wr = 0;
    }
    const float* w = W + wr * ldw;
    float acc[kHeadRows];
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) acc[r] = 0.f;
    for (int c = lane * 4; c < H; c += 128) {  // H % 128 == 0
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c));
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + r * H + c);
        acc[r] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[r]))));
      }
    }
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) acc[r] = warp_sum(acc[r]);
    if (lane < nb) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r)
        if (lane == r) v = acc[r];
      if (bias) v += bias[wr];
      if (act == ACT_GELU) v = v * 0.5f * (1.0f + erff(v * 0.70710678118654752f));
      else if (act == ACT_TANH) v = tanhf(v);
      if (pr.world > 1) {   // fused exchange: the value goes straight into every rank's gather buffer
        const long long at = (long long)((*pr.epoch + 1) & 1u) * pr.half_elems + pr.rank_off + (long long)(b0 + lane) * ldy + o;
        for (int r = 0; r < pr.world; ++r) pr.dst[r][at] = v;
      } else {
        Y[(long long)(b0 + lane) * ldy + o] = v;
      }
    }
  }
  if (pr.world > 1) peers_signal(pr, gridDim.x * gridDim.y);
}

}  // namespace cptk
