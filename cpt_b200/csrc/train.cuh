// Kernels of the training step (forward with a tape, backward for every parameter of
// REC_MLM_CPT / NSPCPT — /root/reference/Oscar/oscar/fewshot/{refcoco_cpt.py:231-250, gqa_cpt.py:428-462,
// vcr_nsp_cpt.py:434-473}).  The matrix products (every dgrad / wgrad) are the tcgen05 GEMM of gemm_sm100.cuh reading
// its operands in place through K-major or MN-major descriptors; the attention backward is attention_bwd_sm100.cuh.
// This file holds the row-wise / element-wise pieces (LayerNorm, GELU, embedding and cross-entropy backward, dropout,
// bias column sums) and CUDA-core attention kernels kept as cross-checks (CPT_B200_ATTN_BWD=simt).
#pragma once
#include "ptx.cuh"
#include "rowwise.cuh"

namespace cptk {

// h32 / h16 rows <- dropout(rows) in place (embedding outputs).  Row m of the site lives at stream row remap(m).
template <typename T16>
__global__ void __launch_bounds__(256) dropout_rows_kernel(float* __restrict__ h32, T16* __restrict__ h16, int n_rows,
                                                           int H, int rin, int rout, int roff, Drop d) {
  const long long total = (long long)n_rows * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / H;
    const int c = (int)(i % H);
    const long long row = (m / rin) * rout + roff + (m % rin);
    const long long e = row * H + c;
    const float v = drop_apply(d, (unsigned long long)e, h32[e]);
    h32[e] = v;
    if (h16) h16[e] = Cvt<T16>::from(v);
  }
}
// x = resid + dropout(delta)   (BertSelfOutput / BertOutput: dropout on the dense output, then the residual add)
__global__ void __launch_bounds__(256) dropout_add_kernel(const float* __restrict__ resid, const float* __restrict__ delta,
                                                          long long n, float* __restrict__ x, Drop d) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = resid[i] + drop_apply(d, (unsigned long long)i, delta[i]);
}

// out[C, ld_out] = in[R, ld_in]^T (16-bit), columns R..ld_out-1 of out zero-filled (TMA pitch padding)
template <typename T16>
__global__ void __launch_bounds__(256) transpose16_kernel(const T16* __restrict__ in, int R, int C, long long ld_in,
                                                          T16* __restrict__ out, long long ld_out) {
  __shared__ T16 tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    tile[k][tx] = (r < R && c < C) ? in[(long long)r * ld_in + c] : Cvt<T16>::from(0.f);
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;
    if (c < C && r < ld_out) out[(long long)c * ld_out + r] = tile[tx][k];
  }
}

// out[n] += sum_m in[m, n]   (bias gradients); T = float or a 16-bit type.  A CTA covers 256 columns x kColsumRows
// rows: thread (x, y) sums columns 4x..4x+3 over rows y, y+4, ... with 8 loads in flight, the 4 row-lanes combine in
// shared memory, one atomic per column per CTA.  N % 4 == 0 and 8-byte (16-bit) / 16-byte (fp32) aligned rows take the
// vector path; anything else a scalar path.
constexpr int kColsumRows = 128;
template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 f = *reinterpret_cast<const float4*>(p);
  v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}
template <>
__device__ __forceinline__ void load4<__half>(const __half* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
// Columns [k*seg, (k+1)*seg) go to out_k (the packed [M, 3H] query|key|value gradient feeds three bias tensors).
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ in, int M, int N, long long ld,
                                                     float* __restrict__ out0, float* __restrict__ out1,
                                                     float* __restrict__ out2, int seg, int vec) {
  __shared__ float red[4][256];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int n0 = blockIdx.x * 256 + tx * 4;
  const int m0 = blockIdx.y * kColsumRows, m1 = min(M, m0 + kColsumRows);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (n0 < N) {
    if (vec && n0 + 3 < N) {
      int m = m0 + ty;
      for (; m + 28 < m1; m += 32) {
        float v[8][4];
#pragma unroll
        for (int u = 0; u < 8; ++u) load4<T>(in + (long long)(m + 4 * u) * ld + n0, v[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          acc[0] += v[u][0]; acc[1] += v[u][1]; acc[2] += v[u][2]; acc[3] += v[u][3];
        }
      }
      for (; m < m1; m += 4) {
        float v[4];
        load4<T>(in + (long long)m * ld + n0, v);
        acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3];
      }
    } else {
      for (int m = m0 + ty; m < m1; m += 4)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n0 + e < N) acc[e] += (float)in[(long long)m * ld + n0 + e];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) red[ty][tx * 4 + e] = acc[e];
  __syncthreads();
  const int c = threadIdx.x, n = blockIdx.x * 256 + c;
  if (n < N) {
    const int k = n / seg;
    float* out = k == 0 ? out0 : k == 1 ? out1 : out2;
    atomicAdd(out + (n - k * seg), (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]));
  }
}

__device__ __forceinline__ float gelu_exact(float x) { return x * 0.5f * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
// 8 elements (one 16-byte vector) per thread per step; n % 8 == 0 (intermediate_size is a multiple of 8)
template <typename T16>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    T16 lo, hi;
    *reinterpret_cast<uint16_t*>(&lo) = (uint16_t)(w[i] & 0xffffu);
    *reinterpret_cast<uint16_t*>(&hi) = (uint16_t)(w[i] >> 16);
    f[2 * i] = Cvt<T16>::to(lo);
    f[2 * i + 1] = Cvt<T16>::to(hi);
  }
}
template <typename T16>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = Cvt<T16>::pack2(f[0], f[1]);
  u.y = Cvt<T16>::pack2(f[2], f[3]);
  u.z = Cvt<T16>::pack2(f[4], f[5]);
  u.w = Cvt<T16>::pack2(f[6], f[7]);
  return u;
}
template <typename T16>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const T16* __restrict__ in, long long n, T16* __restrict__ out) {
  const long long nv = n >> 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8<T16>(reinterpret_cast<const uint4*>(in)[i], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = gelu_exact(f[e]);
    reinterpret_cast<uint4*>(out)[i] = pack8<T16>(f);
  }
}
// dpre = dpost * gelu'(pre) over [M, N] (N % 8 == 0, dense rows), and dbias[n] += sum_m dpre[m, n] (the bias gradient of
// intermediate.dense).  A thread keeps one 8-column group and walks `rows_per_cta` rows (chosen by the host so that the
// grid covers the SMs a few times even for few-shot batches: 24 CTAs took 40 us at M = 480).
template <typename T16>
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const T16* __restrict__ dpost, const T16* __restrict__ pre, int M,
                                                       int N, T16* __restrict__ dpre, float* __restrict__ dbias,
                                                       int rows_per_cta) {
  __shared__ float red[128][9];
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;  // 128 column groups x 2 row lanes
  const int cg = blockIdx.x * 128 + tx;                     // 8-column group
  const bool active = cg * 8 < N;
  const int m0 = blockIdx.y * rows_per_cta, m1 = min(M, m0 + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (active) {
#pragma unroll 4
    for (int m = m0 + ty; m < m1; m += 2) {
      const long long i = ((long long)m * N >> 3) + cg;
      float d[8], x[8];
      unpack8<T16>(reinterpret_cast<const uint4*>(dpost)[i], d);
      unpack8<T16>(reinterpret_cast<const uint4*>(pre)[i], x);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        d[e] *= gelu_grad(x[e]);
        acc[e] += d[e];
      }
      reinterpret_cast<uint4*>(dpre)[i] = pack8<T16>(d);
    }
  }
  if (dbias == nullptr) return;
  if (ty == 1)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[tx][e] = acc[e];
  __syncthreads();
  if (ty == 0 && active)
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(dbias + cg * 8 + e, acc[e] + red[tx][e]);
}
// fp32 variants for the small head tensors
__global__ void __launch_bounds__(256) gelu_fwd32_kernel(const float* __restrict__ in, long long n,
                                                         float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = gelu_exact(in[i]);
}
__global__ void __launch_bounds__(256) gelu_bwd32_kernel(const float* __restrict__ dpost, const float* __restrict__ pre,
                                                         long long n, float* __restrict__ dpre) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dpre[i] = dpost[i] * gelu_grad(pre[i]);
}

// LayerNorm backward, one warp per row (persistent over rows).  y = (x - mu) * rstd * gamma + beta.
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;   dgamma += dy * xhat;  dbeta += dy
// dy row index = remap(m) (region rows live behind the text rows of the [B,S,H] stream); x rows are dense.
template <typename T16, int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, int M,
                                                     int H, const float* __restrict__ gamma, float eps, int do_ln,
                                                     float* __restrict__ dx32, T16* __restrict__ dx16,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     float* __restrict__ dbias, int rin, int rout, int roff,
                                                     Drop drop_dy, Drop drop16) {
  // dbias  : += column sums of the (dropout-masked) dx — the bias gradient of the dense layer that fed this LayerNorm.
  // drop_dy: dropout that sat between this LayerNorm's output and its consumer (embedding sites): masks dy on load.
  // drop16 : dropout that sat on the dense output feeding this LayerNorm's input: the 16-bit copy of dx (the operand
  //          of that dense layer's dgrad / wgrad GEMMs) is masked; the fp32 dx (residual branch) is not.
  __shared__ float red[8][NV * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wstride = gridDim.x * 8;
  float4 ag[NV], ab[NV], ax[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = ab[i] = ax[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = blockIdx.x * 8 + warp; row < M; row += wstride) {
    long long drow = row;
    if (rin > 0) drow = (long long)(row / rin) * rout + roff + (row % rin);
    float4 xv[NV], gv[NV], dv[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      xv[i] = *reinterpret_cast<const float4*>(x + (long long)row * H + col);
      dv[i] = *reinterpret_cast<const float4*>(dy + drow * H + col);
      if (drop_dy.thresh) {
        const unsigned long long e = (unsigned long long)(drow * H + col);
        dv[i].x = drop_apply(drop_dy, e, dv[i].x); dv[i].y = drop_apply(drop_dy, e + 1, dv[i].y);
        dv[i].z = drop_apply(drop_dy, e + 2, dv[i].z); dv[i].w = drop_apply(drop_dy, e + 3, dv[i].w);
      }
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
    float mean = 0.f, rstd = 1.f;
    if (do_ln) {
      mean = warp_sum(s) / (float)H;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float a = xv[i].x - mean, b = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
      rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + eps);
    }
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      float4 gm = make_float4(1.f, 1.f, 1.f, 1.f);
      if (do_ln) gm = __ldg(reinterpret_cast<const float4*>(gamma + col));
      // xhat in xv, g = dy * gamma in gv
      xv[i].x = (xv[i].x - mean) * rstd; xv[i].y = (xv[i].y - mean) * rstd;
      xv[i].z = (xv[i].z - mean) * rstd; xv[i].w = (xv[i].w - mean) * rstd;
      gv[i] = make_float4(dv[i].x * gm.x, dv[i].y * gm.y, dv[i].z * gm.z, dv[i].w * gm.w);
      m1 += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
      m2 += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
      ag[i].x += dv[i].x * xv[i].x; ag[i].y += dv[i].y * xv[i].y; ag[i].z += dv[i].z * xv[i].z; ag[i].w += dv[i].w * xv[i].w;
      ab[i].x += dv[i].x; ab[i].y += dv[i].y; ab[i].z += dv[i].z; ab[i].w += dv[i].w;
    }
    m1 = warp_sum(m1) / (float)H;
    m2 = warp_sum(m2) / (float)H;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      float4 o;
      if (do_ln) {
        o.x = rstd * (gv[i].x - m1 - xv[i].x * m2); o.y = rstd * (gv[i].y - m1 - xv[i].y * m2);
        o.z = rstd * (gv[i].z - m1 - xv[i].z * m2); o.w = rstd * (gv[i].w - m1 - xv[i].w * m2);
      } else {
        o = dv[i];
      }
      if (dx32) *reinterpret_cast<float4*>(dx32 + (long long)row * H + col) = o;
      if (drop16.thresh) {
        const unsigned long long e = (unsigned long long)((long long)row * H + col);
        o.x = drop_apply(drop16, e, o.x); o.y = drop_apply(drop16, e + 1, o.y);
        o.z = drop_apply(drop16, e + 2, o.z); o.w = drop_apply(drop16, e + 3, o.w);
      }
      ax[i].x += o.x; ax[i].y += o.y; ax[i].z += o.z; ax[i].w += o.w;
      if (dx16) {
        uint2 u;
        u.x = Cvt<T16>::pack2(o.x, o.y);
        u.y = Cvt<T16>::pack2(o.z, o.w);
        *reinterpret_cast<uint2*>(dx16 + (long long)row * H + col) = u;
      }
    }
  }
  // block reduction of the per-warp partial dgamma / dbeta / dbias, then one atomic per column per CTA
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? dgamma : which == 1 ? dbeta : dbias;
    if (dst == nullptr || (which < 2 && !do_ln)) continue;  // uniform across the CTA
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      *reinterpret_cast<float4*>(&red[warp][col]) = which == 0 ? ag[i] : which == 1 ? ab[i] : ax[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NV * 128; c += 256) {
      float sg = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sg += red[w][c];
      atomicAdd(dst + c, sg);
    }
    __syncthreads();
  }
}

// Text-embedding backward: recompute x = word[id] + pos[p] + type[s], LayerNorm backward, scatter-add dx into the three
// tables' gradients.  dy rows are b*S + t of the [B,S,H] gradient.  Persistent warps: dgamma / dbeta and the (two-row)
// token-type table gradient accumulate per warp / per CTA and reach global memory once per CTA — per-row atomics on
// those few addresses serialised at L2 (251 us at B=64 before); word / position rows take fp32 atomics (rows of the
// same token or position collide only B-fold).
constexpr int kEmbTypeSlots = 4;  // token-type rows accumulated in shared memory (type_vocab_size is 2)
template <int NV>
__global__ void __launch_bounds__(256) embed_bwd_kernel(
    const long long* __restrict__ ids, const long long* __restrict__ seg, const long long* __restrict__ pos_ids,
    const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type,
    const float* __restrict__ gamma, float eps, const float* __restrict__ dy, int B, int T, int S, int H, int vocab,
    int max_pos, int n_type, float* __restrict__ dword, float* __restrict__ dpos, float* __restrict__ dtype,
    float* __restrict__ dgamma, float* __restrict__ dbeta, Drop drop_dy) {
  __shared__ float red[8][NV * 128];
  __shared__ float stype[kEmbTypeSlots][NV * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool type_in_smem = n_type <= kEmbTypeSlots;
  for (int c = threadIdx.x; c < kEmbTypeSlots * NV * 128; c += 256) (&stype[0][0])[c] = 0.f;
  __syncthreads();
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = blockIdx.x * 8 + warp; row < B * T; row += gridDim.x * 8) {
    const int b = row / T, t = row % T;
    const long long id = min(max(ids[row], 0ll), (long long)vocab - 1);
    const long long sg = seg ? min(max(seg[row], 0ll), (long long)n_type - 1) : 0;
    const long long ps = pos_ids ? min(max(pos_ids[row], 0ll), (long long)max_pos - 1) : t;
    float4 xv[NV], dv[NV], gv[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 w = __ldg(reinterpret_cast<const float4*>(word + id * H + col));
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos + ps * H + col));
      const float4 y = __ldg(reinterpret_cast<const float4*>(type + sg * H + col));
      xv[i] = make_float4((w.x + p.x) + y.x, (w.y + p.y) + y.y, (w.z + p.z) + y.z, (w.w + p.w) + y.w);
      dv[i] = *reinterpret_cast<const float4*>(dy + ((long long)b * S + t) * H + col);
      if (drop_dy.thresh) {
        const unsigned long long e = (unsigned long long)(((long long)b * S + t) * H + col);
        dv[i].x = drop_apply(drop_dy, e, dv[i].x); dv[i].y = drop_apply(drop_dy, e + 1, dv[i].y);
        dv[i].z = drop_apply(drop_dy, e + 2, dv[i].z); dv[i].w = drop_apply(drop_dy, e + 3, dv[i].w);
      }
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
    const float mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = xv[i].x - mean, bb = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
      q += (a * a + bb * bb) + (c * c + d * d);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + col));
      xv[i].x = (xv[i].x - mean) * rstd; xv[i].y = (xv[i].y - mean) * rstd;
      xv[i].z = (xv[i].z - mean) * rstd; xv[i].w = (xv[i].w - mean) * rstd;
      gv[i] = make_float4(dv[i].x * gm.x, dv[i].y * gm.y, dv[i].z * gm.z, dv[i].w * gm.w);
      m1 += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
      m2 += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
      ag[i].x += dv[i].x * xv[i].x; ag[i].y += dv[i].y * xv[i].y; ag[i].z += dv[i].z * xv[i].z; ag[i].w += dv[i].w * xv[i].w;
      ab[i].x += dv[i].x; ab[i].y += dv[i].y; ab[i].z += dv[i].z; ab[i].w += dv[i].w;
    }
    m1 = warp_sum(m1) / (float)H;
    m2 = warp_sum(m2) / (float)H;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float o[4] = {rstd * (gv[i].x - m1 - xv[i].x * m2), rstd * (gv[i].y - m1 - xv[i].y * m2),
                          rstd * (gv[i].z - m1 - xv[i].z * m2), rstd * (gv[i].w - m1 - xv[i].w * m2)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (id != 0) atomicAdd(dword + id * H + col + e, o[e]);  // padding_idx = 0 receives no gradient (nn.Embedding)
        atomicAdd(dpos + ps * H + col + e, o[e]);
        if (type_in_smem) atomicAdd(&stype[sg][col + e], o[e]);
        else atomicAdd(dtype + sg * H + col + e, o[e]);
      }
    }
  }
  // per-CTA reduction of dgamma / dbeta (and the token-type rows), one atomic per column per CTA
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      *reinterpret_cast<float4*>(&red[warp][col]) = which == 0 ? ag[i] : ab[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NV * 128; c += 256) {
      float sg = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sg += red[w][c];
      atomicAdd((which == 0 ? dgamma : dbeta) + c, sg);
    }
    __syncthreads();
  }
  if (type_in_smem)
    for (int c = threadIdx.x; c < n_type * NV * 128; c += 256) {
      const float v = stype[c / (NV * 128)][c % (NV * 128)];
      if (v != 0.f) atomicAdd(dtype + (long long)(c / (NV * 128)) * H + c % (NV * 128), v);
    }
}

// X16[i, :] = X32[rows[i], :]  (+ fp32 copy)
template <typename T16>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ X, const long long* __restrict__ rows,
                                                          int n, int H, T16* __restrict__ out16,
                                                          float* __restrict__ out32) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const long long r = rows[i];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float v = X[r * H + c];
    if (out16) out16[(long long)i * H + c] = Cvt<T16>::from(v);
    if (out32) out32[(long long)i * H + c] = v;
  }
}
// dX[rows[i], :] += d[i, :]
__global__ void __launch_bounds__(256) scatter_rows_add_kernel(const float* __restrict__ d,
                                                               const long long* __restrict__ rows, int n, int H,
                                                               float* __restrict__ dX) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const long long r = rows[i];
  for (int c = threadIdx.x; c < H; c += blockDim.x) atomicAdd(dX + r * H + c, d[(long long)i * H + c]);
}

// Cross entropy over the vocabulary at the n labelled rows (CrossEntropyLoss(ignore_index=-1) averages over them,
// modeling_rec.py:148-149).  One CTA per row.  lse[i] saved for the backward.
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ logits, long long ld, int n, int V,
                                                     const long long* __restrict__ targets, float* __restrict__ lse,
                                                     float* __restrict__ loss) {
  __shared__ float sm[8];
  const int i = blockIdx.x;
  const float* row = logits + (long long)i * ld;
  float mx = -INFINITY;
  for (int v = threadIdx.x; v < V; v += 256) mx = fmaxf(mx, row[v]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = sm[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, sm[w]);
  __syncthreads();
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) s += expf(row[v] - mx);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w];
    const float l = mx + logf(t);
    lse[i] = l;
    atomicAdd(loss, (l - row[targets[i]]) / (float)n);
  }
}
// dlogits[i, v] = g * (softmax - onehot) / n   (16-bit, row pitch ld16, zero padded)
template <typename T16>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ logits, long long ld, int n, int V,
                                                     const long long* __restrict__ targets,
                                                     const float* __restrict__ lse, const float* __restrict__ g,
                                                     T16* __restrict__ dlogits, long long ld16) {
  const int i = blockIdx.x;
  const float* row = logits + (long long)i * ld;
  const float scale = g[0] / (float)n, l = lse[i];
  const long long tgt = targets[i];
  for (int v = blockIdx.y * 256 + threadIdx.x; v < ld16; v += gridDim.y * 256) {
    float d = 0.f;
    if (v < V) d = scale * (expf(row[v] - l) - (v == tgt ? 1.f : 0.f));
    dlogits[(long long)i * ld16 + v] = Cvt<T16>::from(d);
  }
}

// fp32 dlogits for the (tiny) NSP head: d[i, c] = g * (softmax - onehot) / n
__global__ void __launch_bounds__(128) ce_bwd32_kernel(const float* __restrict__ logits, int n, int Cn,
                                                       const long long* __restrict__ targets,
                                                       const float* __restrict__ lse, const float* __restrict__ g,
                                                       float* __restrict__ d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * Cn) return;
  const int r = i / Cn, c = i % Cn;
  d[i] = g[0] / (float)n * (expf(logits[i] - lse[r]) - (c == targets[r] ? 1.f : 0.f));
}
// Generic strided fp32 product on CUDA cores for the NSP head (a few hundred rows at most):
//   C[i*sc0 + j] (+)= act'( sum_k A[i*sa0 + k*sa1] * B[k*sb0 + j*sb1] )     act' : 0 none, 1 multiply by (1 - t^2), t = T[i*N+j]
__global__ void __launch_bounds__(256) small_matmul_kernel(const float* __restrict__ A, long long sa0, long long sa1,
                                                           const float* __restrict__ B, long long sb0, long long sb1,
                                                           int M, int N, int K, float* __restrict__ Cm, long long sc0,
                                                           int accumulate, const float* __restrict__ tanh_out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)M * N) return;
  const int i = (int)(idx / N), j = (int)(idx % N);
  float s = 0.f;
  for (int k = 0; k < K; ++k) s = fmaf(A[i * sa0 + k * sa1], B[k * sb0 + j * sb1], s);
  if (tanh_out) {
    const float t = tanh_out[(long long)i * N + j];
    s *= 1.f - t * t;
  }
  if (accumulate) Cm[i * sc0 + j] += s;
  else Cm[i * sc0 + j] = s;
}

// out[i] (+)= a[i]   (fp32, strided rows: copies a padded [rows, ld_a] gradient into its [rows, cols] home)
__global__ void __launch_bounds__(256) add_rows_kernel(const float* __restrict__ a, long long ld_a, int rows, int cols,
                                                       float* __restrict__ out, long long ld_out) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i % cols);
    out[r * ld_out + c] += a[r * ld_a + c];
  }
}
__global__ void __launch_bounds__(256) cast32to16_kernel(const float* __restrict__ in, long long n, __half* o16,
                                                         __nv_bfloat16* ob16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (o16) o16[i] = __float2half_rn(in[i]);
    if (ob16) ob16[i] = __float2bfloat16_rn(in[i]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Attention backward on CUDA cores, one CTA per (head, sample); P is recomputed from Q, K (never stored).
//   phase A (thread = query i): m_i, l_i, D_i = sum_j p_ij (dO_i . v_j), dQ_i = sum_j dS_ij k_j / sqrt(dH)
//   phase B (thread = key j)  : dK_j = sum_i dS_ij q_i / sqrt(dH), dV_j = sum_i p_ij dO_i
// with dS_ij = p_ij (dO_i . v_j - D_i).  No atomics.  Few-shot tuning runs 4-16 rows per GPU; this kernel is not
// meant for large batches.
template <typename T16>
__global__ void __launch_bounds__(128) attn_bwd_simt_kernel(const T16* __restrict__ qkv, const T16* __restrict__ dctx,
                                                            const float* __restrict__ ext_mask, int S, int H,
                                                            float scale, T16* __restrict__ dqkv, Drop drop) {
  // with dropout on the probabilities (P~ = mask * P / (1-p) multiplies V): dV uses P~, and dP = mask/(1-p) * (dO . v)
  extern __shared__ uint8_t smem_raw[];
  T16* sQ = reinterpret_cast<T16*>(smem_raw);
  T16* sK = sQ + (size_t)S * kAttnDH;
  T16* sV = sK + (size_t)S * kAttnDH;
  T16* sdO = sV + (size_t)S * kAttnDH;
  float* sM = reinterpret_cast<float*>(sdO + (size_t)S * kAttnDH);  // mask[S], m[S], l[S], D[S]
  float* sMx = sM + S;
  float* sL = sMx + S;
  float* sD = sL + S;
  const int h = blockIdx.x, b = blockIdx.y;
  const unsigned long long pbase = ((unsigned long long)b * gridDim.x + h) * S * S;  // index of P[b, h, 0, 0]
  const T16* base = qkv + (long long)b * S * 3 * H;
  const T16* dob = dctx + (long long)b * S * H;
  for (int i = threadIdx.x; i < S * 8; i += blockDim.x) {
    const int j = i >> 3, c = (i & 7) * 8;
    *reinterpret_cast<uint4*>(sQ + j * kAttnDH + c) = *reinterpret_cast<const uint4*>(base + (long long)j * 3 * H + h * kAttnDH + c);
    *reinterpret_cast<uint4*>(sK + j * kAttnDH + c) = *reinterpret_cast<const uint4*>(base + (long long)j * 3 * H + H + h * kAttnDH + c);
    *reinterpret_cast<uint4*>(sV + j * kAttnDH + c) = *reinterpret_cast<const uint4*>(base + (long long)j * 3 * H + 2 * H + h * kAttnDH + c);
    *reinterpret_cast<uint4*>(sdO + j * kAttnDH + c) = *reinterpret_cast<const uint4*>(dob + (long long)j * H + h * kAttnDH + c);
  }
  for (int j = threadIdx.x; j < S; j += blockDim.x) sM[j] = ext_mask[(long long)b * S + j];
  __syncthreads();
  T16* dq_out = dqkv + (long long)b * S * 3 * H + h * kAttnDH;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    float q[kAttnDH], dO[kAttnDH];
#pragma unroll
    for (int d = 0; d < kAttnDH; ++d) {
      q[d] = Cvt<T16>::to(sQ[i * kAttnDH + d]);
      dO[d] = Cvt<T16>::to(sdO[i * kAttnDH + d]);
    }
    float mx = -INFINITY;
    for (int j = 0; j < S; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) s = fmaf(q[d], Cvt<T16>::to(sK[j * kAttnDH + d]), s);
      mx = fmaxf(mx, fmaf(s, scale, sM[j]));
    }
    float l = 0.f, D = 0.f;
    for (int j = 0; j < S; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) {
        s = fmaf(q[d], Cvt<T16>::to(sK[j * kAttnDH + d]), s);
        dp = fmaf(dO[d], Cvt<T16>::to(sV[j * kAttnDH + d]), dp);
      }
      const float e = expf(fmaf(s, scale, sM[j]) - mx);
      l += e;
      D += e * drop_apply(drop, pbase + (unsigned long long)i * S + j, dp);
    }
    D /= l;
    sMx[i] = mx; sL[i] = l; sD[i] = D;
    float dq[kAttnDH];
#pragma unroll
    for (int d = 0; d < kAttnDH; ++d) dq[d] = 0.f;
    for (int j = 0; j < S; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) {
        s = fmaf(q[d], Cvt<T16>::to(sK[j * kAttnDH + d]), s);
        dp = fmaf(dO[d], Cvt<T16>::to(sV[j * kAttnDH + d]), dp);
      }
      const float pij = expf(fmaf(s, scale, sM[j]) - mx) / l;
      const float ds = pij * (drop_apply(drop, pbase + (unsigned long long)i * S + j, dp) - D) * scale;
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) dq[d] = fmaf(ds, Cvt<T16>::to(sK[j * kAttnDH + d]), dq[d]);
    }
#pragma unroll
    for (int d = 0; d < kAttnDH; ++d) dq_out[(long long)i * 3 * H + d] = Cvt<T16>::from(dq[d]);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    float k[kAttnDH], v[kAttnDH], dk[kAttnDH], dv[kAttnDH];
#pragma unroll
    for (int d = 0; d < kAttnDH; ++d) {
      k[d] = Cvt<T16>::to(sK[j * kAttnDH + d]);
      v[d] = Cvt<T16>::to(sV[j * kAttnDH + d]);
      dk[d] = 0.f;
      dv[d] = 0.f;
    }
    const float mj = sM[j];
    for (int i = 0; i < S; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) {
        s = fmaf(Cvt<T16>::to(sQ[i * kAttnDH + d]), k[d], s);
        dp = fmaf(Cvt<T16>::to(sdO[i * kAttnDH + d]), v[d], dp);
      }
      const float pij = expf(fmaf(s, scale, mj) - sMx[i]) / sL[i];
      const unsigned long long e = pbase + (unsigned long long)i * S + j;
      const float ds = pij * (drop_apply(drop, e, dp) - sD[i]) * scale;
      const float pdrop = drop_apply(drop, e, pij);
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) {
        dk[d] = fmaf(ds, Cvt<T16>::to(sQ[i * kAttnDH + d]), dk[d]);
        dv[d] = fmaf(pdrop, Cvt<T16>::to(sdO[i * kAttnDH + d]), dv[d]);
      }
    }
#pragma unroll
    for (int d = 0; d < kAttnDH; ++d) {
      dq_out[(long long)j * 3 * H + H + d] = Cvt<T16>::from(dk[d]);
      dq_out[(long long)j * 3 * H + 2 * H + d] = Cvt<T16>::from(dv[d]);
    }
  }
}

// Attention forward with dropout on the probabilities (training with attention_probs_dropout_prob > 0), CUDA cores,
// one CTA per (head, sample): ctx_i = sum_j mask_ij / (1-p) * softmax_j(q_i . k_j / sqrt(dH) + ext_mask_j) v_j.
template <typename T16>
__global__ void __launch_bounds__(128) attn_fwd_drop_kernel(const T16* __restrict__ qkv,
                                                            const float* __restrict__ ext_mask, int S, int H,
                                                            float scale, T16* __restrict__ ctx, Drop drop) {
  extern __shared__ uint8_t smem_raw[];
  T16* sK = reinterpret_cast<T16*>(smem_raw);
  T16* sV = sK + (size_t)S * kAttnDH;
  float* sM = reinterpret_cast<float*>(sV + (size_t)S * kAttnDH);
  const int h = blockIdx.x, b = blockIdx.y;
  const unsigned long long pbase = ((unsigned long long)b * gridDim.x + h) * S * S;
  const T16* base = qkv + (long long)b * S * 3 * H;
  for (int i = threadIdx.x; i < S * 8; i += blockDim.x) {
    const int j = i >> 3, c = (i & 7) * 8;
    *reinterpret_cast<uint4*>(sK + j * kAttnDH + c) =
        *reinterpret_cast<const uint4*>(base + (long long)j * 3 * H + H + h * kAttnDH + c);
    *reinterpret_cast<uint4*>(sV + j * kAttnDH + c) =
        *reinterpret_cast<const uint4*>(base + (long long)j * 3 * H + 2 * H + h * kAttnDH + c);
  }
  for (int j = threadIdx.x; j < S; j += blockDim.x) sM[j] = ext_mask[(long long)b * S + j];
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    float q[kAttnDH], o[kAttnDH];
    const T16* qp = base + (long long)i * 3 * H + h * kAttnDH;
#pragma unroll
    for (int d = 0; d < kAttnDH; ++d) {
      q[d] = Cvt<T16>::to(qp[d]);
      o[d] = 0.f;
    }
    float mx = -INFINITY;
    for (int j = 0; j < S; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) s = fmaf(q[d], Cvt<T16>::to(sK[j * kAttnDH + d]), s);
      mx = fmaxf(mx, fmaf(s, scale, sM[j]));
    }
    float sum = 0.f;
    for (int j = 0; j < S; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) s = fmaf(q[d], Cvt<T16>::to(sK[j * kAttnDH + d]), s);
      const float e = expf(fmaf(s, scale, sM[j]) - mx);
      sum += e;
      const float ed = drop_apply(drop, pbase + (unsigned long long)i * S + j, e);
#pragma unroll
      for (int d = 0; d < kAttnDH; ++d) o[d] = fmaf(ed, Cvt<T16>::to(sV[j * kAttnDH + d]), o[d]);
    }
    const float inv = 1.0f / sum;
    T16* dst = ctx + ((long long)b * S + i) * H + h * kAttnDH;
#pragma unroll
    for (int d = 0; d < kAttnDH; ++d) dst[d] = Cvt<T16>::from(o[d] * inv);
  }
}

}  // namespace cptk
