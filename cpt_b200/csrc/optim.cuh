// Multi-tensor AdamW: ONE launch updates every parameter of every group (the reference's optimizers are a Python loop of
// ~10 small torch ops per parameter tensor — pytorch-transformers 1.x `AdamW.step`, used by fewshot/gqa_cpt.py:342 and
// vcr_nsp_cpt.py — or torch.optim.AdamW, fewshot/refcoco_cpt.py:342).  Work is cut into chunks of kAdamChunk elements;
// a chunk table maps blockIdx.x to (tensor, offset).  Both tables live in device memory and are rebuilt by the host
// binding whenever gradient pointers change.
#pragma once
#include <cstdint>

namespace cptk {

struct AdamTensor {  // mirrors cpt_adam_tensor (include/cpt_b200.h)
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  float lr, wd, bc1, bc2;  // group hyper-parameters and this tensor's bias corrections 1 - beta^t (1 when disabled)
};
struct AdamChunk {  // mirrors cpt_adam_chunk
  int tensor, count;
  long long offset;
};
constexpr int kAdamChunk = 16384;

// mode 0: torch.optim.AdamW   p *= 1 - lr*wd;  m, v;  p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// mode 1: pytorch-transformers 1.x AdamW   m, v;  p -= lr*sqrt(bc2)/bc1 * m / (sqrt(v) + eps);  p -= lr*wd*p
__global__ void __launch_bounds__(256) adamw_kernel(const AdamTensor* __restrict__ tensors,
                                                    const AdamChunk* __restrict__ chunks, float beta1, float beta2,
                                                    float eps, int mode, const float* __restrict__ grad_scale) {
  const AdamChunk c = chunks[blockIdx.x];
  const AdamTensor t = tensors[c.tensor];
  const float gs = grad_scale ? *grad_scale : 1.f;
  const float sq2 = sqrtf(t.bc2);
  for (int i = threadIdx.x; i < c.count; i += 256) {
    const long long k = c.offset + i;
    const float g = t.g[k] * gs;
    float p = t.p[k];
    const float m = beta1 * t.m[k] + (1.f - beta1) * g;
    const float v = beta2 * t.v[k] + (1.f - beta2) * g * g;
    t.m[k] = m;
    t.v[k] = v;
    if (mode == 0) {
      p *= 1.f - t.lr * t.wd;
      p -= (t.lr / t.bc1) * (m / (sqrtf(v) / sq2 + eps));
    } else {
      p -= (t.lr * sq2 / t.bc1) * (m / (sqrtf(v) + eps));
      if (t.wd > 0.f) p -= t.lr * t.wd * p;
    }
    t.p[k] = p;
  }
}

// Global gradient norm + clipping coefficient of torch.nn.utils.clip_grad_norm_(params, max_norm) (the GQA / VCR few-shot
// loops call it right before optimizer.step(), Oscar/oscar/fewshot/gqa_cpt.py:454) WITHOUT touching the gradients: one
// launch over the same (tensor, chunk) tables as the update reduces sum g^2 (per block in fp32 pairs, across blocks in
// double precision), the last block to finish writes
//     norm = sqrt(sum (gs g)^2),   scale = gs * min(1, max_norm / (norm + 1e-6))
// and clears the scratch for the next step; adamw_kernel then multiplies every gradient by `scale` as it reads it.
// scratch: 16 bytes of zeroed device memory {double sum; unsigned done; unsigned pad}.
__global__ void __launch_bounds__(256) grad_clip_scale_kernel(const AdamTensor* __restrict__ tensors,
                                                              const AdamChunk* __restrict__ chunks, float max_norm,
                                                              const float* __restrict__ grad_scale_in,
                                                              double* __restrict__ scratch, float* __restrict__ norm_out,
                                                              float* __restrict__ scale_out) {
  const AdamChunk c = chunks[blockIdx.x];
  const float* g = tensors[c.tensor].g + c.offset;
  float s0 = 0.f, s1 = 0.f;
  for (int i = threadIdx.x; i < c.count; i += 512) {
    const float a = g[i];
    s0 = fmaf(a, a, s0);
    if (i + 256 < c.count) {
      const float b = g[i + 256];
      s1 = fmaf(b, b, s1);
    }
  }
  double s = (double)s0 + (double)s1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double ws[8];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    atomicAdd(scratch, t);
    __threadfence();
    unsigned* done = reinterpret_cast<unsigned*>(scratch + 1);
    last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const double total = *reinterpret_cast<volatile double*>(scratch);
    const float gs = grad_scale_in ? *grad_scale_in : 1.f;
    const float norm = (float)sqrt(total) * fabsf(gs);
    const float coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
    *norm_out = norm;
    *scale_out = gs * coef;
    *scratch = 0.0;
    *reinterpret_cast<unsigned*>(scratch + 1) = 0u;
  }
}

}  // namespace cptk
