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

}  // namespace cptk
