// Micro-experiment: may a kernel launched as CTA pairs (TMEM allocated with cta_group::2) ALSO issue
// tcgen05.mma.cta_group::1 (each CTA on its own operands / its own TMEM)?  Decides whether the single-CTA attention
// pipeline can become a task type of the pair-based dataflow chain kernel.
//   mode 0: only the cta_group::2 MMA (known-good pattern, validates the harness)
//   mode 1: cta_group::1 MMA in both CTAs, then the cta_group::2 MMA
//   mode 2: cta_group::1 MMA only
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mix_cta_group mix_cta_group.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../cpt_b200/csrc/ptx.cuh"
using namespace cptk;

// A: [128 rows][64 k] fp16 per CTA; B1: [64 n][64 k] per CTA (cta_group::1); B2 half: [32 n][64 k] per CTA (pair)
// swizzled K-major tile: element (r, c) at r*128 + ((c/8) ^ (r&7))*16 + (c%8)*2
__device__ __forceinline__ void put(uint8_t* tile, int r, int c, float v) {
  *reinterpret_cast<__half*>(tile + r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2) = __float2half_rn(v);
}
__host__ __device__ inline float aval(int cta, int r, int c) { return (float)(((r * 7 + c * 3 + cta * 5) % 17) - 8) * 0.125f; }
__host__ __device__ inline float bval(int cta, int n, int c) { return (float)(((n * 5 + c * 11 + cta * 3) % 13) - 6) * 0.25f; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int mode, float* out1, float* out2) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  uint8_t* sA = gen;                 // 16 KB
  uint8_t* sB1 = gen + 16384;        // 8 KB
  uint8_t* sB2 = gen + 16384 + 8192; // 4 KB
  const uint32_t bar1 = base + 32768, bar2 = bar1 + 8, slot = bar1 + 16;
  const uint32_t crank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = (int)crank;
  for (int i = threadIdx.x; i < 128 * 64; i += 128) put(sA, i / 64, i % 64, aval(cta, i / 64, i % 64));
  for (int i = threadIdx.x; i < 64 * 64; i += 128) put(sB1, i / 64, i % 64, bval(cta, i / 64, i % 64));
  // pair MMA: B = 64 rows, CTA r holds rows [32 r, 32 r + 32) (values of "cta 7")
  for (int i = threadIdx.x; i < 32 * 64; i += 128) put(sB2, i / 64, i % 64, bval(7, cta * 32 + i / 64, i % 64));
  if (threadIdx.x == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc_2cta(slot, 256);
    tmem_relinquish_2cta();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tb;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tb) : "r"(slot));
  if (mode != 0 && threadIdx.x == 0) {  // every CTA: its own 128 x 64 x 64 product into its own TMEM columns 0..63
    const uint32_t idesc = make_idesc_f16(128, 64, 0, 0, 0);
    const uint64_t ad = make_smem_desc(base, 16, 1024), bd = make_smem_desc(base + 16384, 16, 1024);
    for (int kk = 0; kk < 4; ++kk) umma_f16(tb, ad + 2 * kk, bd + 2 * kk, idesc, kk != 0);
    umma_commit(bar1);
  }
  if (mode != 0) {
    mbar_wait(bar1, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int c = 0; c < 2; ++c) {
      tmem_ld_32x32b_x32(tb + (uint32_t(warp * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out1[(cta * 128 + warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (mode != 2) {
    if (threadIdx.x == 0 && crank == 0) {  // leader: [A0; A1] (256 rows) x B(64 rows, split over the pair) -> columns 64..127
      const uint32_t idesc = make_idesc_f16(256, 64, 0, 0, 0);
      const uint64_t ad = make_smem_desc(base, 16, 1024), bd = make_smem_desc(base + 16384 + 8192, 16, 1024);
      for (int kk = 0; kk < 4; ++kk) umma_f16_2cta(tb + 64, ad + 2 * kk, bd + 2 * kk, idesc, kk != 0);
      umma_commit_2cta_mc(bar2, 3);
    }
    mbar_wait(bar2, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int c = 0; c < 2; ++c) {
      tmem_ld_32x32b_x32(tb + (uint32_t(warp * 32) << 16) + 64 + c * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out2[(cta * 128 + warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_2cta(tb, 256);
  }
}

int main() {
  float *o1, *o2;
  cudaMalloc(&o1, 256 * 64 * 4);
  cudaMalloc(&o2, 256 * 64 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  std::vector<float> h1(256 * 64), h2(256 * 64);
  auto h = [](float x) { return x; };  // values are exactly representable in fp16
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(o1, 0, 256 * 64 * 4);
    cudaMemset(o2, 0, 256 * 64 * 4);
    k<<<2, 128, 40000>>>(mode, o1, o2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(h1.data(), o1, 256 * 64 * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), o2, 256 * 64 * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int cta = 0; cta < 2; ++cta)
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 64; ++n) {
          double s1 = 0, s2 = 0;
          for (int c = 0; c < 64; ++c) {
            s1 += (double)h(aval(cta, r, c)) * h(bval(cta, n, c));
            s2 += (double)h(aval(cta, r, c)) * h(bval(7, n, c));
          }
          if (mode != 0) e1 = fmax(e1, fabs(s1 - h1[(cta * 128 + r) * 64 + n]));
          if (mode != 2) e2 = fmax(e2, fabs(s2 - h2[(cta * 128 + r) * 64 + n]));
        }
    printf("mode %d: max err cta_group::1 product %.3g, cta_group::2 product %.3g\n", mode, e1, e2);
  }
  return 0;
}
