// Microbenchmark: TMEM -> register read throughput per SM for tcgen05.ld shapes, 4 / 8 warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../cpt_b200/csrc/ptx.cuh"
using namespace cptk;

__device__ __forceinline__ void ld_x64(uint32_t taddr, uint32_t* r) {
  tmem_ld_32x32b_x32(taddr, r);
  tmem_ld_32x32b_x32(taddr + 32, r + 32);
}
__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t* r) {  // 16 lanes x 256 bits x8 = 32 regs
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, long long* out, int nwarps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t row = tb + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    for (int i = 0; i < iters; ++i) {
      uint32_t r[64];
      if (MODE == 0) {  // 8 x (32x32b.x32) per iteration, wait after each
#pragma unroll
        for (int c = 0; c < 8; ++c) { tmem_ld_32x32b_x32(row + c * 32, r); tmem_ld_wait(); acc += r[0] ^ r[31]; }
      } else if (MODE == 1) {  // 4 x (2 x x32 back to back), one wait per pair
#pragma unroll
        for (int c = 0; c < 4; ++c) { ld_x64(row + c * 64, r); tmem_ld_wait(); acc += r[0] ^ r[63]; }
      } else {  // 16x256b.x8: a warp covers 16 lanes x 64 columns... issue twice for 32 lanes
#pragma unroll
        for (int c = 0; c < 8; ++c) { ld_16x256b_x8(row + c * 32, r); tmem_ld_wait(); acc += r[0] ^ r[31]; }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) out[1000] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8192 * 8); long long h[148];
  const int iters = 200;
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 4, 8}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, 256>>>(iters, d, nw);
        if (mode == 1) k<1><<<148, 256>>>(iters, d, nw);
        if (mode == 2) k<2><<<148, 256>>>(iters, d, nw);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
      double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
      // bytes per iteration per warp: 8 chunks x 32 lanes x 32 cols x 4 B = 32 KB  (mode 2: 16 lanes x 8 x 32B... = 16 KB)
      const double bytes = (mode == 2 ? 16384.0 : 32768.0) * nw * iters;
      printf("mode %d warps %d: %.0f cycles, %.1f B/clk/SM\n", mode, nw, cyc, bytes / cyc);
    }
  return 0;
}
