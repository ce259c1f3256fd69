// Does a TMEM load steal MUFU time?  Two warps per SM sub-partition run the exp2 body of the softmax
// step (64 exponentials per iteration) and additionally issue kLd tcgen05.ld.32x32b.x32 per iteration
// whose results are only waited for at the end of the iteration.  Development aid.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../quantizedmha_b200/csrc/sm100_ptx.cuh"
using namespace qmha::ptx;
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int kLd, int kSt>
__global__ void __launch_bounds__(256, 1) k(int iters, float c, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t s[64], junk[8][32];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = (uint32_t)((int)((threadIdx.x * 37 + i * 101) % 4001) - 2000);
  tmem_st32(base, s); tmem_st32(base + 32, s + 32); tmem_wait_st();
  uint64_t lsum[2] = {0ull, 0ull};
  uint32_t acc = 0;
  const uint64_t c2 = pack2(c, c), b2 = pack2(-3.f, -3.f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t p[32];
#pragma unroll
    for (int l = 0; l < kLd; ++l) tmem_ld32(base + (l & 1) * 32, junk[l]);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float x0, x1;
      unpack2(ffma2(pack2(__int_as_float((int)s[2 * i] + 0x4B400000), __int_as_float((int)s[2 * i + 1] + 0x4B400000)), c2, b2), x0, x1);
      const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
      lsum[i & 1] = fadd2(lsum[i & 1], pack2(e0, e1));
      p[i] = pack_f16x2(e0, e1);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= p[i];
    if (kLd > 0) {
      tmem_wait_ld();
#pragma unroll
      for (int l = 0; l < kLd; ++l) acc ^= junk[l][0] ^ junk[l][31] ^ junk[l][13];
    }
#pragma unroll
    for (int l = 0; l < kSt; ++l) tmem_st32(base + 64 + (l & 1) * 32, p);
    if (kSt > 0) tmem_wait_st();
    s[it & 63] ^= acc & 1;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float a, b; unpack2(lsum[0], a, b);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a + b + __uint_as_float(acc) + __uint_as_float(s[5]);
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
}
template <int L, int S>
double run(int threads, int iters) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 148 * 256 * sizeof(float));
  k<L, S><<<148, threads>>>(10, 1e-3f, cyc, sink);
  k<L, S><<<148, threads>>>(iters, 1e-3f, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
  cudaFree(cyc); cudaFree(sink);
  return s / 148 / iters;
}
int main() {
  const int it = 2000;
  printf("2 warps/SMSP, 64 exp2 per warp per iteration (MUFU bound 1024 clk):\n");
  printf("  no TMEM traffic            %7.1f clk\n", run<0, 0>(256, it));
  printf("  + 1 tcgen05.ld x32 / warp  %7.1f clk\n", run<1, 0>(256, it));
  printf("  + 2 tcgen05.ld x32 / warp  %7.1f clk\n", run<2, 0>(256, it));
  printf("  + 4 tcgen05.ld x32 / warp  %7.1f clk\n", run<4, 0>(256, it));
  printf("  + 8 tcgen05.ld x32 / warp  %7.1f clk\n", run<8, 0>(256, it));
  printf("  + 1 tcgen05.st x32 / warp  %7.1f clk\n", run<0, 1>(256, it));
  printf("  + 4 tcgen05.st x32 / warp  %7.1f clk\n", run<0, 4>(256, it));
  printf("  + 2 ld + 1 st (real step)  %7.1f clk\n", run<2, 1>(256, it));
  return 0;
}
