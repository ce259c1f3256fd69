// Micro-benchmarks of the SIMT instruction mix of the softmax step on one SM sub-partition:
// how many clocks per warp-level MUFU.EX2 when it is issued alone, together with the conversion /
// FMA / pack / add instructions of the real loop, and together with tcgen05.ld/st.  Development aid.
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

// mode 0: MUFU only.  1: + int->float add + FFMA2.  2: + F2FP pack.  3: + FADD2 row sum (full exp body).
// 4: full body + integer row max of 64 values.  5: mode 4 + tcgen05.ld x64 / st x32 per 64 elements.
template <int kMode>
__global__ void __launch_bounds__(256, 1) mix_kernel(int iters, float c, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (kMode == 5) {
    if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
  }
  const uint32_t base = kMode == 5 ? tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128 : 0;
  uint32_t s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = (uint32_t)((int)((threadIdx.x * 37 + i * 101) % 4001) - 2000);
  if (kMode == 5) { tmem_st32(base, s); tmem_st32(base + 32, s + 32); tmem_wait_st(); }
  uint64_t lsum[2] = {0ull, 0ull};
  uint32_t acc = 0; float facc = 0.f; int macc = 0;
  const uint64_t c2 = pack2(c, c), b2 = pack2(-3.f, -3.f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t p[32];
    if (kMode == 5) { tmem_ld32(base, s); tmem_ld32(base + 32, s + 32); tmem_wait_ld(); }
    if (kMode >= 4) {
      int m0 = max((int)s[0], (int)s[1]), m1 = max((int)s[2], (int)s[3]);
#pragma unroll
      for (int i = 4; i < 64; i += 4) { m0 = max(max(m0, (int)s[i]), (int)s[i + 1]); m1 = max(max(m1, (int)s[i + 2]), (int)s[i + 3]); }
      macc ^= max(m0, m1);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float x0, x1;
      if (kMode >= 1) {
        unpack2(ffma2(pack2(__int_as_float((int)s[2 * i] + 0x4B400000), __int_as_float((int)s[2 * i + 1] + 0x4B400000)), c2, b2), x0, x1);
      } else { x0 = __uint_as_float(s[2 * i]); x1 = __uint_as_float(s[2 * i + 1]); }
      const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
      if (kMode >= 3) lsum[i & 1] = fadd2(lsum[i & 1], pack2(e0, e1));
      if (kMode >= 2) p[i] = pack_f16x2(e0, e1);
      else { facc += e0; facc += e1; }
      if (kMode == 0) { s[2 * i] = __float_as_uint(e0); s[2 * i + 1] = __float_as_uint(e1); }
    }
    if (kMode >= 2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= p[i];
      if (kMode < 5) { s[it & 63] ^= acc & 1; }
    }
    if (kMode == 5) { tmem_st32(base + 64, p); tmem_wait_st(); }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float a, b; unpack2(lsum[0], a, b);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a + b + __uint_as_float(acc) + facc + (float)macc + __uint_as_float(s[5]);
  __syncthreads();
  if (kMode == 5 && warp == 0) tmem_dealloc(tmem_slot, 512);
}

template <int M>
double run(int threads, int iters) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 148 * 256 * sizeof(float));
  mix_kernel<M><<<148, threads>>>(10, 1e-3f, cyc, sink);
  mix_kernel<M><<<148, threads>>>(iters, 1e-3f, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
  cudaFree(cyc); cudaFree(sink);
  return s / 148 / iters;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  const char* names[6] = {"MUFU only", "+VIADD+FFMA2", "+F2FP", "+FADD2 (full exp body)", "+row max", "+tcgen05.ld/st"};
  for (int wps = 1; wps <= 2; ++wps) {
    const int th = 128 * wps;
    double r[6] = {run<0>(th, iters), run<1>(th, iters), run<2>(th, iters), run<3>(th, iters), run<4>(th, iters), run<5>(th, iters)};
    for (int m = 0; m < 6; ++m)
      printf("warps/SMSP=%d %-26s %7.1f clk per 64-element step per warp set  (%.2f clk per warp-level MUFU)\n", wps, names[m], r[m], r[m] / (64.0 * wps));
  }
  return 0;
}
