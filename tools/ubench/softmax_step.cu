// Micro-benchmark of ONE softmax half-step (tcgen05.ld S -> row max -> exp2 -> pack -> tcgen05.st P)
// without MMA / mbarriers, to measure what the SIMT side alone can sustain per SM.
//   softmax_step <warps_per_smsp(1|2)> <poly_every> <iters>
// Development aid only.
#include <cstdio>
#include <cstdlib>
#include "../../quantizedmha_b200/csrc/attn_fwd.cu"

using namespace qmha;
using namespace qmha::ptx;

template <int kPolyEvery, bool kInt8>
__global__ void __launch_bounds__(256, 1) step_kernel(int iters, float c, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t init[32];
  for (int i = 0; i < 32; ++i) init[i] = kInt8 ? (uint32_t)((int)((threadIdx.x * 37 + i * 101) % 4001) - 2000)
                                               : __float_as_uint(((threadIdx.x * 37 + i * 101) % 4001) * 1e-3f - 2.f);
  tmem_st32(base, init); tmem_st32(base + 32, init); tmem_st32(base + 64, init); tmem_st32(base + 96, init);
  tmem_wait_st();
  __syncthreads();
  float m_used = -INFINITY;
  uint64_t lsum[2] = {0ull, 0ull};
  uint32_t acc = 0;
  const uint32_t pdst = base + 256 - (warp >> 2) * 128 + (warp >> 2) * 32;
  auto fetch = [&](int it, uint32_t (&dst)[kHN]) {
    const int buf = it & 1;
    tmem_ld32(base + buf * 64, &dst[0]);
    tmem_ld32(base + buf * 64 + 32, &dst[32]);
  };
  auto upd = [&](float mt) {
    const bool need = mt > m_used + kRescaleThreshold;
    if (__any_sync(0xffffffffu, need)) m_used = need ? mt : m_used;
  };
  auto pipe_step = [&](int it, uint32_t (&cur)[kHN], float mt_cur, uint32_t (&nxt)[kHN], float& mt_nxt) {
    upd(mt_cur);
    fetch(it + 1, nxt);
    uint32_t p[kHN / 2];
    tile_row_exp<kInt8, false, kPolyEvery, 0, kHN / 4>(cur, p, c, m_used, kHN, lsum);
    tmem_wait_ld();
    mt_nxt = tile_row_max<kInt8, false>(nxt, c, kHN);
    tile_row_exp<kInt8, false, kPolyEvery, kHN / 4, kHN / 2>(cur, p, c, m_used, kHN, lsum);
    tmem_st32(pdst, p);
    tmem_wait_st();
    acc ^= p[it & 31];
  };
  uint32_t sA[kHN], sB[kHN];
  float mtA, mtB = 0.f;
  fetch(0, sA); tmem_wait_ld();
  mtA = tile_row_max<kInt8, false>(sA, c, kHN);
  const long long t0 = clock64();
  for (int it = 0; it < iters; it += 2) {
    pipe_step(it, sA, mtA, sB, mtB);
    pipe_step(it + 1, sB, mtB, sA, mtA);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float a, b; unpack2(lsum[0], a, b);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a + b + __uint_as_float(acc);
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

template <int P, bool I8>
double run(int threads, int iters) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 148 * 256 * sizeof(float));
  step_kernel<P, I8><<<148, threads>>>(10, 1e-3f, cyc, sink);
  step_kernel<P, I8><<<148, threads>>>(iters, 1e-3f, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
  cudaFree(cyc); cudaFree(sink);
  return s / 148 / iters;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  for (int wps = 1; wps <= 2; ++wps) {
    const int threads = 128 * wps;
    printf("warps/SMSP=%d int8 poly0: %.1f clk/half-step\n", wps, run<0, true>(threads, iters));
    printf("warps/SMSP=%d int8 poly4: %.1f clk/half-step\n", wps, run<4, true>(threads, iters));
    printf("warps/SMSP=%d int8 poly2: %.1f clk/half-step\n", wps, run<2, true>(threads, iters));
    printf("warps/SMSP=%d f16  poly0: %.1f clk/half-step\n", wps, run<0, false>(threads, iters));
  }
  return 0;
}
