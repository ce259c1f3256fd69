// Micro-benchmark: SIMT side of one 64-key softmax half-step with the 64 columns of a row SPLIT over
// two warps (4 softmax warps per SM sub-partition, 32 columns each) versus the shipped layout (2 warps
// per sub-partition, 64 columns each).  No MMA / mbarriers.  Development aid only.
//   softmax_split <iters>
#include <cstdio>
#include <cstdlib>
#include "../../quantizedmha_b200/csrc/attn_fwd.cu"

using namespace qmha;
using namespace qmha::ptx;

// kCols = columns per warp per step (64 or 32); blockDim = 128 * warps_per_smsp
template <int kCols, int kMode>
__global__ void __launch_bounds__(512, 1) split_kernel(int iters, float c, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_slot;
  __shared__ float xchg[2][512];
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const int group = warp >> 2;  // 0..3 (or 0..1): which 64/32-column slab of TMEM this warp works in
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + group * 128;
  uint32_t init[32];
  for (int i = 0; i < 32; ++i) init[i] = (uint32_t)((int)((threadIdx.x * 37 + i * 101) % 4001) - 2000);
  tmem_st32(base, init); tmem_st32(base + 32, init); tmem_st32(base + 64, init); tmem_st32(base + 96, init);
  tmem_wait_st();
  __syncthreads();
  float m_used = -INFINITY;
  uint64_t lsum[2] = {0ull, 0ull};
  uint32_t acc = 0;
  constexpr int kPairs = kCols / 2;
  auto rowmax = [&](const uint32_t (&s)[kCols]) {
    int m0 = max((int)s[0], (int)s[1]), m1 = max((int)s[2], (int)s[3]);
#pragma unroll
    for (int i = 4; i < kCols; i += 4) { m0 = max(max(m0, (int)s[i]), (int)s[i + 1]); m1 = max(max(m1, (int)s[i + 2]), (int)s[i + 3]); }
    return (float)max(m0, m1) * c;
  };
  auto expsN = [&](const uint32_t (&s)[kCols], uint32_t (&p)[kPairs], int b, int e) {
    const float bias = -fmaf(kMagicF, c, m_used);
    const uint64_t c2 = pack2(c, c), bias2 = pack2(bias, bias);
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      if (i < b || i >= e) continue;
      float x0, x1;
      unpack2(ffma2(pack2(__int_as_float((int)s[2 * i] + kMagicI2F), __int_as_float((int)s[2 * i + 1] + kMagicI2F)), c2, bias2), x0, x1);
      const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
      lsum[i & 1] = fadd2(lsum[i & 1], pack2(e0, e1));
      p[i] = pack_f16x2(e0, e1);
    }
  };
  auto ld = [&](int it, uint32_t (&dst)[kCols]) {
    const uint32_t a = base + (it & 1) * 64;
    if constexpr (kCols == 64) { tmem_ld32(a, &dst[0]); tmem_ld32(a + 32, &dst[32]); }
    else tmem_ld32(a, &dst[0]);
  };
  auto st = [&](const uint32_t (&p)[kPairs]) {
    if constexpr (kCols == 64) tmem_st32(base + 96, &p[0]);
    else {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::
                   "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]), "r"(p[8]), "r"(p[9]),
                   "r"(p[10]), "r"(p[11]), "r"(p[12]), "r"(p[13]), "r"(p[14]), "r"(p[15]), "r"(base + 96) : "memory");
    }
  };
  auto step = [&](int it, uint32_t (&cur)[kCols], float mt_cur, uint32_t (&nxt)[kCols], float& mt_nxt) {
    const bool need = mt_cur > m_used + kRescaleThreshold;
    if (__any_sync(0xffffffffu, need)) m_used = need ? mt_cur : m_used;
    uint32_t p[kPairs];
    expsN(cur, p, 0, kPairs / 4);
    if (kMode != 1) ld(it + 1, nxt);
    else { _Pragma("unroll") for (int q = 0; q < kCols; ++q) nxt[q] = cur[q] ^ (uint32_t)it; }
    expsN(cur, p, kPairs / 4, (3 * kPairs) / 4);
    tmem_wait_ld();
    float mt = rowmax(nxt);
    if constexpr (kCols == 32) {  // exchange the partial row max with the partner warp (other column half)
      xchg[it & 1][threadIdx.x] = mt;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3) + 4 * ((warp >> 3) & 1)) : "memory");
      mt = fmaxf(mt, xchg[it & 1][threadIdx.x ^ 128]);
    }
    mt_nxt = mt;
    expsN(cur, p, (3 * kPairs) / 4, kPairs);
    if (kMode != 2) { st(p); tmem_wait_st(); }
    acc ^= p[it & (kPairs - 1)];
  };
  uint32_t sA[kCols], sB[kCols];
  float mtA, mtB = 0.f;
  ld(0, sA); tmem_wait_ld();
  mtA = rowmax(sA);
  const long long t0 = clock64();
  for (int it = 0; it < iters; it += 2) {
    step(it, sA, mtA, sB, mtB);
    step(it + 1, sB, mtB, sA, mtA);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float a, b; unpack2(lsum[0], a, b);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a + b + __uint_as_float(acc);
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

template <int kCols, int kMode>
double run(int threads, int iters) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 148 * 512 * sizeof(float));
  split_kernel<kCols, kMode><<<148, threads>>>(10, 1e-3f, cyc, sink);
  split_kernel<kCols, kMode><<<148, threads>>>(iters, 1e-3f, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
  cudaFree(cyc); cudaFree(sink);
  return s / 148 / iters;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  printf("2 warps/SMSP x 64 cols: %.1f clk per step (128 MUFU warp-instr per sub-partition; bound 1024)\n", run<64, 0>(256, iters));
  printf("  same, no tcgen05.ld : %.1f\n", run<64, 1>(256, iters));
  printf("  same, no tcgen05.st : %.1f\n", run<64, 2>(256, iters));
  printf("4 warps/SMSP x 32 cols: %.1f clk per step (same work; partial-max exchange through smem)\n", run<32, 0>(512, iters));
  printf("  same, no tcgen05.ld : %.1f\n", run<32, 1>(512, iters));
  printf("1 warp/SMSP x 64 cols: %.1f ; no ld %.1f\n", run<64, 0>(128, iters), run<64, 1>(128, iters));
  return 0;
}
