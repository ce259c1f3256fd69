// tcgen05.mma issue/throughput probes for the attention kernel's shapes: a chain of 4 K-steps into ONE
// accumulator vs the same work interleaved over 2 / 4 independent accumulators, kind::i8 SS (Q.K^T),
// kind::f16 SS (FP16 Q.K^T) and kind::f16 TS (P.V), N = 64 / 128 / 256.  Operands are zero-filled shared
// memory; one issuing thread; clocks per "tile" (4 K-steps of one accumulator).  Development aid.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../quantizedmha_b200/csrc/sm100_ptx.cuh"
using namespace qmha::ptx;

// kind: 0 = i8 SS, 1 = f16 SS, 2 = f16 TS (A from TMEM columns 448..)
__global__ void __launch_bounds__(128, 1) k(int kind, int N, int nacc, int interleave, int tiles, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;            // 16 KB
  uint8_t* sB = smem + 16384;    // 32 KB (256 rows x 128 B)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tmem_slot;
  {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0;
    const uint32_t la = tb + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 512; c += 32) tmem_st32(la + c, z);
    tmem_wait_st();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = kind == 0 ? make_idesc(kAccS32, kFmtS8, kFmtS8, 128, N) : make_idesc(kAccF32, kFmtF16, kFmtF16, 128, N);
    const uint64_t a_desc = make_smem_desc(smem_u32(sA), 128);
    const uint64_t b_desc = make_smem_desc(smem_u32(sB), 128);
    const long long t0 = clock64();
    for (int t = 0; t < tiles; t += nacc) {
      // nacc accumulators of N columns each (N * nacc <= 448); interleave: ks outer, accumulator inner
      if (interleave) {
        for (int ks = 0; ks < 4; ++ks)
          for (int a = 0; a < nacc; ++a) {
            const uint32_t d = tb + a * N;
            if (kind == 0) mma_i8_ss(d, advance_smem_desc(a_desc, ks * 32), advance_smem_desc(b_desc, ks * 32), idesc, ks > 0);
            else if (kind == 1) mma_f16_ss(d, advance_smem_desc(a_desc, ks * 32), advance_smem_desc(b_desc, ks * 32), idesc, ks > 0);
            else mma_f16_ts(d, tb + 448 + ks * 8, advance_smem_desc(b_desc, ks * 32), idesc, ks > 0);
          }
      } else {
        for (int a = 0; a < nacc; ++a)
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t d = tb + a * N;
            if (kind == 0) mma_i8_ss(d, advance_smem_desc(a_desc, ks * 32), advance_smem_desc(b_desc, ks * 32), idesc, ks > 0);
            else if (kind == 1) mma_f16_ss(d, advance_smem_desc(a_desc, ks * 32), advance_smem_desc(b_desc, ks * 32), idesc, ks > 0);
            else mma_f16_ts(d, tb + 448 + ks * 8, advance_smem_desc(b_desc, ks * 32), idesc, ks > 0);
          }
      }
    }
    mma_commit(&bar);
    while (!mbar_try_wait(&bar, 0)) {}
    cyc[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024);
  const char* kn[3] = {"i8 SS ", "f16 SS", "f16 TS"};
  for (int kind = 0; kind < 3; ++kind)
    for (int N : {64, 128, 256})
      for (int nacc : {1, 2, 4}) {
        if (N * nacc > 448) continue;
        for (int il = 0; il < (nacc > 1 ? 2 : 1); ++il) {
          const int tiles = 512;
          k<<<148, 128, 56 * 1024>>>(kind, N, nacc, il, 8, cyc);
          k<<<148, 128, 56 * 1024>>>(kind, N, nacc, il, tiles, cyc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
          long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
          double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
          printf("%s M128 N%-3d K-step 32B x4: %d accumulator(s) %-11s %7.1f clk per tile (floor %d)\n", kn[kind], N, nacc,
                 nacc > 1 ? (il ? "interleaved" : "sequential") : "", s / 148 / tiles, (kind == 0 ? N / 2 : N) * 4 / 2);
        }
      }
  return 0;
}
