// Straight-line tcgen05.mma throughput probes (development aid).  Unlike mma_rates.cu the issue loop is
// fully unrolled with the descriptors of every K-step in registers, so the numbers are the tensor pipe's,
// not the issuing thread's.  One CTA per SM; warp 4 lane 0 issues (warp 5 lane 0 is a second issuer for the
// two-issuer patterns); warps 0-3 optionally stream tcgen05.ld/st traffic like the softmax warps do.
// Reports clocks per pattern repetition.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../quantizedmha_b200/csrc/sm100_ptx.cuh"
using namespace qmha::ptx;

enum Pat {
  kI8ss64 = 0, kI8ss64acc, kI8ss128, kI8ss256, kF16ts128, kPvQk, kPvQk2tiles, kI8ts64, kI8ts128, kF16ss64x8,
  kF16ss128x8, kI8ss64x2tiles, kPvQk128, kF16ts128x8, kI8ss64one, kF16ts256, kCrossSeq, kCrossIl, kF16CrossSeq, kF16CrossIl, kF16SameSeq, kNumPat
};
static const char* kNames[kNumPat] = {
    "i8 SS N64 x4 (K=128)", "i8 SS N64 x4 acc=1", "i8 SS N128 x4", "i8 SS N256 x4", "f16 TS N128 K16 x4 (P.V 64 keys)",
    "PV(f16 TS N128 x4) + QK(i8 SS N64 x4)", "same, tile0 + tile1 interleaved per op", "i8 TS N64 x4", "i8 TS N128 x4",
    "f16 SS N64 K16 x8", "f16 SS N128 K16 x8", "i8 SS N64 x4, tile0 then tile1", "PV(f16 TS N128 x8) + QK(i8 SS N128 x4)",
    "f16 TS N128 K16 x8 (P.V 128 keys)", "i8 SS N64 x1", "f16 TS N256 K16 x4",
    "INT8 kernel: PV(t0) x4 then QK(t1) x4", "INT8 kernel: PV(t0) / QK(t1) interleaved 1:1",
    "FP16 kernel: PV(t0) x4 then QK(t1) f16 SS N64 x8", "FP16 kernel: PV(t0) / QK(t1) interleaved 1:2",
    "FP16 kernel: PV(t0) x4 then QK(t0) f16 SS N64 x8"};

// TMEM: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); Q in TMEM (TS int8): columns 192.. (overlaps S1, timing only)
template <int PAT>
__device__ __forceinline__ void pattern(uint32_t tb, const uint64_t (&qa)[2][8], const uint64_t (&kb)[8],
                                        const uint64_t (&vb)[8], int rep) {
  constexpr uint32_t i8_64 = make_idesc(kAccS32, kFmtS8, kFmtS8, 128, 64);
  constexpr uint32_t i8_128 = make_idesc(kAccS32, kFmtS8, kFmtS8, 128, 128);
  constexpr uint32_t i8_256 = make_idesc(kAccS32, kFmtS8, kFmtS8, 128, 256);
  constexpr uint32_t f16_64 = make_idesc(kAccF32, kFmtF16, kFmtF16, 128, 64);
  constexpr uint32_t f16_128 = make_idesc(kAccF32, kFmtF16, kFmtF16, 128, 128);
  constexpr uint32_t f16_256 = make_idesc(kAccF32, kFmtF16, kFmtF16, 128, 256);
  const uint32_t sbuf = (rep & 1) * 64;
  auto qk64 = [&](int t, bool acc0) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_i8_ss(tb + t * 128 + sbuf, qa[t][ks], kb[ks], i8_64, (acc0 || ks > 0) ? 1u : 0u);
  };
  auto pv64 = [&](int t) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_f16_ts(tb + 256 + t * 128, tb + t * 128 + (64 - sbuf) + ks * 8, vb[ks], f16_128, 1u);
  };
  if constexpr (PAT == kI8ss64) qk64(0, false);
  if constexpr (PAT == kI8ss64acc) qk64(0, true);
  if constexpr (PAT == kI8ss64one) mma_i8_ss(tb + sbuf, qa[0][0], kb[0], i8_64, 1u);
  if constexpr (PAT == kI8ss128) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_i8_ss(tb, qa[0][ks], kb[ks], i8_128, ks > 0);
  }
  if constexpr (PAT == kI8ss256) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_i8_ss(tb, qa[0][ks], kb[ks], i8_256, ks > 0);
  }
  if constexpr (PAT == kF16ts128) pv64(0);
  if constexpr (PAT == kF16ts256) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_f16_ts(tb + 256, tb + ks * 8, vb[ks], f16_256, 1u);
  }
  if constexpr (PAT == kF16ts128x8) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) mma_f16_ts(tb + 256, tb + ks * 8, vb[ks], f16_128, 1u);
  }
  if constexpr (PAT == kPvQk) { pv64(0); qk64(0, false); }
  if constexpr (PAT == kPvQk2tiles) { pv64(0); qk64(0, false); pv64(1); qk64(1, false); }
  if constexpr (PAT == kI8ss64x2tiles) { qk64(0, false); qk64(1, false); }
  if constexpr (PAT == kCrossSeq) { pv64(0); qk64(1, false); }
  if constexpr (PAT == kCrossIl) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      mma_f16_ts(tb + 256, tb + (64 - sbuf) + ks * 8, vb[ks], f16_128, 1u);
      mma_i8_ss(tb + 128 + sbuf, qa[1][ks], kb[ks], i8_64, ks > 0);
    }
  }
  if constexpr (PAT == kF16CrossSeq || PAT == kF16SameSeq) {
    pv64(0);
    constexpr int t = PAT == kF16CrossSeq ? 1 : 0;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) mma_f16_ss(tb + t * 128 + sbuf, qa[t][ks], kb[ks], f16_64, ks > 0);
  }
  if constexpr (PAT == kF16CrossIl) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      mma_f16_ts(tb + 256, tb + (64 - sbuf) + ks * 8, vb[ks], f16_128, 1u);
      mma_f16_ss(tb + 128 + sbuf, qa[1][2 * ks], kb[2 * ks], f16_64, ks > 0);
      mma_f16_ss(tb + 128 + sbuf, qa[1][2 * ks + 1], kb[2 * ks + 1], f16_64, 1u);
    }
  }
  if constexpr (PAT == kI8ts64) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_i8_ts(tb + sbuf, tb + 192 + ks * 8, kb[ks], i8_64, ks > 0);
  }
  if constexpr (PAT == kI8ts128) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_i8_ts(tb, tb + 192 + ks * 8, kb[ks], i8_128, ks > 0);
  }
  if constexpr (PAT == kF16ss64x8) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) mma_f16_ss(tb + sbuf, qa[0][ks], kb[ks], f16_64, ks > 0);
  }
  if constexpr (PAT == kF16ss128x8) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) mma_f16_ss(tb, qa[0][ks], kb[ks], f16_128, ks > 0);
  }
  if constexpr (PAT == kPvQk128) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) mma_f16_ts(tb + 256, tb + ks * 8, vb[ks], f16_128, 1u);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_i8_ss(tb, qa[0][ks], kb[ks], i8_128, ks > 0);
  }
}

// issuers: 1 = warp 4 only; 2 = warps 4 and 5 issue the same pattern on tile 0 / tile 1 resources concurrently
template <int PAT>
__global__ void __launch_bounds__(192, 1) k(int reps, int issuers, int traffic, long long* cyc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                  // 2 x 32 KB (two query tiles, up to 256 B rows as two 128 B sub-tiles)
  uint8_t* sK = smem + 65536;          // 64 KB: 256 rows x 128 B x 2 sub-tiles
  uint8_t* sV = smem + 131072;         // 64 KB
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 196608 / 4; i += 192) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); stop = 0; }
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (warp < 4) {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0;
    const uint32_t la = tb + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 512; c += 32) tmem_st32(la + c, z);
    tmem_wait_st();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp >= 4 && lane == 0 && warp - 4 < issuers) {
    const int me = warp - 4;
    uint64_t qa[2][8], kb[8], vb[8];
    const uint64_t q0 = make_smem_desc(smem_u32(sQ), 128), k0 = make_smem_desc(smem_u32(sK), 128),
                   v0 = make_smem_desc(smem_u32(sV), 128);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t in_atom = (uint32_t)(ks % 4) * 32, sub = (uint32_t)(ks / 4);
      qa[0][ks] = advance_smem_desc(q0, (me ? 32768u : 0u) + sub * 16384u + in_atom);
      qa[1][ks] = advance_smem_desc(q0, (me ? 0u : 32768u) + sub * 16384u + in_atom);
      kb[ks] = advance_smem_desc(k0, sub * 32768u + in_atom);
      vb[ks] = advance_smem_desc(v0, sub * 32768u + in_atom);
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; r += 2) {
      pattern<PAT>(me ? tb + 128 : tb, qa, kb, vb, 0);
      pattern<PAT>(me ? tb + 128 : tb, qa, kb, vb, 1);
    }
    mma_commit(&bar[me]);
    while (!mbar_try_wait(&bar[me], 0)) {}
    cyc[blockIdx.x * 2 + me] = clock64() - t0;
    if (me == 0) stop = 1;
  } else if (warp < 4 && traffic) {
    // softmax-like TMEM traffic: per iteration two x32 loads of S columns and one x32 store (per warp)
    uint32_t v[32], w[32];
    const uint32_t la = tb + ((uint32_t)(warp * 32) << 16);
    uint32_t acc = 0;
    while (!stop) {
      tmem_ld32(la + 448, v);
      tmem_ld32(la + 480, w);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += v[i] ^ w[i];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = acc & 0;   // keep zeros
      tmem_st32(la + 448, v);
      tmem_wait_st();
    }
    if (acc == 0x12345678u) cyc[0] = 0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int PAT>
void run(long long* cyc) {
  cudaFuncSetAttribute(k<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 193 * 1024 + 1024);
  constexpr bool two_ok = PAT != kPvQk2tiles && PAT != kI8ss64x2tiles && PAT != kI8ss256 && PAT != kF16ts256 && PAT < kCrossSeq;
  for (int issuers : {1, 2}) {
    if (issuers == 2 && !two_ok) continue;
    for (int traffic : {0, 1}) {
      const int reps = 512;
      k<PAT><<<148, 192, 193 * 1024 + 1024>>>(8, issuers, traffic, cyc);
      k<PAT><<<148, 192, 193 * 1024 + 1024>>>(reps, issuers, traffic, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error: %s\n", kNames[PAT], cudaGetErrorString(e)); exit(1); }
      long long h[296]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      double s = 0; int n = 0;
      for (int i = 0; i < 148; ++i) for (int m = 0; m < issuers; ++m) { s += h[i * 2 + m]; ++n; }
      printf("%-46s issuers %d  tmem traffic %d: %8.1f clk per repetition%s\n", kNames[PAT], issuers, traffic, s / n / reps,
             issuers == 2 ? " (each issuer)" : "");
    }
  }
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 296 * 8);
  run<kI8ss64one>(cyc); run<kI8ss64>(cyc); run<kI8ss64acc>(cyc); run<kI8ss128>(cyc); run<kI8ss256>(cyc);
  run<kI8ts64>(cyc); run<kI8ts128>(cyc);
  run<kF16ts128>(cyc); run<kF16ts128x8>(cyc); run<kF16ts256>(cyc); run<kF16ss64x8>(cyc); run<kF16ss128x8>(cyc);
  run<kI8ss64x2tiles>(cyc); run<kPvQk>(cyc); run<kPvQk2tiles>(cyc); run<kPvQk128>(cyc);
  run<kCrossSeq>(cyc); run<kCrossIl>(cyc); run<kF16SameSeq>(cyc); run<kF16CrossSeq>(cyc); run<kF16CrossIl>(cyc);
  return 0;
}
