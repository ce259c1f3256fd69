// tcgen05.cp (shared -> tensor memory) throughput and its overlap with tcgen05.mma, to decide whether
// the score accumulator can be pre-loaded with the int->float bias 0x4B400000 by the tensor pipe
// (so that bits(S) = float(12582912 + s) and the softmax warps need no per-element integer add).
// One CTA per SM, one issuing thread; T "tiles" are issued back to back, one commit at the end.
// Development aid.  Also checks that cp + accumulate-MMA on zero operands leaves the constant.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../quantizedmha_b200/csrc/sm100_ptx.cuh"
using namespace qmha::ptx;

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void cp_32x128b_warpx4(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void cp_128x128b(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}

// mode bits: 1 = cp 32x128b.warpx4 (16 per tile), 2 = cp 128x256b (8 per tile), 4 = cp 128x128b (16 per tile),
//            8 = 4 x i8 MMA 128x64x32 (accumulate onto the tile), 16 = 4 x f16 TS MMA 128x128x16 (P.V)
__global__ void __launch_bounds__(128, 1) k(int mode, int tiles, long long* cyc, uint32_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 16 KB zeros (Q tile, int8 128 x 128)
  uint8_t* sB = smem + 16384;         // 16 KB zeros (K half tile / V^T tile)
  uint32_t* sC = reinterpret_cast<uint32_t*>(smem + 32768);  // 8 KB of the constant
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  for (int i = threadIdx.x; i < 8192 / 4; i += 128) sC[i] = 0x4B400000u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tmem_slot;
  const uint32_t idesc_qk = make_idesc(kAccS32, kFmtS8, kFmtS8, 128, 64);
  const uint32_t idesc_pv = make_idesc(kAccF32, kFmtF16, kFmtF16, 128, 128);
  // zero the P / O columns so that the f16 MMAs read defined data
  {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0;
    const uint32_t la = tb + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 512; c += 32) tmem_st32(la + c, z);
    tmem_wait_st();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint64_t a_desc = make_smem_desc(smem_u32(sA), 128);
    const uint64_t b_desc = make_smem_desc(smem_u32(sB), 128);
    const uint64_t c4 = desc_noswz(smem_u32(sC), 128, 128);    // 32 rows x 16 B: 4 row groups, SBO 128
    const uint64_t c256 = desc_noswz(smem_u32(sC), 128, 256);  // 128 rows x 32 B: LBO 128, SBO 256
    const uint64_t c128 = desc_noswz(smem_u32(sC), 128, 128);  // 128 rows x 16 B
    t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const uint32_t s_tile = tb + (t & 1) * 64;   // two score buffers
      if (mode & 1)
#pragma unroll
        for (int j = 0; j < 16; ++j) cp_32x128b_warpx4(s_tile + j * 4, c4);
      if (mode & 2)
#pragma unroll
        for (int j = 0; j < 8; ++j) cp_128x256b(s_tile + j * 8, c256);
      if (mode & 4)
#pragma unroll
        for (int j = 0; j < 16; ++j) cp_128x128b(s_tile + j * 4, c128);
      if (mode & 8)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_i8_ss(s_tile, advance_smem_desc(a_desc, ks * 32), advance_smem_desc(b_desc, ks * 32), idesc_qk,
                    (mode & 7) ? 1u : (ks > 0));
      if (mode & 16)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_f16_ts(tb + 256, tb + 128 + ks * 8, advance_smem_desc(b_desc, ks * 32), idesc_pv, 1u);
    }
    mma_commit(&bar);
    while (!mbar_try_wait(&bar, 0)) {}
    t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc_fence_after();
  // read back score buffer 0, column `lane` of every warp's lanes: must be the constant (cp modes)
  uint32_t v[32];
  tmem_ld32(tb + ((uint32_t)(warp * 32) << 16), v);
  tmem_wait_ld();
  uint32_t bad = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) bad += (v[i] != 0x4B400000u);
  uint32_t v2[32];
  tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + 32, v2);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) bad += (v2[i] != 0x4B400000u);
  if (blockIdx.x == 0) out[threadIdx.x] = bad;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[128] = v[0];
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* cyc; uint32_t* out;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&out, 129 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  struct { int mode; const char* name; } cases[] = {
      {1, "cp 32x128b.warpx4 x16"}, {2, "cp 128x256b x8"}, {4, "cp 128x128b x16"},
      {8, "4 i8 MMA 128x64x32"}, {16, "4 f16 TS MMA 128x128x16"}, {24, "i8 + f16 MMAs"},
      {9, "cp warpx4 x16 + i8 MMAs"}, {10, "cp 128x256b x8 + i8 MMAs"}, {12, "cp 128x128b x16 + i8 MMAs"},
      {25, "cp warpx4 + i8 + f16"}, {26, "cp 128x256b + i8 + f16"}, {28, "cp 128x128b + i8 + f16"}};
  for (auto& c : cases) {
    for (int grid : {1, 148}) {
      const int tiles = 512;
      k<<<grid, 128, 48 * 1024>>>(c.mode, 8, cyc, out);
      k<<<grid, 128, 48 * 1024>>>(c.mode, tiles, cyc, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%-32s CUDA error: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      long long h[148]; uint32_t ho[129];
      cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, sizeof ho, cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
      uint32_t bad = 0; for (int i = 0; i < 128; ++i) bad += ho[i];
      printf("%-32s grid %3d: %8.1f clk per tile   (constant check: %u mismatches of 8192, word0 = 0x%08x)\n", c.name, grid,
             s / grid / tiles, bad, ho[128]);
    }
  }
  return 0;
}
