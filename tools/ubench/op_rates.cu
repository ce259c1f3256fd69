// Throughput of the individual SASS instructions of the softmax loop on one SM sub-partition
// (clocks per warp-level instruction, 1 / 2 / 4 warps per sub-partition, 8 independent chains).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
enum Op { VIMNMX3, VIMNMX2, FMNMX3, FMNMX2, VIADD_, LOP3_, FFMA2_, FADD2_, F2FP_, MUFU_, I2FP_, HMNMX2_, MUFU16_, MUFU16X2_, HADD2_, NOPS };
template <int OP>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cyc, int* sink, int seed) {
  int a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 7 + i + seed; b[i] = threadIdx.x * 3 + i * 5 + seed; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == VIMNMX3) a[i] = max(max(a[i], b[i]), b[(i + 1) & 7] + u);
        if (OP == VIMNMX2) a[i] = max(a[i], b[i]) ^ b[(i + 1) & 7];
        if (OP == FMNMX3) a[i] = __float_as_int(fmaxf(fmaxf(__int_as_float(a[i]), __int_as_float(b[i])), __int_as_float(b[(i + 1) & 7])));
        if (OP == FMNMX2) a[i] = __float_as_int(fmaxf(__int_as_float(a[i]), __int_as_float(b[i])) + __int_as_float(b[(i + 1) & 7]));
        if (OP == VIADD_) a[i] = a[i] + b[i];
        if (OP == LOP3_) a[i] = (a[i] ^ b[i]) | b[(i + 3) & 7];
        if (OP == FFMA2_ || OP == FADD2_) {
          if ((i & 1) == 0) {
            uint64_t x, y, z;
            asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(a[i]), "r"(a[i + 1]));
            asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "r"(b[i]), "r"(b[i + 1]));
            if (OP == FFMA2_) asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(z) : "l"(x), "l"(y));
            else asm("add.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(x), "l"(y));
            asm("mov.b64 {%0, %1}, %2;" : "=r"(a[i]), "=r"(a[i + 1]) : "l"(z));
          }
        }
        if (OP == F2FP_) { uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__int_as_float(a[i])), "f"(__int_as_float(b[i]))); a[i] = r; }
        if (OP == MUFU_) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__int_as_float(a[i]))); a[i] = __float_as_int(y); }
        if (OP == I2FP_) a[i] = __float_as_int((float)a[i]);
        if (OP == MUFU16_) { unsigned short h = (unsigned short)a[i], r; asm("ex2.approx.f16 %0, %1;" : "=h"(r) : "h"(h)); a[i] = r; }
        if (OP == HADD2_) { uint32_t r; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b[i])); a[i] = r; }
        if (OP == MUFU16X2_) { uint32_t r; asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a[i])); a[i] = r; }
        if (OP == HMNMX2_) { uint32_t r; asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b[i])); a[i] = r; }
      }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, int per_iter) {
  long long* cyc; int* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  printf("%-10s", name);
  for (int th : {128, 256, 512}) {
    k<OP><<<148, th>>>(10, cyc, sink, 1); k<OP><<<148, th>>>(2000, cyc, sink, 1);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
    s /= 148.0 * 2000;
    printf("  %d w/SMSP: %6.2f clk/inst", th / 128, s / per_iter / (th / 128));
  }
  printf("\n");
}
int main() {
  run<VIMNMX3>("VIMNMX3", 32); run<VIMNMX2>("VIMNMX", 32); run<FMNMX3>("FMNMX3", 32); run<FMNMX2>("FMNMX", 32);
  run<VIADD_>("VIADD", 32); run<LOP3_>("LOP3", 32); run<FFMA2_>("FFMA2", 16); run<FADD2_>("FADD2", 16);
  run<F2FP_>("F2FP", 32); run<MUFU_>("MUFU.EX2", 32); run<I2FP_>("I2FP", 32); run<HMNMX2_>("HMNMX2", 32);
  run<MUFU16_>("EX2.F16", 32); run<MUFU16X2_>("EX2.F16x2", 32); run<HADD2_>("HADD2", 32);
  return 0;
}
