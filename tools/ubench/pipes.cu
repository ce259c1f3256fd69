// Which pipe does each SASS instruction of the softmax loop occupy, and what do PAIRS of them cost
// when interleaved 1:1?  (If t(A+B) == max(tA, tB) the two sit on different pipes; if it is the sum
// they share one.)  All bodies are inline PTX on runtime operands so that nothing is folded; check
// `cuobjdump -sass bin/ubench_pipes` for what ptxas picked.  Development aid.
//   clocks per warp-level instruction (or per A+B pair) per SM sub-partition, 1 / 2 / 4 warps per SMSP.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { IADD_IMM, IMAD_ONE, LOP_XOR, FFMA_R, FFMA_IMM, FFMA2, FADD2, F2FP, I2FP, VIMNMX3, MUFU, LEA_, PRMT_, FMNMX_, NONE };

template <int OP>
__device__ __forceinline__ void op(uint32_t& a, uint32_t& a2, uint32_t b, uint32_t one, uint32_t c) {
  if constexpr (OP == IADD_IMM) asm volatile("add.s32 %0, %0, 0x4B400000;" : "+r"(a));
  if constexpr (OP == IMAD_ONE) asm volatile("mad.lo.s32 %0, %0, %1, 0x4B400000;" : "+r"(a) : "r"(one));
  if constexpr (OP == LOP_XOR) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a) : "r"(b));
  if constexpr (OP == FFMA_R) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  if constexpr (OP == FFMA_IMM) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F800000;" : "+r"(a) : "r"(b));
  if constexpr (OP == FFMA2) {
    asm volatile("{\n\t.reg .b64 x, y, z;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\tmov.b64 z, {%3, %3};\n\t"
                 "fma.rn.f32x2 x, x, y, z;\n\tmov.b64 {%0, %1}, x;\n\t}" : "+r"(a), "+r"(a2) : "r"(b), "r"(c));
  }
  if constexpr (OP == FADD2) {
    asm volatile("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\t"
                 "add.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}" : "+r"(a), "+r"(a2) : "r"(b));
  }
  if constexpr (OP == F2FP) asm volatile("cvt.rn.f16x2.f32 %0, %0, %1;" : "+r"(a) : "r"(b));
  if constexpr (OP == I2FP) asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(a));
  if constexpr (OP == VIMNMX3) asm volatile("{\n\t.reg .s32 t;\n\tmax.s32 t, %0, %1;\n\tmax.s32 %0, t, %2;\n\t}" : "+r"(a) : "r"(b), "r"(c));
  if constexpr (OP == MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a));
  if constexpr (OP == LEA_) asm volatile("{\n\t.reg .b32 t;\n\tshl.b32 t, %0, 23;\n\tadd.s32 %0, t, %1;\n\t}" : "+r"(a) : "r"(b));
  if constexpr (OP == PRMT_) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(a) : "r"(b));
  if constexpr (OP == FMNMX_) asm volatile("max.f32 %0, %0, %1;" : "+r"(a) : "r"(b));
}

// NA instructions of A then NB of B per chain slot (8 chains, 4x unrolled)
template <int A, int B, int NA, int NB>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cyc, uint32_t* sink, uint32_t one, uint32_t bb, uint32_t cc) {
  uint32_t a[8], a2[8], x[8], x2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 7 + i; a2[i] = a[i] ^ 0x55; x[i] = threadIdx.x * 3 + i * 5; x2[i] = x[i] + 9; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int r = 0; r < NA; ++r) op<A>(a[i], a2[i], bb, one, cc);
#pragma unroll
        for (int r = 0; r < NB; ++r) op<B>(x[i], x2[i], bb, one, cc);
      }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i] ^ a2[i] ^ x[i] ^ x2[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int A, int B, int NA = 1, int NB = 1>
void run(const char* name) {
  long long* cyc; uint32_t* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
  printf("%-28s", name);
  for (int th : {128, 256, 512}) {
    k<A, B, NA, NB><<<148, th>>>(10, cyc, sink, 1, 0x3f800123, 0x3e000000);
    k<A, B, NA, NB><<<148, th>>>(1000, cyc, sink, 1, 0x3f800123, 0x3e000000);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
    s /= 148.0 * 1000;
    printf("  %dw: %6.2f", th / 128, s / 32 / (th / 128));
  }
  printf("   clk per group (%d A + %d B)\n", NA, B == NONE ? 0 : NB);
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  printf("# single instructions\n");
  run<IADD_IMM, NONE>("add.s32 imm (VIADD/IADD3)");
  run<IMAD_ONE, NONE>("mad.lo r*one+imm (IMAD)");
  run<LOP_XOR, NONE>("xor (LOP3)");
  run<FFMA_R, NONE>("fma reg (FFMA)");
  run<FFMA_IMM, NONE>("fma imm (FFMA)");
  run<FFMA2, NONE>("fma.f32x2 (FFMA2)");
  run<FADD2, NONE>("add.f32x2 (FADD2)");
  run<F2FP, NONE>("cvt.f16x2.f32 (F2FP)");
  run<I2FP, NONE>("cvt.f32.s32 (I2FP)");
  run<VIMNMX3, NONE>("max3.s32 (VIMNMX3)");
  run<MUFU, NONE>("ex2 (MUFU)");
  run<LEA_, NONE>("shl+add (LEA?)");
  run<PRMT_, NONE>("prmt (PRMT)");
  run<FMNMX_, NONE>("max.f32 (FMNMX)");
  printf("# pairs, 1:1\n");
  run<IADD_IMM, IMAD_ONE>("IADD + IMAD");
  run<IADD_IMM, FFMA2>("IADD + FFMA2");
  run<IMAD_ONE, FFMA2>("IMAD + FFMA2");
  run<IMAD_ONE, FFMA_R>("IMAD + FFMA");
  run<IADD_IMM, F2FP>("IADD + F2FP");
  run<IADD_IMM, VIMNMX3>("IADD + VIMNMX3");
  run<IADD_IMM, FFMA_R>("IADD + FFMA");
  run<F2FP, FFMA2>("F2FP + FFMA2");
  run<F2FP, VIMNMX3>("F2FP + VIMNMX3");
  run<FFMA2, FADD2>("FFMA2 + FADD2");
  run<IADD_IMM, LOP_XOR>("IADD + LOP3");
  run<I2FP, IADD_IMM>("I2FP + IADD");
  run<I2FP, FFMA2>("I2FP + FFMA2");
  printf("# with the MUFU (8 clk): how much other work hides under one ex2\n");
  run<MUFU, IADD_IMM, 1, 1>("MUFU + 1 IADD");
  run<MUFU, IADD_IMM, 1, 2>("MUFU + 2 IADD");
  run<MUFU, IADD_IMM, 1, 4>("MUFU + 4 IADD");
  run<MUFU, IMAD_ONE, 1, 2>("MUFU + 2 IMAD");
  run<MUFU, IMAD_ONE, 1, 4>("MUFU + 4 IMAD");
  run<MUFU, FFMA2, 1, 2>("MUFU + 2 FFMA2");
  run<MUFU, FFMA2, 1, 4>("MUFU + 4 FFMA2");
  run<MUFU, F2FP, 1, 2>("MUFU + 2 F2FP");
  run<MUFU, F2FP, 1, 4>("MUFU + 4 F2FP");
  run<MUFU, FFMA_R, 1, 4>("MUFU + 4 FFMA");
  run<MUFU, FFMA_R, 1, 8>("MUFU + 8 FFMA");
  return 0;
}
