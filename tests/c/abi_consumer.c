/* A plain-C consumer of include/qmha.h — what a cgo / JNI / FFI binding of the reference would see.
 * Built by tests/test_host_cpu.py (gcc -std=c99 -Wall -Werror, linked against libqmha.so: the header is valid C and
 * every symbol used resolves) and run two ways:
 *   abi_consumer sizes      no GPU needed: prints sizeof(qmha_args) and the struct_size qmha_args_init() stamps
 *   abi_consumer run        on a B200: the reference's all-ones known-answer test (drivers/main.cu:73-101) through
 *                           solve() and through qmha_forward_ex() with a strided output slab; exit 0 = pass
 * Device memory comes from the CUDA runtime through its C API (cuda_runtime_api.h). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "qmha.h"

static int check_ones(const float* h, size_t n, const char* what) {
  size_t i;
  for (i = 0; i < n; ++i)
    if (fabsf(h[i] - 1.0f) > 2e-2f) {
      fprintf(stderr, "%s: element %zu = %f, expected 1\n", what, i, h[i]);
      return 1;
    }
  return 0;
}

int main(int argc, char** argv) {
  qmha_args a;
  qmha_args_init(&a);
  if (argc > 1 && strcmp(argv[1], "sizes") == 0) {
    printf("%zu %zu %d\n", sizeof(qmha_args), a.struct_size, qmha_kernel_from_name("fa_tc_int8_b"));
    return a.struct_size == sizeof(qmha_args) ? 0 : 1;
  }
  {
    const int N = 512, d_model = 256, h = 2, wide = 2 * d_model;
    const size_t n = (size_t)N * d_model;
    float *q = NULL, *o = NULL, *big = NULL, *host = (float*)malloc(n * sizeof(float) * 2);
    size_t i, r;
    if (!host || cudaMalloc((void**)&q, n * 4) != cudaSuccess || cudaMalloc((void**)&o, n * 4) != cudaSuccess ||
        cudaMalloc((void**)&big, n * 4 * 2) != cudaSuccess) {
      fprintf(stderr, "allocation failed\n");
      return 2;
    }
    for (i = 0; i < n; ++i) host[i] = 1.0f;
    cudaMemcpy(q, host, n * 4, cudaMemcpyHostToDevice);
    /* 1. the reference's entry point: synchronous, void */
    solve(q, q, q, o, N, d_model, h);
    if (qmha_last_error()[0]) { fprintf(stderr, "solve: %s\n", qmha_last_error()); return 1; }
    cudaMemcpy(host, o, n * 4, cudaMemcpyDeviceToHost);
    if (check_ones(host, n, "solve")) return 1;
    /* 2. the extended entry: FP16 kernel, output written into the right half of a [N, 2*d_model] tensor */
    cudaMemset(big, 0, n * 4 * 2);
    a.Q = q; a.K = q; a.V = q; a.O = big + d_model;
    a.B = 1; a.N = N; a.d_model = d_model; a.h = h;
    a.kernel = QMHA_KERNEL_F16;
    a.o_row_stride = wide;
    if (qmha_forward_ex(&a) != 0 || qmha_synchronize(NULL) != 0) { fprintf(stderr, "forward_ex: %s\n", qmha_last_error()); return 1; }
    cudaMemcpy(host, big, n * 4 * 2, cudaMemcpyDeviceToHost);
    for (r = 0; r < (size_t)N; ++r) {
      if (check_ones(host + r * wide + d_model, (size_t)d_model, "forward_ex slab")) return 1;
      for (i = 0; i < (size_t)d_model; ++i)
        if (host[r * wide + i] != 0.0f) { fprintf(stderr, "forward_ex wrote outside its slab\n"); return 1; }
    }
    qmha_shutdown();
    cudaFree(q); cudaFree(o); cudaFree(big); free(host);
    printf("abi_consumer ok (%s)\n", qmha_version());
  }
  return 0;
}
