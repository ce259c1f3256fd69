/* check_cuda.h — same contract as the reference's tools/check_cuda.h:8-14:
 * print file:line + cudaGetErrorString and exit(1). */
#ifndef CHECK_CUDA_H
#define CHECK_CUDA_H
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CHECK_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t err__ = (call);                                                       \
    if (err__ != cudaSuccess) {                                                       \
      fprintf(stderr, "CUDA Error at: %s:%d: %s \n", __FILE__, __LINE__,              \
              cudaGetErrorString(err__));                                             \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)
#endif
