// common.cuh — shared helpers for libbsrnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bsrnn_b200.h"

namespace bsrnn {

void set_error(const char* fmt, ...);
void count_launches(long n);   // bookkeeping for bsrnn_launch_count()

#define BSRNN_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      bsrnn::set_error(__VA_ARGS__);          \
      return 1;                               \
    }                                         \
  } while (0)

#define BSRNN_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      bsrnn::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

#define BSRNN_LAUNCH_OK()                  \
  do {                                     \
    BSRNN_CUDA_OK(cudaGetLastError());     \
    bsrnn::count_launches(1);              \
  } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace bsrnn
