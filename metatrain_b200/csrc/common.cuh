// Shared helpers for libpetb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/petb200.h"

namespace petb200 {

// thread-local last-error message, read through petb200_last_error()
void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t err = cudaPeekAtLastError();
  if (err != cudaSuccess) {
    cudaGetLastError();  // clear
    set_error("%s: %s", what, cudaGetErrorString(err));
    return PETB200_ERR_CUDA;
  }
  return PETB200_OK;
}

#define PETB200_REQUIRE(cond, ...)          \
  do {                                      \
    if (!(cond)) {                          \
      petb200::set_error(__VA_ARGS__);      \
      return PETB200_ERR_INVALID_ARGUMENT;  \
    }                                       \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// d/dx [x * sigmoid(x)]
__device__ __forceinline__ float dsiluf_(float x) {
  float s = sigmoidf_(x);
  return s * (1.0f + x * (1.0f - s));
}

}  // namespace petb200
