// Shared helpers for the zest_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/zest_b200.h"

namespace zest {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define ZEST_CHECK_ARG(cond, ...)                  \
  do {                                             \
    if (!(cond)) {                                 \
      zest::set_error(__VA_ARGS__);                \
      return ZEST_E_ARG;                           \
    }                                              \
  } while (0)

#define ZEST_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      zest::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                      __LINE__);                                                          \
      return ZEST_E_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#ifndef ZEST_TRY
#define ZEST_TRY(expr) do { int _r = (expr); if (_r != ZEST_OK) return _r; } while (0)
#endif

// call after every kernel launch: catches launch-configuration errors without synchronising
#define ZEST_LAUNCH_CHECK()                                                                \
  do {                                                                                     \
    zest::count_launch();                                                                  \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      zest::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                           \
      return ZEST_E_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

static inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// dot of a 3-vector with a matrix row, in the summation order of ATen's CPU matmul for K = 3
// (fma(v2, r2, fma(v1, r1, v0 * r0)), verified bit-exact).
__device__ __forceinline__ float dot3(float v0, float v1, float v2, float r0, float r1, float r2) {
  return __fmaf_rn(v2, r2, __fmaf_rn(v1, r1, __fmul_rn(v0, r0)));
}

// ATen GridSampler.cuh safe_downgrade_to_int_range
__device__ __forceinline__ float safe_int_range(float x) {
  if (x > 2147483646.f || x < -2147483648.f || !isfinite(x)) return -100.f;
  return x;
}

// align_corners=True un-normalisation, one rounding per torch op: ((g + 1) / 2) * (size - 1)
__device__ __forceinline__ float unnormalize(float g, int size) {
  return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), (float)(size - 1));
}

}  // namespace zest
