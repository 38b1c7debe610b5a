// Shared helpers for the libidv_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/idv.h"

namespace idv {

void set_error(const char* fmt, ...);

#define IDV_CHECK_ARG(cond, ...)             \
  do {                                       \
    if (!(cond)) {                           \
      ::idv::set_error(__VA_ARGS__);         \
      return IDV_E_ARG;                      \
    }                                        \
  } while (0)

#define IDV_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::idv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                        \
      return IDV_E_CUDA;                                                                 \
    }                                                                                    \
  } while (0)

#define IDV_LAUNCH_CHECK(name)                                                  \
  do {                                                                          \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) {                                                    \
      ::idv::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
      return IDV_E_CUDA;                                                        \
    }                                                                           \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float prelu_f(float v, float a) { return v > 0.f ? v : a * v; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

}  // namespace idv
