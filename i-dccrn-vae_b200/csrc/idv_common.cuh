// Shared helpers for the libidv_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/idv.h"

namespace idv {

void set_error(const char* fmt, ...);
int option_lstm_ncols();
int option_dynamic_tiles();
int option_gemm_pairs();
int option_lstm_wave_pairs();
int option_tile_order();
int option_tma_store();
int option_lstm_interleave();
int option_lstm_sync_mode();
int option_launch_pdl();
int option_splitk();
int option_lstm_cluster_alt();
int option_lstm_chunk_sync();
int option_lstm_tma_publish();

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

// ---- programmatic dependent launch (option launch_pdl) ----------------------------------------------------------------
// A kernel launched through launch_pdl may start while its predecessor on the stream is still running (the predecessor
// triggers with pdl_trigger(), or implicitly when it exits): its launch latency, block scheduling and set-up (barrier
// initialisation, TMEM allocation, descriptor prefetch) overlap the predecessor's tail.  Contract of every kernel
// launched this way: EVERY thread of EVERY block executes pdl_wait() before its first global-memory access that is not
// to launch-invariant data (weights, tables); pdl_wait() returns when all predecessor grids have completed and their
// writes are visible.  Both instructions are no-ops in a normal launch.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  if (option_launch_pdl()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float prelu_f(float v, float a) { return v > 0.f ? v : a * v; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ---- split-bf16 activation format: x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi) ----------------
__device__ __forceinline__ void split_bf16(float x, unsigned short& hi, unsigned short& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = __bfloat16_as_ushort(h);
  lo = __bfloat16_as_ushort(l);
}
__device__ __forceinline__ float bf16_bits_to_float(unsigned short b) { return __uint_as_float((unsigned int)b << 16); }
// 4 consecutive channels of a split tensor -> float4 (hi + lo)
__device__ __forceinline__ float4 ld_split4(const unsigned short* hi, long long hl, long long idx) {
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(hi + idx));
  const uint2 l = __ldg(reinterpret_cast<const uint2*>(hi + hl + idx));
  float4 r;
  r.x = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
  r.y = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
  r.z = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
  r.w = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
  return r;
}
__device__ __forceinline__ void st_split4(unsigned short* hi, long long hl, long long idx, float4 v) {
  unsigned short h[4], l[4];
  split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
  *reinterpret_cast<uint2*>(hi + idx) = make_uint2((unsigned)h[0] | ((unsigned)h[1] << 16), (unsigned)h[2] | ((unsigned)h[3] << 16));
  *reinterpret_cast<uint2*>(hi + hl + idx) = make_uint2((unsigned)l[0] | ((unsigned)l[1] << 16), (unsigned)l[2] | ((unsigned)l[3] << 16));
}
__device__ __forceinline__ void st_split2(unsigned short* hi, long long hl, long long idx, float2 v) {
  unsigned short h[2], l[2];
  split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]);
  *reinterpret_cast<unsigned*>(hi + idx) = (unsigned)h[0] | ((unsigned)h[1] << 16);
  *reinterpret_cast<unsigned*>(hi + hl + idx) = (unsigned)l[0] | ((unsigned)l[1] << 16);
}
__device__ __forceinline__ float ld_split1(const unsigned short* hi, long long hl, long long idx) {
  return bf16_bits_to_float(__ldg(hi + idx)) + bf16_bits_to_float(__ldg(hi + hl + idx));
}
__device__ __forceinline__ void st_split1(unsigned short* hi, long long hl, long long idx, float v) {
  unsigned short h, l;
  split_bf16(v, h, l);
  hi[idx] = h;
  hi[hl + idx] = l;
}

}  // namespace idv
