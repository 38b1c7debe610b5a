// Layout conversion at the module boundary: reference layout (B, C, F, T, 2) <-> planes [F][R][Cp].
// Only used where a caller hands in / reads back reference-layout tensors (skiper entries,
// stand-alone primitive modules); the encoder->decoder path stays in planes.
#include "idv_common.cuh"

namespace idv {

__device__ __forceinline__ int round_up8(int c) { return (c + 7) & ~7; }

// grid (ceil(T/32)*ceil(C/32), F, NB), block (32, 8)
__global__ void __launch_bounds__(256) planes_to_user_kernel(const void* __restrict__ planesv, int in_split, int NB,
                                                             int C, int F, int T, float* __restrict__ user) {
  const float* planes = reinterpret_cast<const float*>(planesv);
  const unsigned short* psp = reinterpret_cast<const unsigned short*>(planesv);
  __shared__ float tile[2][32][33];
  const int Ch = round_up8(C), Cp = 2 * Ch, Tp = T + 1;
  const int64_t R = (int64_t)NB * Tp;
  const int ct = (C + 31) / 32;
  const int c0 = (blockIdx.x % ct) * 32, t0 = (blockIdx.x / ct) * 32;
  const int f = blockIdx.y, b = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    float re = 0.f, im = 0.f;
    if (t < T && c < C) {
      const long long ridx = ((int64_t)f * R + (int64_t)b * Tp + 1 + t) * Cp;
      if (in_split) {
        const long long hl = (long long)F * R * Cp;
        re = ld_split1(psp, hl, ridx + c);
        im = ld_split1(psp, hl, ridx + Ch + c);
      } else {
        re = __ldg(planes + ridx + c);
        im = __ldg(planes + ridx + Ch + c);
      }
    }
    tile[0][i][tx] = re;
    tile[1][i][tx] = im;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    if (c < C && t < T) {
      *reinterpret_cast<float2*>(user + ((((int64_t)b * C + c) * F + f) * T + t) * 2) =
          make_float2(tile[0][tx][i], tile[1][tx][i]);
    }
  }
}

__global__ void __launch_bounds__(256) user_to_planes_kernel(const float* __restrict__ user, int NB, int C, int F,
                                                             int T, void* __restrict__ planesv, int out_split) {
  float* planes = reinterpret_cast<float*>(planesv);
  unsigned short* psp = reinterpret_cast<unsigned short*>(planesv);
  __shared__ float tile[2][32][33];
  const int Ch = round_up8(C), Cp = 2 * Ch, Tp = T + 1;
  const int64_t R = (int64_t)NB * Tp;
  const int ct = (C + 31) / 32;
  const int c0 = (blockIdx.x % ct) * 32, t0 = (blockIdx.x / ct) * 32;
  const int f = blockIdx.y, b = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (c < C && t < T)
      v = __ldg(reinterpret_cast<const float2*>(user + ((((int64_t)b * C + c) * F + f) * T + t) * 2));
    tile[0][i][tx] = v.x;
    tile[1][i][tx] = v.y;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < T && c < C) {
      const long long ridx = ((int64_t)f * R + (int64_t)b * Tp + 1 + t) * Cp;
      if (out_split) {
        const long long hl = (long long)F * R * Cp;
        st_split1(psp, hl, ridx + c, tile[0][tx][i]);
        st_split1(psp, hl, ridx + Ch + c, tile[1][tx][i]);
      } else {
        planes[ridx + c] = tile[0][tx][i];
        planes[ridx + Ch + c] = tile[1][tx][i];
      }
    }
  }
}

__global__ void __launch_bounds__(256) z_to_planes_kernel(const float* __restrict__ z, int NB, int S, int s, int T,
                                                          int zdim, void* __restrict__ planesv, int out_split) {
  float* planes = reinterpret_cast<float*>(planesv);
  unsigned short* psp = reinterpret_cast<unsigned short*>(planesv);
  const int Ch = round_up8(zdim), Cp = 2 * Ch, Tp = T + 1;
  const int64_t n = (int64_t)NB * T * zdim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % zdim);
    const int64_t bt = i / zdim;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const float2 v = __ldg(reinterpret_cast<const float2*>(z + ((((int64_t)b * S + s) * T + t) * zdim + j) * 2));
    const long long ridx = ((int64_t)b * Tp + 1 + t) * Cp;
    if (out_split) {
      const long long hl = (long long)NB * Tp * Cp;
      st_split1(psp, hl, ridx + j, v.x);
      st_split1(psp, hl, ridx + Ch + j, v.y);
    } else {
      planes[ridx + j] = v.x;
      planes[ridx + Ch + j] = v.y;
    }
  }
}

// stand-alone ComplexBatchNormal(train=False) on the reference layout: out = Z (x) + b' per channel
__global__ void __launch_bounds__(256) cbn_eval_user_kernel(const float* __restrict__ x, int64_t outer, int C,
                                                            int64_t inner, const float* __restrict__ zb,
                                                            float* __restrict__ out) {
  const int64_t n = outer * C * inner;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / inner) % C);
    const float2 v = __ldg(reinterpret_cast<const float2*>(x + i * 2));
    const float* k = zb + c * 6;
    const float re = fmaf(__ldg(k + 0), v.x, fmaf(__ldg(k + 1), v.y, __ldg(k + 4)));
    const float im = fmaf(__ldg(k + 2), v.x, fmaf(__ldg(k + 3), v.y, __ldg(k + 5)));
    *reinterpret_cast<float2*>(out + i * 2) = make_float2(re, im);
  }
}

}  // namespace idv

extern "C" int idv_cbn_eval_user(const float* x, int64_t outer, int C, int64_t inner, const float* zb, float* out,
                                 void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(x && zb && out && outer > 0 && C > 0 && inner > 0, "idv_cbn_eval_user: bad argument");
  const int64_t n = outer * C * inner;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  cbn_eval_user_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, outer, C, inner, zb, out);
  IDV_LAUNCH_CHECK("cbn_eval_user_kernel");
  return IDV_OK;
}

extern "C" int idv_planes_to_user(const void* planes, int in_split, int NB, int C, int F, int T, float* user,
                                  void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(planes && user, "idv_planes_to_user: null pointer");
  IDV_CHECK_ARG(NB > 0 && NB <= 65535 && C > 0 && F > 0 && F <= 65535 && T > 0, "idv_planes_to_user: bad shape");
  dim3 grid(cdiv(T, 32) * cdiv(C, 32), F, NB), block(32, 8);
  planes_to_user_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(planes, in_split, NB, C, F, T, user);
  IDV_LAUNCH_CHECK("planes_to_user_kernel");
  return IDV_OK;
}

extern "C" int idv_user_to_planes(const float* user, int NB, int C, int F, int T, void* planes, int out_split,
                                  void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(planes && user, "idv_user_to_planes: null pointer");
  IDV_CHECK_ARG(NB > 0 && NB <= 65535 && C > 0 && F > 0 && F <= 65535 && T > 0, "idv_user_to_planes: bad shape");
  const int Cp = 2 * ((C + 7) / 8 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  // fp32: F*R*Cp floats; split: 2 bf16 plane sets = the same number of bytes
  IDV_CUDA(cudaMemsetAsync(planes, 0, (size_t)F * NB * (T + 1) * Cp * sizeof(float), st));
  dim3 grid(cdiv(T, 32) * cdiv(C, 32), F, NB), block(32, 8);
  user_to_planes_kernel<<<grid, block, 0, st>>>(user, NB, C, F, T, planes, out_split);
  IDV_LAUNCH_CHECK("user_to_planes_kernel");
  return IDV_OK;
}

extern "C" int idv_z_to_planes(const float* z, int NB, int S, int s, int T, int zdim, void* planes, int out_split,
                               void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(planes && z, "idv_z_to_planes: null pointer");
  IDV_CHECK_ARG(NB > 0 && S > 0 && s >= 0 && s < S && T > 0 && zdim > 0, "idv_z_to_planes: bad shape");
  const int Cp = 2 * ((zdim + 7) / 8 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  IDV_CUDA(cudaMemsetAsync(planes, 0, (size_t)NB * (T + 1) * Cp * sizeof(float), st));
  const int64_t n = (int64_t)NB * T * zdim;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  z_to_planes_kernel<<<blocks, 256, 0, st>>>(z, NB, S, s, T, zdim, planes, out_split);
  IDV_LAUNCH_CHECK("z_to_planes_kernel");
  return IDV_OK;
}
