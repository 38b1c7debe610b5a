// Layout conversion at the module boundary: reference layout (B, C, F, T, 2) <-> planes [F][R][Cp].
// Only used where a caller hands in / reads back reference-layout tensors (skiper entries,
// stand-alone primitive modules); the encoder->decoder path stays in planes.
#include "idv_common.cuh"

namespace idv {

__device__ __forceinline__ int round_up8(int c) { return (c + 7) & ~7; }

// grid (ceil(T/32)*ceil(C/32), F, NB), block (32, 8)
__global__ void __launch_bounds__(256) planes_to_user_kernel(const void* __restrict__ planesv, int in_split, int NB,
                                                             int C, int F, int Talloc, int T, float* __restrict__ user) {
  const float* planes = reinterpret_cast<const float*>(planesv);
  const unsigned short* psp = reinterpret_cast<const unsigned short*>(planesv);
  __shared__ float tile[2][32][33];
  const int Ch = round_up8(C), Cp = 2 * Ch, Tp = Talloc + 1;
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
                                                             int Talloc, int T, void* __restrict__ planesv,
                                                             int out_split) {
  float* planes = reinterpret_cast<float*>(planesv);
  unsigned short* psp = reinterpret_cast<unsigned short*>(planesv);
  __shared__ float tile[2][32][33];
  const int Ch = round_up8(C), Cp = 2 * Ch, Tp = Talloc + 1;
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

__global__ void __launch_bounds__(256) z_to_planes_kernel(const float* __restrict__ z, int NB, int S, int s, int Talloc,
                                                          int T, int zdim, void* __restrict__ planesv, int out_split) {
  float* planes = reinterpret_cast<float*>(planesv);
  unsigned short* psp = reinterpret_cast<unsigned short*>(planesv);
  const int Ch = round_up8(zdim), Cp = 2 * Ch, Tp = Talloc + 1;
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
                                  int t_valid, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(planes && user, "idv_planes_to_user: null pointer");
  IDV_CHECK_ARG(NB > 0 && NB <= 65535 && C > 0 && F > 0 && F <= 65535 && T > 0, "idv_planes_to_user: bad shape");
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;      // the user tensor has Tv frames
  dim3 grid(cdiv(Tv, 32) * cdiv(C, 32), F, NB), block(32, 8);
  planes_to_user_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(planes, in_split, NB, C, F, T, Tv, user);
  IDV_LAUNCH_CHECK("planes_to_user_kernel");
  return IDV_OK;
}

extern "C" int idv_user_to_planes(const float* user, int NB, int C, int F, int T, void* planes, int out_split,
                                  int t_valid, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(planes && user, "idv_user_to_planes: null pointer");
  IDV_CHECK_ARG(NB > 0 && NB <= 65535 && C > 0 && F > 0 && F <= 65535 && T > 0, "idv_user_to_planes: bad shape");
  const int Cp = 2 * ((C + 7) / 8 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  // fp32: F*R*Cp floats; split: 2 bf16 plane sets = the same number of bytes
  IDV_CUDA(cudaMemsetAsync(planes, 0, (size_t)F * NB * (T + 1) * Cp * sizeof(float), st));
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  dim3 grid(cdiv(Tv, 32) * cdiv(C, 32), F, NB), block(32, 8);
  user_to_planes_kernel<<<grid, block, 0, st>>>(user, NB, C, F, T, Tv, planes, out_split);
  IDV_LAUNCH_CHECK("user_to_planes_kernel");
  return IDV_OK;
}

extern "C" int idv_z_to_planes(const float* z, int NB, int S, int s, int T, int zdim, void* planes, int out_split,
                               int t_valid, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(planes && z, "idv_z_to_planes: null pointer");
  IDV_CHECK_ARG(NB > 0 && S > 0 && s >= 0 && s < S && T > 0 && zdim > 0, "idv_z_to_planes: bad shape");
  const int Cp = 2 * ((zdim + 7) / 8 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  IDV_CUDA(cudaMemsetAsync(planes, 0, (size_t)NB * (T + 1) * Cp * sizeof(float), st));
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  const int64_t n = (int64_t)NB * Tv * zdim;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  z_to_planes_kernel<<<blocks, 256, 0, st>>>(z, NB, S, s, T, Tv, zdim, planes, out_split);
  IDV_LAUNCH_CHECK("z_to_planes_kernel");
  return IDV_OK;
}

// ------------------------------------------------------------------------------------------------------------
// ComplexBatchNormal, train=True branch (model/complex_progress.py:L131-160 + cbn L168-209), forward only:
//   1. idv_cbn_stats_planes   : per complex channel, over all planes and all non-pad rows: sum r, sum i, sum r^2,
//                               sum i^2, sum r*i (double accumulation)
//   2. idv_cbn_train_finalize : batch mean / biased (co)variances (+eps as the reference adds it), running-stat
//                               update (first call copies, later calls EMA with `momentum`), Z and b' from the
//                               BATCH statistics
//   3. idv_cbn_apply_planes   : y <- PReLU(Z y + b') in place on the planes (pad rows stay zero)
// ------------------------------------------------------------------------------------------------------------
namespace idv {

// grid (F, chunks), block = Ch threads (one complex channel per thread; Ch <= 1024)
__global__ void __launch_bounds__(256) cbn_stats_planes_kernel(const void* __restrict__ planesv, int split, int NB, int C,
                                                               int F, int T, int Tv, int rows_per_chunk,
                                                               double* __restrict__ acc) {
  // block (round32(C), 256 / round32(C)): thread = (complex channel, row lane), shared-memory reduce over the lanes
  __shared__ double red[256];
  const int c = threadIdx.x, ly = threadIdx.y, ny = blockDim.y, Cw = blockDim.x;
  const int Ch = round_up8(C), Cp = 2 * Ch, Tp = T + 1;
  const long long R = (long long)NB * Tp;
  const int f = blockIdx.x;
  const long long r_begin = (long long)blockIdx.y * rows_per_chunk;
  const long long r_end = r_begin + rows_per_chunk < R ? r_begin + rows_per_chunk : R;
  const float* pf = reinterpret_cast<const float*>(planesv);
  const unsigned short* ps = reinterpret_cast<const unsigned short*>(planesv);
  const long long hl = (long long)F * R * Cp;
  double s[5] = {0, 0, 0, 0, 0};
  if (c < C) {
    for (long long r = r_begin + ly; r < r_end; r += ny) {
      const int tt = (int)(r % Tp);
      if (tt == 0 || tt > Tv) continue;                  // causal pad row / beyond the valid frames
      const long long idx = ((long long)f * R + r) * Cp;
      float re, im;
      if (split) {
        re = ld_split1(ps, hl, idx + c);
        im = ld_split1(ps, hl, idx + Ch + c);
      } else {
        re = __ldg(pf + idx + c);
        im = __ldg(pf + idx + Ch + c);
      }
      s[0] += re; s[1] += im;
      s[2] += (double)re * re; s[3] += (double)im * im; s[4] += (double)re * im;
    }
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    red[ly * Cw + c] = s[k];
    __syncthreads();
    if (ly == 0 && c < C) {
      double t = 0.0;
      for (int q = 0; q < ny; ++q) t += red[q * Cw + c];
      atomicAdd(acc + c * 5 + k, t);
    }
    __syncthreads();
  }
}

__global__ void cbn_train_finalize_kernel(const double* __restrict__ acc, double count, int C,
                                          const float* __restrict__ g_rr, const float* __restrict__ g_ri,
                                          const float* __restrict__ g_ii, const float* __restrict__ beta_r,
                                          const float* __restrict__ beta_i, float* __restrict__ run_mr,
                                          float* __restrict__ run_mi, float* __restrict__ run_vrr,
                                          float* __restrict__ run_vri, float* __restrict__ run_vii, float momentum,
                                          int first, float* __restrict__ zb, float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float eps = 1e-5f;
  const double mr = acc[c * 5 + 0] / count, mi = acc[c * 5 + 1] / count;
  // biased second moments of the centred values; the reference adds eps to Vrr / Vii right away (L141-142)
  const float mu_r = (float)mr, mu_i = (float)mi;
  const float vrr = (float)(acc[c * 5 + 2] / count - mr * mr) + eps;
  const float vii = (float)(acc[c * 5 + 3] / count - mi * mi) + eps;
  const float vri = (float)(acc[c * 5 + 4] / count - mr * mi);
  if (stats) {                               // saved for the backward pass (idv_cbn_bwd_*)
    stats[c * 5 + 0] = mu_r; stats[c * 5 + 1] = mu_i; stats[c * 5 + 2] = vrr; stats[c * 5 + 3] = vri; stats[c * 5 + 4] = vii;
  }
  if (first) {
    run_mr[c] = mu_r; run_mi[c] = mu_i; run_vrr[c] = vrr; run_vri[c] = vri; run_vii[c] = vii;
  } else {
    run_mr[c] = momentum * run_mr[c] + (1.f - momentum) * mu_r;
    run_mi[c] = momentum * run_mi[c] + (1.f - momentum) * mu_i;
    run_vrr[c] = momentum * run_vrr[c] + (1.f - momentum) * vrr;
    run_vri[c] = momentum * run_vri[c] + (1.f - momentum) * vri;
    run_vii[c] = momentum * run_vii[c] + (1.f - momentum) * vii;
  }
  // cbn() with the batch statistics (same operation order as model/complex_progress.py:L168-205)
  float delta = vrr * vii - vri * vri + eps;
  delta = delta < 1e-8f ? 1e-8f : delta;
  const float s = sqrtf(delta);
  const float t = sqrtf(vrr + vii + 2.f * s + eps);
  const float ist = 1.f / (s * t + eps);
  const float wrr = (vii + s) * ist, wii = (vrr + s) * ist, wri = -vri * ist;
  const float zrr = g_rr[c] * wrr + g_ri[c] * wri, zri = g_rr[c] * wri + g_ri[c] * wii;
  const float zir = g_ri[c] * wrr + g_ii[c] * wri, zii = g_ri[c] * wri + g_ii[c] * wii;
  zb[c * 6 + 0] = zrr; zb[c * 6 + 1] = zri; zb[c * 6 + 2] = zir; zb[c * 6 + 3] = zii;
  zb[c * 6 + 4] = beta_r[c] - (zrr * mu_r + zri * mu_i);
  zb[c * 6 + 5] = beta_i[c] - (zir * mu_r + zii * mu_i);
}

// grid-stride over (plane, row, channel)
__global__ void __launch_bounds__(256) cbn_apply_planes_kernel(void* __restrict__ planesv, int split, int NB, int C,
                                                               int F, int T, int Tv, const float* __restrict__ zb,
                                                               int apply_prelu, float slope, void* __restrict__ outv,
                                                               int out_split, int in_place) {
  // one thread per (row, complex channel incl. the padding channels); out of place: pad rows / channels written as 0
  const int Ch = round_up8(C), Cp = 2 * Ch, Tp = T + 1;
  const long long R = (long long)NB * Tp;
  const long long n = (long long)F * R * Ch;
  float* pf = reinterpret_cast<float*>(planesv);
  unsigned short* ps = reinterpret_cast<unsigned short*>(planesv);
  float* of = reinterpret_cast<float*>(outv);
  unsigned short* os = reinterpret_cast<unsigned short*>(outv);
  const long long hl = (long long)F * R * Cp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Ch);
    const long long fr = i / Ch;
    const long long r = fr % R;
    const int tt = (int)(r % Tp);
    const bool live = c < C && tt != 0 && tt <= Tv;
    if (!live && in_place) continue;
    const long long idx = fr * Cp;
    float orr = 0.f, oi = 0.f;
    if (live) {
      float re, im;
      if (split) {
        re = ld_split1(ps, hl, idx + c);
        im = ld_split1(ps, hl, idx + Ch + c);
      } else {
        re = pf[idx + c];
        im = pf[idx + Ch + c];
      }
      const float* k = zb + c * 6;
      orr = fmaf(k[0], re, fmaf(k[1], im, k[4]));
      oi = fmaf(k[2], re, fmaf(k[3], im, k[5]));
      if (apply_prelu) {
        orr = prelu_f(orr, slope);
        oi = prelu_f(oi, slope);
      }
    }
    if (out_split) {
      st_split1(os, hl, idx + c, orr);
      st_split1(os, hl, idx + Ch + c, oi);
    } else {
      of[idx + c] = orr;
      of[idx + Ch + c] = oi;
    }
  }
}

// out[b][f][t][p] = x * scale[f][p] + shift[f][p]; zero_edge_imag: the imaginary part of bins 0 and F-1 is set to 0
__global__ void __launch_bounds__(256) bin_affine_kernel(const float* __restrict__ x, long long n, int F, int T,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         int zero_edge_imag, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)((i / T) % F);
    const float2 v = *reinterpret_cast<const float2*>(x + i * 2);
    float2 o;
    o.x = fmaf(v.x, scale[2 * f], shift[2 * f]);
    o.y = fmaf(v.y, scale[2 * f + 1], shift[2 * f + 1]);
    if (zero_edge_imag && (f == 0 || f == F - 1)) o.y = 0.f;
    *reinterpret_cast<float2*>(out + i * 2) = o;
  }
}

}  // namespace idv

extern "C" int idv_bin_affine(const float* x, int B, int F, int T, const float* scale, const float* shift,
                              int zero_edge_imag, float* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(x && scale && shift && out && B > 0 && F > 0 && T > 0, "idv_bin_affine: bad argument");
  const long long n = (long long)B * F * T;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  bin_affine_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n, F, T, scale, shift, zero_edge_imag, out);
  IDV_LAUNCH_CHECK("bin_affine_kernel");
  return IDV_OK;
}

extern "C" int idv_cbn_stats_planes(const void* planes, int split, int NB, int C, int F, int T, double* acc,
                                    int t_valid, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(planes && acc && NB > 0 && C > 0 && C <= 256 && F > 0 && F <= 65535 && T > 0, "idv_cbn_stats_planes: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  IDV_CUDA(cudaMemsetAsync(acc, 0, (size_t)C * 5 * sizeof(double), st));
  const long long R = (long long)NB * (T + 1);
  int chunks = (int)(148LL * 8 / F);
  if (chunks < 1) chunks = 1;
  if (chunks > R) chunks = (int)R;
  const int rows_per_chunk = (int)((R + chunks - 1) / chunks);
  const int Cw = ((C + 31) / 32) * 32;
  dim3 grid(F, chunks), block(Cw, 256 / Cw);
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  cbn_stats_planes_kernel<<<grid, block, 0, st>>>(planes, split, NB, C, F, T, Tv, rows_per_chunk, acc);
  IDV_LAUNCH_CHECK("cbn_stats_planes_kernel");
  return IDV_OK;
}

extern "C" int idv_cbn_train_finalize(const double* acc, double count, int C, const float* gamma_rr,
                                      const float* gamma_ri, const float* gamma_ii, const float* beta_r,
                                      const float* beta_i, float* run_mean_r, float* run_mean_i, float* run_vrr,
                                      float* run_vri, float* run_vii, float momentum, int first, float* zb,
                                      float* stats, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(acc && gamma_rr && gamma_ri && gamma_ii && beta_r && beta_i && run_mean_r && run_mean_i && run_vrr &&
                    run_vri && run_vii && zb && C > 0 && count > 0,
                "idv_cbn_train_finalize: bad argument");
  cbn_train_finalize_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(acc, count, C, gamma_rr, gamma_ri, gamma_ii,
                                                                             beta_r, beta_i, run_mean_r, run_mean_i,
                                                                             run_vrr, run_vri, run_vii, momentum, first,
                                                                             zb, stats);
  IDV_LAUNCH_CHECK("cbn_train_finalize_kernel");
  return IDV_OK;
}

extern "C" int idv_cbn_apply_planes(void* planes, int split, int NB, int C, int F, int T, const float* zb,
                                    int apply_prelu, float prelu_slope, int t_valid, void* out, int out_split,
                                    void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(planes && zb && NB > 0 && C > 0 && F > 0 && T > 0, "idv_cbn_apply_planes: bad argument");
  const int in_place = (!out || out == planes) ? 1 : 0;
  if (in_place) out_split = split;
  const long long n = (long long)F * NB * (T + 1) * ((C + 7) / 8 * 8);
  const int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  cbn_apply_planes_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(planes, split, NB, C, F, T, Tv, zb, apply_prelu,
                                                                    prelu_slope, in_place ? planes : out, out_split ? 1 : 0,
                                                                    in_place);
  IDV_LAUNCH_CHECK("cbn_apply_planes_kernel");
  return IDV_OK;
}

// ---- the same on the reference layout (outer, C, inner, 2): stand-alone ComplexBatchNormal(train=True) and the
//      last decoder layer (C = 1) whose output is produced directly in the reference layout ----------------------
namespace idv {

__global__ void __launch_bounds__(256) cbn_stats_user_kernel(const float* __restrict__ x, long long outer, int C,
                                                             long long inner, double* __restrict__ acc) {
  // grid (chunks, C): block reduces a slice of (outer x inner) for channel c
  const int c = blockIdx.y;
  const long long n = outer * inner;
  double s[5] = {0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long o = i / inner, j = i % inner;
    const float2 v = __ldg(reinterpret_cast<const float2*>(x + ((o * C + c) * inner + j) * 2));
    s[0] += v.x; s[1] += v.y;
    s[2] += (double)v.x * v.x; s[3] += (double)v.y * v.y; s[4] += (double)v.x * v.y;
  }
  __shared__ double red[5][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    double v = s[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    atomicAdd(acc + c * 5 + threadIdx.x, v);
  }
}

// y (NBtot, F, T, 2) in place: PReLU(slope) on re/im, then optionally the mask head with the noisy STFT
// (model/pvae_module.py:L2594-2609); utterance b of y uses stft_x[b / s_rep]
__global__ void __launch_bounds__(256) head_user_kernel(float* __restrict__ y, long long n_per_utt, long long n_utt,
                                                        float slope, int mask, const float* __restrict__ stft_x,
                                                        int s_rep) {
  const long long n = n_per_utt * n_utt;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float2 v = *reinterpret_cast<float2*>(y + i * 2);
    float yr = prelu_f(v.x, slope), yi = prelu_f(v.y, slope);
    if (mask) {
      const long long b = i / n_per_utt, j = i % n_per_utt;
      const float mag = tanhf(sqrtf(yr * yr + yi * yi));
      const float ph = atan2f(yi / (mag + 1e-8f), yr / (mag + 1e-8f));
      const float2 X = __ldg(reinterpret_cast<const float2*>(stft_x + ((b / s_rep) * n_per_utt + j) * 2));
      const float in_mag = sqrtf(X.x * X.x + X.y * X.y);
      const float in_ph = atan2f(X.y, X.x);
      float sn, cs;
      sincosf(in_ph + ph, &sn, &cs);
      yr = in_mag * mag * cs;
      yi = in_mag * mag * sn;
    }
    *reinterpret_cast<float2*>(y + i * 2) = make_float2(yr, yi);
  }
}

}  // namespace idv

extern "C" int idv_cbn_stats_user(const float* x, int64_t outer, int C, int64_t inner, double* acc, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(x && acc && outer > 0 && C > 0 && C <= 65535 && inner > 0, "idv_cbn_stats_user: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  IDV_CUDA(cudaMemsetAsync(acc, 0, (size_t)C * 5 * sizeof(double), st));
  const long long n = outer * inner;
  int chunks = (int)((n + 256 * 64 - 1) / (256 * 64));
  if (chunks < 1) chunks = 1;
  if (chunks > 1024) chunks = 1024;
  dim3 grid(chunks, C);
  cbn_stats_user_kernel<<<grid, 256, 0, st>>>(x, outer, C, inner, acc);
  IDV_LAUNCH_CHECK("cbn_stats_user_kernel");
  return IDV_OK;
}

extern "C" int idv_head_user(float* y, int64_t n_per_utt, int64_t n_utt, float prelu_slope, int mask,
                             const float* stft_x, int s_rep, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(y && n_per_utt > 0 && n_utt > 0 && s_rep > 0 && (!mask || stft_x), "idv_head_user: bad argument");
  const long long n = n_per_utt * n_utt;
  const int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
  head_user_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y, n_per_utt, n_utt, prelu_slope, mask, stft_x, s_rep);
  IDV_LAUNCH_CHECK("head_user_kernel");
  return IDV_OK;
}
