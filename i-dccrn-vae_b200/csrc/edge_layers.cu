// The two layers that are not GEMM-shaped, written as HBM-bound SIMT kernels:
//   enc0: causal ComplexConv2d(1 -> Cout) + folded CBN + PReLU, reading the user-layout STFT
//   dec5: causal ComplexConvTranspose2d(Cin -> 1) + folded CBN + PReLU (+ mask head), writing the
//         user-layout spectrum `predict`.
#include "idv_common.cuh"

namespace idv {

// grid: (row tiles of 32, Fout); block 256 = 32 rows x 8 column groups
__global__ void __launch_bounds__(256) enc0_kernel(const float* __restrict__ stft, int NB, int Fin, int T,
                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                   int N, float slope, void* __restrict__ outv, int Fout,
                                                   int out_split) {
  extern __shared__ __align__(16) float ws[];       // [20][N] then bias [N]
  const int tid = threadIdx.x;
  for (int i = tid; i < 20 * N; i += 256) ws[i] = __ldg(w + i);
  for (int i = tid; i < N; i += 256) ws[20 * N + i] = __ldg(bias + i);
  __syncthreads();
  const int Tp = T + 1;
  const int R = NB * Tp;
  const int r = blockIdx.x * 32 + (tid >> 3);
  const int cg = tid & 7;
  const int fo = blockIdx.y;
  if (r >= R) return;
  const int b = r / Tp, t = r % Tp - 1;
  const long long oidx = ((long long)fo * R + r) * N;
  const long long hl = (long long)Fout * R * N;
  float* orow = reinterpret_cast<float*>(outv) + oidx;
  unsigned short* osp = reinterpret_cast<unsigned short*>(outv);
  if (t < 0) {                                       // causal pad row
    for (int n = cg * 8; n < N; n += 64) {
      if (out_split) {
        *reinterpret_cast<uint4*>(osp + oidx + n) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(osp + hl + oidx + n) = make_uint4(0u, 0u, 0u, 0u);
      } else {
        *reinterpret_cast<float4*>(orow + n) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(orow + n + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    return;
  }
  float xin[20];                                     // [kf][kt][part]
#pragma unroll
  for (int kf = 0; kf < 5; ++kf) {
    const int fi = 2 * fo + kf - 2;
    const bool fok = (fi >= 0 && fi < Fin);
    const float* xp = stft + ((int64_t)(b * Fin + (fok ? fi : 0)) * T) * 2;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      const int ti = t - 1 + kt;
      float2 v = make_float2(0.f, 0.f);
      if (fok && ti >= 0) v = __ldg(reinterpret_cast<const float2*>(xp + (int64_t)ti * 2));
      xin[(kf * 2 + kt) * 2 + 0] = v.x;
      xin[(kf * 2 + kt) * 2 + 1] = v.y;
    }
  }
  // 8 consecutive channels per thread: one 16-byte store per plane (hi / lo) or two float4
  for (int n = cg * 8; n < N; n += 64) {
    float acc[8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(&ws[20 * N + n]);
      const float4 b1 = *reinterpret_cast<const float4*>(&ws[20 * N + n + 4]);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
      acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
#pragma unroll
    for (int k = 0; k < 20; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[k * N + n]);
      const float4 w1 = *reinterpret_cast<const float4*>(&ws[k * N + n + 4]);
      acc[0] = fmaf(xin[k], w0.x, acc[0]); acc[1] = fmaf(xin[k], w0.y, acc[1]);
      acc[2] = fmaf(xin[k], w0.z, acc[2]); acc[3] = fmaf(xin[k], w0.w, acc[3]);
      acc[4] = fmaf(xin[k], w1.x, acc[4]); acc[5] = fmaf(xin[k], w1.y, acc[5]);
      acc[6] = fmaf(xin[k], w1.z, acc[6]); acc[7] = fmaf(xin[k], w1.w, acc[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = prelu_f(acc[j], slope);
    if (out_split) {
      unsigned short h[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_bf16(acc[j], h[j], l[j]);
      *reinterpret_cast<uint4*>(osp + oidx + n) =
          make_uint4((unsigned)h[0] | ((unsigned)h[1] << 16), (unsigned)h[2] | ((unsigned)h[3] << 16),
                     (unsigned)h[4] | ((unsigned)h[5] << 16), (unsigned)h[6] | ((unsigned)h[7] << 16));
      *reinterpret_cast<uint4*>(osp + hl + oidx + n) =
          make_uint4((unsigned)l[0] | ((unsigned)l[1] << 16), (unsigned)l[2] | ((unsigned)l[3] << 16),
                     (unsigned)l[4] | ((unsigned)l[5] << 16), (unsigned)l[6] | ((unsigned)l[7] << 16));
    } else {
      *reinterpret_cast<float4*>(orow + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(orow + n + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

// grid: (row tiles of 32, Fout); block 256 = 8 warps x 4 rows each; one warp per output bin
constexpr int D5_ROWS = 32;
__global__ void __launch_bounds__(256) dec5_head_kernel(const void* __restrict__ pv, int p_cp,
                                                        const void* __restrict__ skipv, int s_cp, int in_split, int NB,
                                                        int Fin, int T, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float slope, int mask,
                                                        const float* __restrict__ stft_x,
                                                        float* __restrict__ predict, int out_bmul,
                                                        int out_boff) {
  extern __shared__ __align__(16) float ws[];       // [10][ktot][2]
  const int ktot = p_cp + s_cp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 10 * ktot * 2; i += 256) ws[i] = __ldg(w + i);
  __syncthreads();
  const int Tp = T + 1;
  const int R = NB * Tp;
  const int Fout = 2 * Fin - 1;
  const int fo = blockIdx.y;
  const float b_r = __ldg(bias), b_i = __ldg(bias + 1);
  const int nq = ktot / 4;                           // float4 chunks per input row (p then skip)
  const int pq = p_cp / 4;
  const float* p = reinterpret_cast<const float*>(pv);
  const float* skip = reinterpret_cast<const float*>(skipv);
  const unsigned short* psp = reinterpret_cast<const unsigned short*>(pv);
  const unsigned short* ssp = reinterpret_cast<const unsigned short*>(skipv);
  const long long p_hl = (long long)Fin * R * p_cp, s_hl = (long long)Fin * R * s_cp;
  for (int rr = 0; rr < D5_ROWS / 8; ++rr) {
    const int r = blockIdx.x * D5_ROWS + warp * (D5_ROWS / 8) + rr;
    if (r >= R) break;                               // warp-uniform
    const int b = r / Tp, t = r % Tp - 1;
    if (t < 0) continue;
    float yr = 0.f, yi = 0.f;
    for (int kf = (fo & 1) ? 1 : 0; kf < 5; kf += 2) {
      const int fi2 = fo + 2 - kf;                   // even by construction
      const int fi = fi2 >> 1;
      if (fi < 0 || fi >= Fin) continue;
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {
        const int ra = r - kt;                       // row r-1 of t == 0 is the zero pad row
        const float* wt = ws + (kf * 2 + kt) * ktot * 2;
        for (int q = lane; q < nq; q += 32) {
          float4 a;
          if (in_split) {
            if (q < pq) a = ld_split4(psp, p_hl, ((long long)fi * R + ra) * p_cp + q * 4);
            else a = ld_split4(ssp, s_hl, ((long long)fi * R + ra) * s_cp + (q - pq) * 4);
          } else {
            if (q < pq) a = ldg4(p + ((int64_t)fi * R + ra) * p_cp + q * 4);
            else a = ldg4(skip + ((int64_t)fi * R + ra) * s_cp + (q - pq) * 4);
          }
          const float4 w0 = *reinterpret_cast<const float4*>(wt + q * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wt + q * 8 + 4);
          yr = fmaf(a.x, w0.x, yr); yi = fmaf(a.x, w0.y, yi);
          yr = fmaf(a.y, w0.z, yr); yi = fmaf(a.y, w0.w, yi);
          yr = fmaf(a.z, w1.x, yr); yi = fmaf(a.z, w1.y, yi);
          yr = fmaf(a.w, w1.z, yr); yi = fmaf(a.w, w1.w, yi);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      yr += __shfl_xor_sync(0xffffffffu, yr, o);
      yi += __shfl_xor_sync(0xffffffffu, yi, o);
    }
    if (lane == 0) {
      yr = prelu_f(yr + b_r, slope);
      yi = prelu_f(yi + b_i, slope);
      if (mask) {
        // model/pvae_module.py:L2594-2609 — operation order kept
        const float mag = tanhf(sqrtf(yr * yr + yi * yi));
        const float ph = atan2f(yi / (mag + 1e-8f), yr / (mag + 1e-8f));
        const float2 X = __ldg(reinterpret_cast<const float2*>(stft_x + ((int64_t)(b * Fout + fo) * T + t) * 2));
        const float in_mag = sqrtf(X.x * X.x + X.y * X.y);
        const float in_ph = atan2f(X.y, X.x);
        float s, c;
        sincosf(in_ph + ph, &s, &c);
        const float g = in_mag * mag;
        yr = g * c;
        yi = g * s;
      }
      const int bo = b * out_bmul + out_boff;
      *reinterpret_cast<float2*>(predict + ((int64_t)(bo * Fout + fo) * T + t) * 2) = make_float2(yr, yi);
    }
  }
}

}  // namespace idv

extern "C" int idv_enc0_fwd(const float* stft, int B, int Fin, int T, const float* w, const float* bias,
                            int Cout, float prelu_slope, void* out, int out_split, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(stft && w && bias && out, "idv_enc0_fwd: null pointer");
  IDV_CHECK_ARG(B > 0 && Fin >= 5 && T > 0, "idv_enc0_fwd: bad shape B=%d Fin=%d T=%d", B, Fin, T);
  IDV_CHECK_ARG(Cout > 0 && Cout % 16 == 0, "idv_enc0_fwd: Cout=%d must be a multiple of 16", Cout);
  const int N = 2 * Cout;
  const int Fout = (Fin + 4 - 5) / 2 + 1;
  IDV_CHECK_ARG(Fout <= 65535, "idv_enc0_fwd: Fout too large");
  const int R = B * (T + 1);
  const size_t smem = (size_t)21 * N * sizeof(float);
  IDV_CUDA(cudaFuncSetAttribute(enc0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(R, 32), Fout);
  enc0_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(stft, B, Fin, T, w, bias, N, prelu_slope, out, Fout,
                                                         out_split);
  IDV_LAUNCH_CHECK("enc0_kernel");
  return IDV_OK;
}

extern "C" int idv_dec5_head_fwd(const void* p, int p_cp, const void* skip, int s_cp, int in_split, int NB, int Fin, int T,
                                 const float* w, const float* bias, float prelu_slope, int mask,
                                 const float* stft_x, float* predict, int out_bmul, int out_boff,
                                 void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(p && w && bias && predict, "idv_dec5_head_fwd: null pointer");
  IDV_CHECK_ARG(!mask || stft_x, "idv_dec5_head_fwd: mask head needs stft_x");
  if (!skip) s_cp = 0;
  IDV_CHECK_ARG(p_cp > 0 && p_cp % 4 == 0 && s_cp % 4 == 0, "idv_dec5_head_fwd: channel counts must be multiples of 4");
  IDV_CHECK_ARG(NB > 0 && Fin > 0 && T > 0 && 2 * Fin - 1 <= 65535, "idv_dec5_head_fwd: bad shape");
  const int R = NB * (T + 1);
  const size_t smem = (size_t)10 * (p_cp + s_cp) * 2 * sizeof(float);
  IDV_CHECK_ARG(smem <= 200 * 1024, "idv_dec5_head_fwd: too many input channels");
  IDV_CUDA(cudaFuncSetAttribute(dec5_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(R, D5_ROWS), 2 * Fin - 1);
  dec5_head_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p, p_cp, skip, s_cp, in_split, NB, Fin, T, w, bias,
                                                             prelu_slope, mask, stft_x, predict, out_bmul,
                                                             out_boff);
  IDV_LAUNCH_CHECK("dec5_head_kernel");
  return IDV_OK;
}
