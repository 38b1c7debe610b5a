// STFT as fused reflect-pad + framing + windowed dense DFT, iSTFT as inverse-DFT frames + overlap-add.
// fp32 SIMT version (the two transforms are 0.4 % of the path's FLOPs).
//   stft : model/pvae_module.py:L21-27 (torch.stft)      -> SURVEY §9 V1
//   istft: model/pvae_module.py:L38-42 (torch.istft)     -> SURVEY §9 V2
#include "idv_common.cuh"

namespace idv {

constexpr int ST_FR = 64;        // frames per CTA
constexpr int ST_BN = 128;       // output columns per CTA
constexpr int ST_BK = 16;
constexpr int ST_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// out[b][kbin][t][part] = sum_{j<win} xp[b][hop*t + off + j] * basis[j][2*kbin+part]
// basis is [win][ncol_pad] (ncol_pad multiple of ST_BN, zero padded).
__global__ void __launch_bounds__(ST_THREADS) stft_kernel(const float* __restrict__ x, int L, int T,
                                                          const float* __restrict__ basis, int ncol,
                                                          int ncol_pad, int n_fft, int hop, int win,
                                                          float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* seg = smem;                                   // (ST_FR-1)*hop + win floats
  float* Bs = smem + ((ST_FR - 1) * hop + win + 3) / 4 * 4;   // [ST_BK][ST_BN]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int t0 = blockIdx.x * ST_FR;
  const int b = blockIdx.y;
  const int n0 = blockIdx.z * ST_BN;
  const int off = (n_fft - win) / 2;
  const int half = n_fft / 2;
  const int seg_len = (ST_FR - 1) * hop + win;
  const float* xb = x + (int64_t)b * L;
  for (int i = tid; i < seg_len; i += ST_THREADS) {
    int idx = hop * t0 + off + i - half;               // index into the un-padded signal
    if (idx < 0) idx = -idx;
    if (idx >= L) idx = 2 * (L - 1) - idx;
    seg[i] = (idx >= 0 && idx < L) ? __ldg(xb + idx) : 0.f;
  }
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  __syncthreads();
  for (int k0 = 0; k0 < win; k0 += ST_BK) {
    // B tile: 16 x 128 floats = 512 float4, 2 per thread
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * ST_THREADS;
      const int k = idx >> 5, nq = idx & 31;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < win) v = ldg4(basis + (int64_t)(k0 + k) * ncol_pad + n0 + nq * 4);
      *reinterpret_cast<float4*>(&Bs[k * ST_BN + nq * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ST_BK; ++k) {
      float a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int si = hop * (ty * 4 + i) + k0 + k;
        a[i] = (si < seg_len) ? seg[si] : 0.f;
      }
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k * ST_BN + tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k * ST_BN + 64 + tx * 4]);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int nbins = ncol / 2;
#pragma unroll
  for (int cg = 0; cg < 2; ++cg) {
#pragma unroll
    for (int jb = 0; jb < 2; ++jb) {
      const int n = n0 + cg * 64 + tx * 4 + jb * 2;    // even column -> (bin, re)
      const int kb = n >> 1;
      if (kb >= nbins) continue;
      float* op = out + ((int64_t)(b * nbins + kb) * T) * 2;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty * 4 + i;
        if (t < T) {
          *reinterpret_cast<float2*>(op + (int64_t)t * 2) =
              make_float2(acc[i][cg * 4 + jb * 2], acc[i][cg * 4 + jb * 2 + 1]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// frames[(b*T + t)][n'] = sum_{k'} spec[b][k'/2][t][k'%2] * basis[k'][n'],  k' < 2*nbins
// basis is [kpad][ncol_pad] zero padded (kpad multiple of 16, ncol_pad multiple of 128).
__global__ void __launch_bounds__(ST_THREADS) istft_frames_kernel(const float* __restrict__ spec, int T,
                                                                  int nbins, const float* __restrict__ basis,
                                                                  int kpad, int ncol_pad, int win,
                                                                  float* __restrict__ frames) {
  __shared__ __align__(16) float As[ST_BK][ST_FR];
  __shared__ __align__(16) float Bs[ST_BK][ST_BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int t0 = blockIdx.x * ST_FR;
  const int b = blockIdx.y;
  const int n0 = blockIdx.z * ST_BN;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < kpad; k0 += ST_BK) {
    // A tile: 8 bins x 64 frames of (re,im) float2 -> 512 float2, 2 per thread
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * ST_THREADS;
      const int kb = idx >> 6, tt = idx & 63;
      const int bin = (k0 >> 1) + kb;
      float2 v = make_float2(0.f, 0.f);
      if (bin < nbins && t0 + tt < T)
        v = __ldg(reinterpret_cast<const float2*>(spec + ((int64_t)(b * nbins + bin) * T + t0 + tt) * 2));
      As[2 * kb][tt] = v.x;
      As[2 * kb + 1][tt] = v.y;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * ST_THREADS;
      const int k = idx >> 5, nq = idx & 31;
      *reinterpret_cast<float4*>(&Bs[k][nq * 4]) = ldg4(basis + (int64_t)(k0 + k) * ncol_pad + n0 + nq * 4);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ST_BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int cg = 0; cg < 2; ++cg) {
    const int n = n0 + cg * 64 + tx * 4;
    if (n >= win) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + ty * 4 + i;
      if (t < T)
        *reinterpret_cast<float4*>(frames + ((int64_t)b * T + t) * win + n) =
            make_float4(acc[i][cg * 4 + 0], acc[i][cg * 4 + 1], acc[i][cg * 4 + 2], acc[i][cg * 4 + 3]);
    }
  }
}

// overlap-add + window-envelope normalisation + centre trim (HBM-bound)
__global__ void __launch_bounds__(256) ola_kernel(const float* __restrict__ frames, int frame_ld,
                                                  const float* __restrict__ wsq, int T, int n_fft, int hop, int win,
                                                  int out_len, const int* __restrict__ lengths,
                                                  float* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (s >= out_len) return;
  const int T_all = T;
  if (lengths) {                                     // ragged batch: utterance b has lengths[b] / hop + 1 frames
    T = lengths[b] / hop + 1;
    if (T > T_all) T = T_all;
    if (s >= hop * (T - 1)) {
      out[(int64_t)b * out_len + s] = 0.f;
      return;
    }
  }
  const int off = (n_fft - win) / 2;
  const int q = s + n_fft / 2 - off;                 // position relative to the first window tap
  int t_hi = q / hop;
  int t_lo = (q - win + hop) / hop;                  // ceil((q - win + 1)/hop) for q-win+1 >= 0
  if (q - win + 1 <= 0) t_lo = 0;
  if (t_hi > T - 1) t_hi = T - 1;
  float acc = 0.f, env = 0.f;
  const float* fb = frames + (int64_t)b * T_all * frame_ld;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int j = q - hop * t;
    if (j >= 0 && j < win) {
      acc += __ldg(fb + (int64_t)t * frame_ld + j);
      env += __ldg(wsq + j);
    }
  }
  out[(int64_t)b * out_len + s] = acc / env;
}

// ---- tensor-core path helpers: operands of the two DFT GEMMs in split-bf16, K-major -----------------------
// frames[(b*T + t)][j] = xp[b][hop*t + off + j] for j < win, 0 for win <= j < kpad   (hi / lo planes)
__global__ void __launch_bounds__(256) stft_frames_split_kernel(const float* __restrict__ x, int B, int L, int T,
                                                                int n_fft, int hop, int win, int kpad,
                                                                const int* __restrict__ lengths,
                                                                unsigned short* __restrict__ out) {
  const long long n = (long long)B * T * (kpad / 4);
  const long long hl = (long long)B * T * kpad;
  const int off = (n_fft - win) / 2, half = n_fft / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j4 = (int)(i % (kpad / 4)) * 4;
    const long long bt = i / (kpad / 4);
    const int t = (int)(bt % T), b = (int)(bt / T);
    // ragged batch: utterance b holds lengths[b] samples (reflect padding at ITS end), frames beyond its last are 0
    const int Lb = lengths ? min(lengths[b], L) : L;
    const bool live = t <= Lb / hop;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j4 + e;
      int idx = hop * t + off + j - half;
      if (idx < 0) idx = -idx;
      if (idx >= Lb) idx = 2 * (Lb - 1) - idx;
      v[e] = (live && j < win && idx >= 0 && idx < Lb) ? __ldg(x + (long long)b * L + idx) : 0.f;
    }
    st_split4(out, hl, bt * kpad + j4, make_float4(v[0], v[1], v[2], v[3]));
  }
}

// rows[(b*T + t)][2k + part] = spec[b][k][t][part], zero padded to kpad columns (hi / lo planes)
// grid (ceil(T/32), ceil(nbins/32), B), block (32, 8)
__global__ void __launch_bounds__(256) spec_rows_split_kernel(const float* __restrict__ spec, int B, int nbins, int T,
                                                              int kpad, unsigned short* __restrict__ out) {
  __shared__ float tile[2][32][33];
  const int t0 = blockIdx.x * 32, k0 = blockIdx.y * 32, b = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int k = k0 + i, t = t0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (k < nbins && t < T) v = __ldg(reinterpret_cast<const float2*>(spec + ((long long)(b * nbins + k) * T + t) * 2));
    tile[0][i][tx] = v.x;
    tile[1][i][tx] = v.y;
  }
  __syncthreads();
  const long long hl = (long long)B * T * kpad;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, k = k0 + tx;
    if (t < T && 2 * k + 1 < kpad) {
      const long long idx = ((long long)b * T + t) * kpad + 2 * k;
      unsigned short h0, l0, h1, l1;
      split_bf16(tile[0][tx][i], h0, l0);
      split_bf16(tile[1][tx][i], h1, l1);
      *reinterpret_cast<unsigned int*>(out + idx) = (unsigned)h0 | ((unsigned)h1 << 16);
      *reinterpret_cast<unsigned int*>(out + hl + idx) = (unsigned)l0 | ((unsigned)l1 << 16);
    }
  }
}

}  // namespace idv

extern "C" int idv_stft_frames_split(const float* x, int B, int L, int n_fft, int hop, int win, int kpad,
                                     const int* lengths, void* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(x && out && B > 0 && hop > 0 && win > 0 && n_fft >= win && kpad >= win && kpad % 64 == 0,
                "idv_stft_frames_split: bad argument");
  IDV_CHECK_ARG(L > n_fft / 2, "idv_stft_frames_split: reflect padding needs L > n_fft/2 (L=%d)", L);
  const int T = L / hop + 1;
  const long long n = (long long)B * T * (kpad / 4);
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  stft_frames_split_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, B, L, T, n_fft, hop, win, kpad, lengths,
                                                                      reinterpret_cast<unsigned short*>(out));
  IDV_LAUNCH_CHECK("stft_frames_split_kernel");
  return IDV_OK;
}

extern "C" int idv_spec_rows_split(const float* spec, int B, int nbins, int T, int kpad, void* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(spec && out && B > 0 && B <= 65535 && nbins > 0 && T > 0 && kpad % 64 == 0 && kpad >= 2 * nbins,
                "idv_spec_rows_split: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  // zero the K padding (columns 2*nbins .. kpad) once: the kernel writes every (row, 2k) pair with 2k+1 < kpad,
  // covering bins up to kpad/2 with zeros beyond nbins
  dim3 grid(cdiv(T, 32), cdiv(kpad / 2, 32), B), block(32, 8);
  spec_rows_split_kernel<<<grid, block, 0, st>>>(spec, B, nbins, T, kpad, reinterpret_cast<unsigned short*>(out));
  IDV_LAUNCH_CHECK("spec_rows_split_kernel");
  return IDV_OK;
}

extern "C" int idv_ola_fwd(const float* frames, int frame_ld, const float* wsq, int B, int T, int n_fft, int hop,
                           int win, const int* lengths, float* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(frames && wsq && out && B > 0 && B <= 65535 && T > 1 && frame_ld >= win, "idv_ola_fwd: bad argument");
  const int out_len = hop * (T - 1);
  dim3 g2(cdiv(out_len, 256), B);
  ola_kernel<<<g2, 256, 0, (cudaStream_t)stream>>>(frames, frame_ld, wsq, T, n_fft, hop, win, out_len, lengths, out);
  IDV_LAUNCH_CHECK("ola_kernel");
  return IDV_OK;
}

extern "C" int idv_stft_fwd(const float* x, int B, int L, const float* basis, int n_fft, int hop, int win,
                            float* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(x && basis && out, "idv_stft_fwd: null pointer");
  IDV_CHECK_ARG(B > 0 && B <= 65535, "idv_stft_fwd: batch %d out of range", B);
  IDV_CHECK_ARG(n_fft >= win && hop > 0 && hop <= 128 && win <= 512 && win % 4 == 0 && ((n_fft - win) % 2 == 0),
                "idv_stft_fwd: unsupported n_fft=%d hop=%d win=%d", n_fft, hop, win);
  IDV_CHECK_ARG(L > n_fft / 2, "idv_stft_fwd: reflect padding needs L > n_fft/2 (L=%d)", L);
  const int T = L / hop + 1;
  const int ncol = 2 * (n_fft / 2 + 1);
  const int ncol_pad = cdiv(ncol, ST_BN) * ST_BN;
  const int seg_len = (ST_FR - 1) * hop + win;
  const size_t smem = (size_t)((seg_len + 3) / 4 * 4 + ST_BK * ST_BN) * sizeof(float);
  IDV_CUDA(cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(T, ST_FR), B, ncol_pad / ST_BN);
  stft_kernel<<<grid, ST_THREADS, smem, (cudaStream_t)stream>>>(x, L, T, basis, ncol, ncol_pad, n_fft, hop, win, out);
  IDV_LAUNCH_CHECK("stft_kernel");
  return IDV_OK;
}

extern "C" int idv_istft_fwd(const float* spec, int B, int T, const float* basis, const float* wsq, int n_fft,
                             int hop, int win, float* frames, float* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(spec && basis && wsq && frames && out, "idv_istft_fwd: null pointer");
  IDV_CHECK_ARG(B > 0 && B <= 65535 && T > 1, "idv_istft_fwd: B=%d T=%d out of range", B, T);
  IDV_CHECK_ARG(n_fft >= win && hop > 0 && win % 4 == 0, "idv_istft_fwd: unsupported n_fft=%d hop=%d win=%d", n_fft, hop, win);
  const int nbins = n_fft / 2 + 1;
  const int kpad = cdiv(2 * nbins, ST_BK) * ST_BK;
  const int ncol_pad = cdiv(win, ST_BN) * ST_BN;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(T, ST_FR), B, ncol_pad / ST_BN);
  istft_frames_kernel<<<grid, ST_THREADS, 0, st>>>(spec, T, nbins, basis, kpad, ncol_pad, win, frames);
  IDV_LAUNCH_CHECK("istft_frames_kernel");
  const int out_len = hop * (T - 1);
  dim3 g2(cdiv(out_len, 256), B);
  ola_kernel<<<g2, 256, 0, st>>>(frames, win, wsq, T, n_fft, hop, win, out_len, nullptr, out);
  IDV_LAUNCH_CHECK("ola_kernel");
  return IDV_OK;
}
