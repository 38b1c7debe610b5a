// The two layers that are not GEMM-shaped, written as HBM-bound SIMT kernels:
//   enc0: causal ComplexConv2d(1 -> Cout) + folded CBN + PReLU, reading the user-layout STFT
//   dec5: causal ComplexConvTranspose2d(Cin -> 1) + folded CBN + PReLU (+ mask head), writing the
//         user-layout spectrum `predict`.
#include "idv_common.cuh"

namespace idv {

// grid: (row tiles of 128, Fout); block 256 = 32 row groups (4 consecutive rows each) x 8 channel groups.
// Each thread computes 4 rows x 8 consecutive channels per 64-channel slab: the 4 rows share the weight loads
// (2 LDS.128 per 32 FMA) and, being consecutive frames, 5 instead of 8 time samples of input.
constexpr int E0_ROWS = 128;
__global__ void __launch_bounds__(256, 2) enc0_kernel(const float* __restrict__ stft, int NB, int Fin, int T,
                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                   int N, float slope, void* __restrict__ outv, int Fout,
                                                   int out_split, int causal, int t_valid,
                                                   const float* __restrict__ prev, int keep_pad) {
  // weights [20][N] + bias [N]; inside every 64-channel slab the columns are stored as
  // [cg*4 + j | 32 + cg*4 + j] so that the two float4 of channel group cg are bank-conflict free
  extern __shared__ __align__(16) float ws[];
  const int tid = threadIdx.x;
  pdl_trigger();
  for (int i = tid; i < 21 * N; i += 256) {
    const int k = i / N, n = i % N;
    const int slab = n & ~63, c = n & 63;
    const int phys = slab + ((c & 4) ? 32 : 0) + (c >> 3) * 4 + (c & 3);
    ws[k * N + phys] = (k < 20) ? __ldg(w + i) : __ldg(bias + n);
  }
  pdl_wait();           // (the weights are launch-invariant; the spectrum is the predecessor's output)
  const int Tp = T + 1;
  const int R = NB * Tp;
  const int cg = tid & 7;
  const int fo = blockIdx.y;
  const long long hl = (long long)Fout * R * N;
  float* of32 = reinterpret_cast<float*>(outv);
  unsigned short* osp = reinterpret_cast<unsigned short*>(outv);
  float* xs0 = ws + 21 * N;                          // two im2col buffers [128 rows][21] (double buffered)
  const int n_tiles = (R + E0_ROWS - 1) / E0_ROWS;
  // im2col entry i = (row, tap) of a tile: xs[row][tap*2 + part]; pad rows, rows past R and out-of-range taps are 0.
  // Every thread owns entries tid + j*256, j < 5; the next tile's entries are fetched into registers before the
  // current tile is computed so the global-load latency hides behind the FMAs.
  float2 pre[5];
  auto fetch = [&](int tile) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int i = tid + j * 256;
      const int rl = i / 10, tap = i % 10;
      const int kf = tap >> 1, kt = tap & 1;
      const int r = tile * E0_ROWS + rl;
      float2 v = make_float2(0.f, 0.f);
      if (r < R) {
        const int b = r / Tp, t = r % Tp - 1;
        const int fi = 2 * fo + kf - 2, ti = causal ? t - 1 + kt : t + kt;   // non-causal: x[t], x[t+1]
        if (t >= 0 && fi >= 0 && fi < Fin && ti >= 0 && ti < T)
          v = __ldg(reinterpret_cast<const float2*>(stft + ((int64_t)(b * Fin + fi) * T + ti) * 2));
        else if (prev && t >= 0 && fi >= 0 && fi < Fin && ti == -1)       // streaming: x[-1] = last frame of the
          v = __ldg(reinterpret_cast<const float2*>(prev + (int64_t)(b * Fin + fi) * 2));   // previous step
      }
      pre[j] = v;
    }
  };
  auto stash = [&](float* xs) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int i = tid + j * 256;
      const int rl = i / 10, tap = i % 10;
      xs[rl * 21 + tap * 2] = pre[j].x;
      xs[rl * 21 + tap * 2 + 1] = pre[j].y;
    }
  };
  int cur = 0;
  if ((int)blockIdx.x < n_tiles) {
    fetch(blockIdx.x);
    stash(xs0);
  }
  __syncthreads();                                   // weights + first tile visible
  // each block walks over row tiles blockIdx.x, blockIdx.x + gridDim.x, ... (weights staged once)
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, cur ^= 1) {
  const float* xs = xs0 + cur * (E0_ROWS * 21);
  const int next = tile + gridDim.x;
  if (next < n_tiles) fetch(next);
  const int r0 = tile * E0_ROWS + (tid >> 3) * 4;
  if (r0 < R) {
  bool ok[4], pad[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ok[i] = r0 + i < R;
    pad[i] = ((r0 + i) % Tp) == 0 || (t_valid > 0 && ((r0 + i) % Tp) > t_valid);
  }
  const float* xrow = xs + (tid >> 3) * 4 * 21;
  for (int n0 = 0; n0 < N; n0 += 64) {
    float acc[4][8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(&ws[20 * N + n0 + cg * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&ws[20 * N + n0 + 32 + cg * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = b0.x; acc[i][1] = b0.y; acc[i][2] = b0.z; acc[i][3] = b0.w;
        acc[i][4] = b1.x; acc[i][5] = b1.y; acc[i][6] = b1.z; acc[i][7] = b1.w;
      }
    }
#pragma unroll 4
    for (int k = 0; k < 20; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(&ws[k * N + n0 + cg * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&ws[k * N + n0 + 32 + cg * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = xrow[i * 21 + k];
        acc[i][0] = fmaf(x, w0.x, acc[i][0]); acc[i][1] = fmaf(x, w0.y, acc[i][1]);
        acc[i][2] = fmaf(x, w0.z, acc[i][2]); acc[i][3] = fmaf(x, w0.w, acc[i][3]);
        acc[i][4] = fmaf(x, w1.x, acc[i][4]); acc[i][5] = fmaf(x, w1.y, acc[i][5]);
        acc[i][6] = fmaf(x, w1.z, acc[i][6]); acc[i][7] = fmaf(x, w1.w, acc[i][7]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!ok[i] || (keep_pad && ((r0 + i) % Tp) == 0)) continue;      // keep_pad: pad rows carry state
      const long long oidx = ((long long)fo * R + r0 + i) * N + n0 + cg * 8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = pad[i] ? 0.f : prelu_f(acc[i][j], slope);
      if (out_split) {
        unsigned short h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split_bf16(v[j], h[j], l[j]);
        *reinterpret_cast<uint4*>(osp + oidx) =
            make_uint4((unsigned)h[0] | ((unsigned)h[1] << 16), (unsigned)h[2] | ((unsigned)h[3] << 16),
                       (unsigned)h[4] | ((unsigned)h[5] << 16), (unsigned)h[6] | ((unsigned)h[7] << 16));
        *reinterpret_cast<uint4*>(osp + hl + oidx) =
            make_uint4((unsigned)l[0] | ((unsigned)l[1] << 16), (unsigned)l[2] | ((unsigned)l[3] << 16),
                       (unsigned)l[4] | ((unsigned)l[5] << 16), (unsigned)l[6] | ((unsigned)l[7] << 16));
      } else {
        *reinterpret_cast<float4*>(of32 + oidx) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(of32 + oidx + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  }   // r0 < R
  if (next < n_tiles) stash(xs0 + (cur ^ 1) * (E0_ROWS * 21));
  __syncthreads();
  }   // tile loop
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
                            int Cout, float prelu_slope, void* out, int out_split, int causal, int t_valid,
                            const float* prev, int keep_pad, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(stft && w && bias && out, "idv_enc0_fwd: null pointer");
  IDV_CHECK_ARG(B > 0 && Fin >= 5 && T > 0, "idv_enc0_fwd: bad shape B=%d Fin=%d T=%d", B, Fin, T);
  IDV_CHECK_ARG(Cout > 0 && Cout % 32 == 0, "idv_enc0_fwd: Cout=%d must be a multiple of 32", Cout);
  const int N = 2 * Cout;
  const int Fout = (Fin + 4 - 5) / 2 + 1;
  IDV_CHECK_ARG(Fout <= 65535, "idv_enc0_fwd: Fout too large");
  const int R = B * (T + 1);
  const size_t smem = (size_t)(21 * N + 2 * E0_ROWS * 21) * sizeof(float);
  IDV_CUDA(cudaFuncSetAttribute(enc0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_tiles = cdiv(R, E0_ROWS);
  dim3 grid(n_tiles < 16 ? n_tiles : cdiv(n_tiles, 8), Fout);      // ~8 row tiles per block
  IDV_CUDA(launch_pdl(enc0_kernel, grid, dim3(256), smem, (cudaStream_t)stream, stft, B, Fin, T, w, bias, N, prelu_slope, out,
                      Fout, out_split, causal, t_valid, prev, keep_pad));
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
