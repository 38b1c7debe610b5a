// Backward of the NSVAE encoder (phase-1 training step, /root/reference i_dccrn_vae/nsvae_dccrn/train_nsvae.py:L505-566:
// noisy encoder train=True -> closed-form KL -> loss.backward() -> Adam).  The GEMM-shaped parts of the backward
// (data gradients of the complex convs and LSTM projections, all weight gradients) run on the tcgen05 tap-GEMM
// (csrc/tapgemm_tc.cu) - weight gradients as GEMMs over the ROW dimension on transposed copies of the activations;
// this file holds the HBM-bound pieces around them:
//   planes_transpose_split : [F][R][Cp] -> split-bf16 [2][F][Cp][Rpad] (optionally row-shifted), the K-major operands
//                            of the weight-gradient GEMMs
//   cbn_bwd_*              : ComplexBatchNormal(train=True) + PReLU backward (model/complex_progress.py:L131-209,
//                            differentiated through the batch mean and covariance like the reference's autograd)
//   lstm_scan_c / lstm_cell_bwd_step / lstm_combine_bwd / colsum : BPTT of nn.LSTM (complex_progress.py:L58-74)
//   enc0_wgrad             : weight gradient of the Cin = 1 first layer
//   adam_step              : torch.optim.Adam(lr, weight_decay) update (train_nsvae.py:L200)
#include "idv_common.cuh"

namespace idv {

__device__ __forceinline__ int r8(int c) { return (c + 7) & ~7; }
__device__ __forceinline__ float sgm(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float ld_act(const void* p, int split, long long hl, long long idx) {
  return split ? ld_split1(reinterpret_cast<const unsigned short*>(p), hl, idx)
               : __ldg(reinterpret_cast<const float*>(p) + idx);
}

// out[f][c][k] = in[f][k + shift][c] for 0 <= k + shift < R, else 0;  k < Rpad.  grid (Rpad/64, ceil(Cp/64), F), block 256.
// 64 x 64 tiles; channel pairs are read and row pairs are written as 32-bit words (Cp even, Rpad % 64 == 0).
__global__ void __launch_bounds__(256) planes_transpose_split_kernel(const void* __restrict__ in, int in_split, int F,
                                                                     int R, int Cp, int Rpad, int shift,
                                                                     unsigned short* __restrict__ out) {
  __shared__ float tile[64][65];
  const int k0 = blockIdx.x * 64, c0 = blockIdx.y * 64, f = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long hl_in = (long long)F * R * Cp, hl_out = (long long)F * Cp * Rpad;
  const float* inf = reinterpret_cast<const float*>(in);
  const unsigned short* ins = reinterpret_cast<const unsigned short*>(in);
  for (int i = ty; i < 64; i += 8) {
    const int r = k0 + i + shift, c = c0 + 2 * tx;
    float v0 = 0.f, v1 = 0.f;
    if (r >= 0 && r < R && c < Cp) {
      const long long idx = ((long long)f * R + r) * Cp + c;
      if (in_split) {
        const unsigned h = __ldg(reinterpret_cast<const unsigned*>(ins + idx));
        const unsigned l = __ldg(reinterpret_cast<const unsigned*>(ins + hl_in + idx));
        v0 = __uint_as_float(h << 16) + __uint_as_float(l << 16);
        v1 = __uint_as_float(h & 0xffff0000u) + __uint_as_float(l & 0xffff0000u);
      } else {
        const float2 v = __ldg(reinterpret_cast<const float2*>(inf + idx));
        v0 = v.x; v1 = v.y;
      }
    }
    tile[i][2 * tx] = v0;
    tile[i][2 * tx + 1] = v1;
  }
  __syncthreads();
  for (int j = ty; j < 64; j += 8) {
    const int c = c0 + j, k = k0 + 2 * tx;
    if (c < Cp) {                                                    // k + 1 < Rpad: Rpad is a multiple of 64
      unsigned short h0, l0, h1, l1;
      split_bf16(tile[2 * tx][j], h0, l0);
      split_bf16(tile[2 * tx + 1][j], h1, l1);
      const long long o = ((long long)f * Cp + c) * Rpad + k;
      *reinterpret_cast<unsigned*>(out + o) = (unsigned)h0 | ((unsigned)h1 << 16);
      *reinterpret_cast<unsigned*>(out + hl_out + o) = (unsigned)l0 | ((unsigned)l1 << 16);
    }
  }
}

__global__ void __launch_bounds__(256) f32_to_split_kernel(const float* __restrict__ x, long long n4,
                                                           unsigned short* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    st_split4(out, n4 * 4, i * 4, ldg4(x + i * 4));
}

// ---- ComplexBatchNormal(train) + PReLU backward --------------------------------------------------------------------
// pre = Z y + b' (b' = beta - Z mu), act = PReLU(pre).  gp = g * (pre > 0 ? 1 : slope), xc = y - mu.
// acc[c][8] (double): sum gp_r, gp_i, gp_r xc_r, gp_r xc_i, gp_i xc_r, gp_i xc_i, sum g*pre over pre <= 0, (unused)
// grid (F, chunks), block (round32(C), 256 / round32(C)): thread = (complex channel, row lane)
__global__ void __launch_bounds__(256) cbn_bwd_reduce_kernel(const void* __restrict__ y, int y_split,
                                                             const void* __restrict__ g, int g_split, int NB, int C,
                                                             int F, int T, int Tv, int rows_per_chunk,
                                                             const float* __restrict__ stats, const float* __restrict__ zb,
                                                             float slope, double* __restrict__ acc) {
  __shared__ double red[256];
  const int c = threadIdx.x, ly = threadIdx.y, ny = blockDim.y, Cw = blockDim.x;
  const int Ch = r8(C), Cp = 2 * Ch, Tp = T + 1;
  const long long R = (long long)NB * Tp, hl = (long long)F * R * Cp;
  const int f = blockIdx.x;
  long long r_begin = (long long)blockIdx.y * rows_per_chunk, r_end = r_begin + rows_per_chunk;
  if (r_end > R) r_end = R;
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  if (c < C) {
    const float mu_r = stats[c * 5 + 0], mu_i = stats[c * 5 + 1];
    const float zrr = zb[c * 6 + 0], zri = zb[c * 6 + 1], zir = zb[c * 6 + 2], zii = zb[c * 6 + 3];
    const float br = zb[c * 6 + 4], bi = zb[c * 6 + 5];
    for (long long r = r_begin + ly; r < r_end; r += ny) {
      const int tt = (int)(r % Tp);
      if (tt == 0 || tt > Tv) continue;
      const long long idx = ((long long)f * R + r) * Cp;
      const float yr = ld_act(y, y_split, hl, idx + c), yi = ld_act(y, y_split, hl, idx + Ch + c);
      const float gr = ld_act(g, g_split, hl, idx + c), gi = ld_act(g, g_split, hl, idx + Ch + c);
      const float pr = fmaf(zrr, yr, fmaf(zri, yi, br)), pi = fmaf(zir, yr, fmaf(zii, yi, bi));
      const float gpr = pr > 0.f ? gr : slope * gr, gpi = pi > 0.f ? gi : slope * gi;
      const float xr = yr - mu_r, xi = yi - mu_i;
      s[0] += gpr; s[1] += gpi;
      s[2] += (double)gpr * xr; s[3] += (double)gpr * xi; s[4] += (double)gpi * xr; s[5] += (double)gpi * xi;
      s[6] += (pr > 0.f ? 0.0 : (double)gr * pr) + (pi > 0.f ? 0.0 : (double)gi * pi);
    }
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    red[ly * Cw + c] = s[k];
    __syncthreads();
    if (ly == 0 && c < C) {
      double t = 0.0;
      for (int q = 0; q < ny; ++q) t += red[q * Cw + c];
      atomicAdd(acc + c * 8 + k, t);
    }
    __syncthreads();
  }
}

// per channel: parameter gradients (+=) and the coefficients of the element-wise pass
//   dy = Z^T gp + D xc - m,  coef[c][10] = Zrr, Zir, Zri, Zii (Z^T rows), Drr, Dri, Dii, m_r, m_i, (unused)
__global__ void cbn_bwd_finalize_kernel(const double* __restrict__ acc, double n, int C, const float* __restrict__ stats,
                                        const float* __restrict__ g_rr, const float* __restrict__ g_ri,
                                        const float* __restrict__ g_ii, float* __restrict__ coef,
                                        float* __restrict__ d_grr, float* __restrict__ d_gri, float* __restrict__ d_gii,
                                        float* __restrict__ d_br, float* __restrict__ d_bi,
                                        double* __restrict__ d_slope) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double slope_part = 0.0;
  if (c < C) {
    const double eps = 1e-5;
    const double a = stats[c * 5 + 2], cri = stats[c * 5 + 3], b = stats[c * 5 + 4];      // Vrr, Vri, Vii (eps inside)
    const double* s = acc + c * 8;
    // forward quantities (model/complex_progress.py:L168-200)
    double delta = a * b - cri * cri + eps;
    const bool clamped = delta < 1e-8;
    if (clamped) delta = 1e-8;
    const double sq = sqrt(delta), t = sqrt(a + b + 2 * sq + eps), q = sq * t + eps;
    const double wrr = (b + sq) / q, wii = (a + sq) / q, wri = -cri / q;
    const double grr = g_rr[c], gri = g_ri[c], gii = g_ii[c];
    const double zrr = grr * wrr + gri * wri, zri = grr * wri + gri * wii;
    const double zir = gri * wrr + gii * wri, zii = gri * wri + gii * wii;
    // dZ_ab = sum gp_a xc_b
    const double dzrr = s[2], dzri = s[3], dzir = s[4], dzii = s[5];
    d_grr[c] += (float)(dzrr * wrr + dzri * wri);
    d_gri[c] += (float)(dzrr * wri + dzri * wii + dzir * wrr + dzii * wri);
    d_gii[c] += (float)(dzir * wri + dzii * wii);
    d_br[c] += (float)s[0];
    d_bi[c] += (float)s[1];
    const double dwrr = dzrr * grr + dzir * gri;
    const double dwri = dzrr * gri + dzri * grr + dzir * gii + dzii * gri;
    const double dwii = dzri * gri + dzii * gii;
    // W(V): Wrr = (b + s)/q, Wii = (a + s)/q, Wri = -c/q, q = s t + eps, t = sqrt(a + b + 2 s + eps), s = sqrt(delta)
    double ga = dwii / q, gb = dwrr / q, gc = -dwri / q;
    double gs = (dwrr + dwii) / q;
    const double gq = -(dwrr * (b + sq) + dwii * (a + sq) - dwri * cri) / (q * q);
    gs += gq * t;
    const double gt = gq * sq;
    const double gin = gt / (2 * t);
    ga += gin; gb += gin; gs += 2 * gin;
    const double gdelta = clamped ? 0.0 : gs / (2 * sq);
    ga += gdelta * b; gb += gdelta * a; gc += -2 * cri * gdelta;
    // V = mean(xc xc^T): d/dxc_r = (2 ga xc_r + gc xc_i)/n, d/dxc_i = (2 gb xc_i + gc xc_r)/n
    float* k = coef + c * 10;
    k[0] = (float)zrr; k[1] = (float)zir; k[2] = (float)zri; k[3] = (float)zii;
    k[4] = (float)(2 * ga / n); k[5] = (float)(gc / n); k[6] = (float)(2 * gb / n);
    k[7] = (float)((zrr * s[0] + zir * s[1]) / n);
    k[8] = (float)((zri * s[0] + zii * s[1]) / n);
    k[9] = 0.f;
    slope_part = s[6];
  }
  // one PReLU slope for the whole layer
  for (int o = 16; o > 0; o >>= 1) slope_part += __shfl_down_sync(0xffffffffu, slope_part, o);
  if ((threadIdx.x & 31) == 0 && d_slope) atomicAdd(d_slope, slope_part);
}

// dy (planes, pad rows and invalid frames = 0), fp32 or split
__global__ void __launch_bounds__(256) cbn_bwd_apply_kernel(const void* __restrict__ y, int y_split,
                                                            const void* __restrict__ g, int g_split, int NB, int C,
                                                            int F, int T, int Tv, const float* __restrict__ stats,
                                                            const float* __restrict__ zb, const float* __restrict__ coef,
                                                            float slope, void* __restrict__ dy, int dy_split) {
  const int Ch = r8(C), Cp = 2 * Ch, Tp = T + 1;
  const long long R = (long long)NB * Tp, hl = (long long)F * R * Cp;
  const long long n = (long long)F * R * Ch;
  float* d32 = reinterpret_cast<float*>(dy);
  unsigned short* dsp = reinterpret_cast<unsigned short*>(dy);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Ch);
    const long long fr = i / Ch;
    const int tt = (int)((fr % R) % Tp);
    const long long idx = fr * Cp;
    float dr = 0.f, di = 0.f;
    if (c < C && tt != 0 && tt <= Tv) {
      const float yr = ld_act(y, y_split, hl, idx + c), yi = ld_act(y, y_split, hl, idx + Ch + c);
      const float gr = ld_act(g, g_split, hl, idx + c), gi = ld_act(g, g_split, hl, idx + Ch + c);
      const float* z = zb + c * 6;
      const float pr = fmaf(z[0], yr, fmaf(z[1], yi, z[4])), pi = fmaf(z[2], yr, fmaf(z[3], yi, z[5]));
      const float gpr = pr > 0.f ? gr : slope * gr, gpi = pi > 0.f ? gi : slope * gi;
      const float xr = yr - stats[c * 5 + 0], xi = yi - stats[c * 5 + 1];
      const float* k = coef + c * 10;
      dr = k[0] * gpr + k[1] * gpi + k[4] * xr + k[5] * xi - k[7];
      di = k[2] * gpr + k[3] * gpi + k[5] * xr + k[6] * xi - k[8];
    }
    if (dy_split) {
      st_split1(dsp, hl, idx + c, dr);
      st_split1(dsp, hl, idx + Ch + c, di);
    } else {
      d32[idx + c] = dr;
      d32[idx + Ch + c] = di;
    }
  }
}

// ---- LSTM backward -------------------------------------------------------------------------------------------------
// P: gate pre-activations [4][R][4H] (i, f, g, o); cst[4][R][H] <- c_t (pad rows 0).  One thread per (stream, b, j).
__global__ void __launch_bounds__(256) lstm_scan_c_kernel(const float* __restrict__ P, int NB, int T, int Tv, int H,
                                                          float* __restrict__ cst) {
  const long long n = 4LL * NB * H;
  const int Tp = T + 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % H);
    const long long sb = i / H;                                  // s*NB + b
    const float* p = P + sb * Tp * 4 * H;
    float* co = cst + sb * Tp * H;
    float c = 0.f;
    co[j] = 0.f;
    for (int t = 0; t < Tv; ++t) {
      const float* pt = p + (long long)(1 + t) * 4 * H;
      c = sgm(__ldg(pt + H + j)) * c + sgm(__ldg(pt + j)) * tanhf(__ldg(pt + 2 * H + j));
      co[(long long)(1 + t) * H + j] = c;
    }
  }
}

// one BPTT step (time t): dh = dH[t] + sum over the dh_parts split-K partial planes of dh_rec; dP[t] <- gate
// gradients; dc carried.
__global__ void __launch_bounds__(256) lstm_cell_bwd_step_kernel(const float* __restrict__ P,
                                                                 const float* __restrict__ cst,
                                                                 const float* __restrict__ dH,
                                                                 const float* __restrict__ dh_rec, float* __restrict__ dc,
                                                                 int NB, int T, int H, int t, int last, int dh_parts,
                                                                 float* __restrict__ dP,
                                                                 unsigned short* __restrict__ dP_step) {
  const long long n = 4LL * NB * H;
  const int Tp = T + 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % H);
    const long long sb = i / H;
    const long long row = sb * Tp + 1 + t;
    const float* pt = P + row * 4 * H;
    const float ig = sgm(__ldg(pt + j)), fg = sgm(__ldg(pt + H + j)), gg = tanhf(__ldg(pt + 2 * H + j)),
                og = sgm(__ldg(pt + 3 * H + j));
    const float c = __ldg(cst + row * H + j), cprev = __ldg(cst + (row - 1) * H + j);   // row - 1 = pad row (0) at t = 0
    const float tc = tanhf(c);
    float dh = __ldg(dH + row * H + j);
    if (!last)
      for (int q = 0; q < dh_parts; ++q) dh += __ldg(dh_rec + q * n + i);
    const float dcv = (last ? 0.f : dc[i]) + dh * og * (1.f - tc * tc);
    const float d_o = dh * tc * og * (1.f - og);
    const float d_i = dcv * gg * ig * (1.f - ig);
    const float d_g = dcv * ig * (1.f - gg * gg);
    const float d_f = dcv * cprev * fg * (1.f - fg);
    dc[i] = dcv * fg;
    float* dp = dP + row * 4 * H;
    dp[j] = d_i; dp[H + j] = d_f; dp[2 * H + j] = d_g; dp[3 * H + j] = d_o;
    const long long hl = 4LL * NB * 4 * H, so = sb * 4 * H;
    st_split1(dP_step, hl, so + j, d_i);
    st_split1(dP_step, hl, so + H + j, d_f);
    st_split1(dP_step, hl, so + 2 * H + j, d_g);
    st_split1(dP_step, hl, so + 3 * H + j, d_o);
  }
}

// dlatent (NB, Tv, H, 2) -> dH [4][R][H]: h_rr = +d_re, h_ir = +d_im, h_ri = +d_im, h_ii = -d_re (pad rows 0)
__global__ void __launch_bounds__(256) lstm_combine_bwd_kernel(const float* __restrict__ dl, int NB, int T, int Tv, int H,
                                                               float* __restrict__ dH) {
  const int Tp = T + 1;
  const long long R = (long long)NB * Tp, n = R * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % H);
    const long long r = i / H;
    const int tt = (int)(r % Tp), b = (int)(r / Tp);
    float2 d = make_float2(0.f, 0.f);
    if (tt != 0 && tt <= Tv) d = __ldg(reinterpret_cast<const float2*>(dl + (((long long)b * Tv + tt - 1) * H + j) * 2));
    dH[i] = d.x;
    dH[R * H + i] = d.y;
    dH[2 * R * H + i] = d.y;
    dH[3 * R * H + i] = -d.x;
  }
}

// out[col] += sum_r x[r][col]  (x: [rows][ld], cols <= ld).  grid (ceil(cols/256), chunks)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long rows, int cols, int ld,
                                                     int rows_per_chunk, float* __restrict__ out) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= cols) return;
  long long r0 = (long long)blockIdx.y * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > rows) r1 = rows;
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += __ldg(x + r * ld + col);
  atomicAdd(out + col, s);
}

// dW[tap][part][n] += sum over (fo, rows) x[b][fi][ti][part] * dY[fo][r][n];  grid (Fout, chunks), block = N threads
__global__ void enc0_wgrad_kernel(const float* __restrict__ stft, const float* __restrict__ dY, int NB, int Fin, int T,
                                  int N, int Fout, int causal, int rows_per_chunk, float* __restrict__ dW) {
  const int n = threadIdx.x, fo = blockIdx.x;
  const int Tp = T + 1;
  const long long R = (long long)NB * Tp;
  long long r0 = (long long)blockIdx.y * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > R) r1 = R;
  float acc[20];
#pragma unroll
  for (int k = 0; k < 20; ++k) acc[k] = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const int b = (int)(r / Tp), t = (int)(r % Tp) - 1;
    if (t < 0) continue;
    const float g = __ldg(dY + ((long long)fo * R + r) * N + n);
#pragma unroll
    for (int tap = 0; tap < 10; ++tap) {
      const int kf = tap >> 1, kt = tap & 1;
      const int fi = 2 * fo + kf - 2, ti = causal ? t - 1 + kt : t + kt;
      if (fi >= 0 && fi < Fin && ti >= 0 && ti < T) {
        const float2 x = __ldg(reinterpret_cast<const float2*>(stft + ((long long)(b * Fin + fi) * T + ti) * 2));
        acc[tap * 2] = fmaf(x.x, g, acc[tap * 2]);
        acc[tap * 2 + 1] = fmaf(x.y, g, acc[tap * 2 + 1]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 20; ++k) atomicAdd(dW + k * N + n, acc[k]);
}

__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        float lr, float b1, float b2, float eps, float wd, float bc1,
                                                        float bc2_sqrt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gv = g[i] + wd * pv;                              // torch.optim.Adam: L2 penalty added to the gradient
    const float mv = b1 * m[i] + (1.f - b1) * gv;
    const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
    m[i] = mv;
    v[i] = vv;
    p[i] = pv - (lr / bc1) * mv / (sqrtf(vv) / bc2_sqrt + eps);
  }
}


// ---- closed-form KL between two complex Gaussians, forward + gradient (model/nsvae_loss.py:L275-328) ------------------
// lat1 (n_bt, H1, 2): (mu, log sigma, delta) of distribution 1 at channel ch1 (zdim each); lat2 likewise (constant).
// acc[0] += scale * sum_bt kl, acc[1] += mean_scale * sum_bt kl; dlat1 += scale * d(sum kl)/dlat1.
__global__ void __launch_bounds__(256) kl_fwd_bwd_kernel(const float* __restrict__ lat1, int H1, int ch1,
                                                         const float* __restrict__ lat2, int H2, int ch2,
                                                         long long n_bt, int zdim, float scale, float mean_scale,
                                                         float* __restrict__ dlat1, double* __restrict__ acc) {
  const float e = 1e-10f;                                      // standard_nsvae_loss_true_kl.epsilon (L250)
  const long long n = n_bt * zdim;
  double part = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % zdim);
    const long long bt = i / zdim;
    const float* a = lat1 + (bt * H1 + ch1 + j) * 2;
    const float* b = lat2 + (bt * H2 + ch2 + j) * 2;
    const float m1r = a[0], m1i = a[1], ls1 = a[2 * zdim], d1r = a[4 * zdim], d1i = a[4 * zdim + 1];
    const float m2r = b[0], m2i = b[1], ls2 = b[2 * zdim];
    float p2r = b[4 * zdim], p2i = b[4 * zdim + 1];
    const float s1 = expf(ls1), s2 = expf(ls2);
    const float ad1 = sqrtf(d1r * d1r + d1i * d1i + e), tmp1 = s1 * 0.99f / (ad1 + e);
    const bool cl1 = ad1 >= s1 - 1e-3f;
    const float p1r = cl1 ? d1r * tmp1 : d1r, p1i = cl1 ? d1i * tmp1 : d1i;
    const float a1 = p1r * p1r + p1i * p1i;
    const float ad2 = sqrtf(p2r * p2r + p2i * p2i + e), tmp2 = s2 * 0.99f / (ad2 + e);
    if (ad2 >= s2 - 1e-3f) { p2r *= tmp2; p2i *= tmp2; }
    const float a2 = p2r * p2r + p2i * p2i;
    const float den1 = 0.25f * (s1 * s1 - a1) + e, den2 = 0.25f * (s2 * s2 - a2) + e;
    const float coeff = 2.f / (s2 * s2 - a2 + e);
    const float trace = s1 * s2 - p2r * p1r - p2i * p1i;
    const float dr = m2r - m1r, di = m2i - m1i;
    const float quad = dr * dr * (s2 - p2r) - 2.f * p2i * dr * di + di * di * (s2 + p2r);
    part += (double)(coeff * (trace + quad) + logf(den2) - logf(den1));
    if (dlat1) {
      const float k = 0.5f * scale;
      const float g_m1r = coeff * (-2.f * dr * (s2 - p2r) + 2.f * p2i * di);
      const float g_m1i = coeff * (2.f * p2i * dr - 2.f * di * (s2 + p2r));
      float g_s1 = coeff * s2 - 0.5f * s1 / den1;
      const float g_p1r = -coeff * p2r + 0.5f * p1r / den1, g_p1i = -coeff * p2i + 0.5f * p1i / den1;
      float g_d1r = g_p1r, g_d1i = g_p1i;
      if (cl1) {                                               // p1 = d1 * 0.99 s1 / (|d1| + e)
        const float dot = g_p1r * d1r + g_p1i * d1i;
        g_s1 += dot * 0.99f / (ad1 + e);
        const float dtmp = -0.99f * s1 / ((ad1 + e) * (ad1 + e)) / ad1;
        g_d1r = g_p1r * tmp1 + dot * dtmp * d1r;
        g_d1i = g_p1i * tmp1 + dot * dtmp * d1i;
      }
      float* d = dlat1 + (bt * H1 + ch1 + j) * 2;
      d[0] += k * g_m1r;
      d[1] += k * g_m1i;
      d[2 * zdim] += k * g_s1 * s1;                            // log sigma: only the real part is used (L285)
      d[4 * zdim] += k * g_d1r;
      d[4 * zdim + 1] += k * g_d1i;
    }
  }
  // block reduction of the partial sums
  __shared__ double red[8];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    t *= 0.5;
    if (blockIdx.x == 0) t -= (double)zdim * (double)n_bt;      // kl = 0.5 sum_j(...) - zdim per (b, t)
    atomicAdd(acc, t * (double)scale);
    atomicAdd(acc + 1, t * (double)mean_scale);
  }
}

static inline int grid_for(long long n, int per_sm) {
  const long long b = (n + 255) / 256;
  return (int)(b < 148LL * per_sm ? (b > 0 ? b : 1) : 148LL * per_sm);
}

}  // namespace idv

extern "C" int idv_planes_transpose_split(const void* planes, int in_split, int F, int R, int Cp, int Rpad, int shift,
                                          void* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(planes && out && F > 0 && F <= 65535 && R > 0 && Cp > 0 && Cp % 2 == 0 && Rpad >= R && Rpad % 64 == 0,
                "idv_planes_transpose_split: bad argument (Cp even, Rpad a multiple of 64 >= R)");
  dim3 grid(Rpad / 64, cdiv(Cp, 64), F), block(256);
  IDV_CHECK_ARG(grid.y <= 65535, "idv_planes_transpose_split: Cp too large");
  planes_transpose_split_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(planes, in_split, F, R, Cp, Rpad, shift,
                                                                          reinterpret_cast<unsigned short*>(out));
  IDV_LAUNCH_CHECK("planes_transpose_split_kernel");
  return IDV_OK;
}

extern "C" int idv_f32_to_split(const float* x, int64_t n, void* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(x && out && n > 0 && n % 4 == 0, "idv_f32_to_split: n must be a positive multiple of 4");
  f32_to_split_kernel<<<grid_for(n / 4, 16), 256, 0, (cudaStream_t)stream>>>(x, n / 4,
                                                                             reinterpret_cast<unsigned short*>(out));
  IDV_LAUNCH_CHECK("f32_to_split_kernel");
  return IDV_OK;
}

extern "C" int idv_cbn_bwd_reduce(const void* y, int y_split, const void* g, int g_split, int NB, int C, int F, int T,
                                  const float* stats, const float* zb, float slope, double* acc, int t_valid,
                                  void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(y && g && stats && zb && acc && NB > 0 && C > 0 && C <= 1024 && F > 0 && F <= 65535 && T > 0,
                "idv_cbn_bwd_reduce: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  IDV_CUDA(cudaMemsetAsync(acc, 0, (size_t)C * 8 * sizeof(double), st));
  const long long R = (long long)NB * (T + 1);
  IDV_CHECK_ARG(C <= 256, "idv_cbn_bwd_reduce: at most 256 complex channels");
  int chunks = (int)(148LL * 8 / F);
  if (chunks < 1) chunks = 1;
  if (chunks > R) chunks = (int)R;
  const int rows_per_chunk = (int)((R + chunks - 1) / chunks);
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  dim3 grid(F, chunks);
  const int Cw = ((C + 31) / 32) * 32;
  dim3 block(Cw, 256 / Cw);
  cbn_bwd_reduce_kernel<<<grid, block, 0, st>>>(y, y_split, g, g_split, NB, C, F, T, Tv, rows_per_chunk, stats, zb, slope,
                                                acc);
  IDV_LAUNCH_CHECK("cbn_bwd_reduce_kernel");
  return IDV_OK;
}

extern "C" int idv_cbn_bwd_finalize(const double* acc, double count, int C, const float* stats, const float* gamma_rr,
                                    const float* gamma_ri, const float* gamma_ii, float* coef, float* d_gamma_rr,
                                    float* d_gamma_ri, float* d_gamma_ii, float* d_beta_r, float* d_beta_i,
                                    double* d_slope, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(acc && stats && gamma_rr && gamma_ri && gamma_ii && coef && d_gamma_rr && d_gamma_ri && d_gamma_ii &&
                    d_beta_r && d_beta_i && C > 0 && count > 0,
                "idv_cbn_bwd_finalize: bad argument");
  cbn_bwd_finalize_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(acc, count, C, stats, gamma_rr, gamma_ri,
                                                                          gamma_ii, coef, d_gamma_rr, d_gamma_ri,
                                                                          d_gamma_ii, d_beta_r, d_beta_i, d_slope);
  IDV_LAUNCH_CHECK("cbn_bwd_finalize_kernel");
  return IDV_OK;
}

extern "C" int idv_cbn_bwd_apply(const void* y, int y_split, const void* g, int g_split, int NB, int C, int F, int T,
                                 const float* stats, const float* zb, const float* coef, float slope, void* dy,
                                 int dy_split, int t_valid, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(y && g && stats && zb && coef && dy && NB > 0 && C > 0 && F > 0 && T > 0, "idv_cbn_bwd_apply: bad argument");
  const long long n = (long long)F * NB * (T + 1) * ((C + 7) / 8 * 8);
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  cbn_bwd_apply_kernel<<<grid_for(n, 32), 256, 0, (cudaStream_t)stream>>>(y, y_split, g, g_split, NB, C, F, T, Tv, stats,
                                                                          zb, coef, slope, dy, dy_split);
  IDV_LAUNCH_CHECK("cbn_bwd_apply_kernel");
  return IDV_OK;
}

extern "C" int idv_lstm_scan_c(const float* P, int NB, int T, int H, float* cst, int t_valid, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(P && cst && NB > 0 && T > 0 && H > 0, "idv_lstm_scan_c: bad argument");
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  lstm_scan_c_kernel<<<grid_for(4LL * NB * H, 16), 256, 0, (cudaStream_t)stream>>>(P, NB, T, Tv, H, cst);
  IDV_LAUNCH_CHECK("lstm_scan_c_kernel");
  return IDV_OK;
}

extern "C" int idv_lstm_cell_bwd_step(const float* P, const float* cst, const float* dH, const float* dh_rec, float* dc,
                                      int NB, int T, int H, int t, int last, int dh_parts, float* dP, void* dP_step,
                                      void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(P && cst && dH && dc && dP && dP_step && (last || (dh_rec && dh_parts > 0)) && NB > 0 && T > 0 && H > 0 &&
                    t >= 0 && t < T,
                "idv_lstm_cell_bwd_step: bad argument");
  lstm_cell_bwd_step_kernel<<<grid_for(4LL * NB * H, 8), 256, 0, (cudaStream_t)stream>>>(
      P, cst, dH, dh_rec, dc, NB, T, H, t, last, dh_parts, dP, reinterpret_cast<unsigned short*>(dP_step));
  IDV_LAUNCH_CHECK("lstm_cell_bwd_step_kernel");
  return IDV_OK;
}

extern "C" int idv_lstm_combine_bwd(const float* dlatent, int NB, int T, int H, float* dH, int t_valid, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(dlatent && dH && NB > 0 && T > 0 && H > 0, "idv_lstm_combine_bwd: bad argument");
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  lstm_combine_bwd_kernel<<<grid_for((long long)NB * (T + 1) * H, 16), 256, 0, (cudaStream_t)stream>>>(dlatent, NB, T, Tv,
                                                                                                    H, dH);
  IDV_LAUNCH_CHECK("lstm_combine_bwd_kernel");
  return IDV_OK;
}

extern "C" int idv_colsum_add(const float* x, int64_t rows, int cols, int ld, float* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(x && out && rows > 0 && cols > 0 && ld >= cols, "idv_colsum_add: bad argument");
  int chunks = 148 * 4 / cdiv(cols, 256);
  if (chunks < 1) chunks = 1;
  if (chunks > rows) chunks = (int)rows;
  const int rpc = (int)((rows + chunks - 1) / chunks);
  dim3 grid(cdiv(cols, 256), chunks);
  colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, rpc, out);
  IDV_LAUNCH_CHECK("colsum_kernel");
  return IDV_OK;
}

extern "C" int idv_enc0_wgrad(const float* stft, const float* dY, int B, int Fin, int T, int Cout, int causal, float* dW,
                              void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(stft && dY && dW && B > 0 && Fin >= 5 && T > 0 && Cout > 0 && 2 * Cout <= 1024, "idv_enc0_wgrad: bad argument");
  const int Fout = (Fin + 4 - 5) / 2 + 1, N = 2 * Cout;
  const long long R = (long long)B * (T + 1);
  int chunks = 148 * 8 / Fout;
  if (chunks < 1) chunks = 1;
  if (chunks > R) chunks = (int)R;
  const int rpc = (int)((R + chunks - 1) / chunks);
  cudaStream_t st = (cudaStream_t)stream;
  IDV_CUDA(cudaMemsetAsync(dW, 0, (size_t)20 * N * sizeof(float), st));
  dim3 grid(Fout, chunks);
  enc0_wgrad_kernel<<<grid, N, 0, st>>>(stft, dY, B, Fin, T, N, Fout, causal, rpc, dW);
  IDV_LAUNCH_CHECK("enc0_wgrad_kernel");
  return IDV_OK;
}

extern "C" int idv_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                             float eps, float weight_decay, int step, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "idv_adam_step: bad argument");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = sqrtf(1.f - powf(beta2, (float)step));
  adam_step_kernel<<<grid_for(n, 16), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                      bc1, bc2);
  IDV_LAUNCH_CHECK("adam_step_kernel");
  return IDV_OK;
}

extern "C" int idv_kl_fwd_bwd(const float* lat1, int H1, int ch1, const float* lat2, int H2, int ch2, int64_t n_bt,
                              int zdim, float scale, float mean_scale, float* dlat1, double* acc, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(lat1 && lat2 && acc && n_bt > 0 && zdim > 0 && ch1 >= 0 && ch1 + 3 * zdim <= H1 && ch2 >= 0 &&
                    ch2 + 3 * zdim <= H2,
                "idv_kl_fwd_bwd: bad argument");
  kl_fwd_bwd_kernel<<<grid_for(n_bt * zdim, 8), 256, 0, (cudaStream_t)stream>>>(lat1, H1, ch1, lat2, H2, ch2,
                                                                                (long long)n_bt, zdim, scale, mean_scale,
                                                                                dlat1, acc);
  IDV_LAUNCH_CHECK("kl_fwd_bwd_kernel");
  return IDV_OK;
}

// ====================================================================================================================
// decoder side (phase-2 training step, i_dccrn_vae/nsvae_dccrn/train_second_phase_decoder.py:L376-433: frozen encoder,
// decoder train=True, SI-SNR loss): SI-SNR value + gradient, adjoint of the overlap-add, backward of the reconstruction
// head.  The transposed convs / ComplexBatchNormal / PReLU / dense reuse the kernels above and the tap-GEMM.
// ====================================================================================================================
namespace idv {

// sums[b][3] = <est, src>, |src|^2, |est|^2   (grid (chunks, B))
__global__ void __launch_bounds__(256) sisnr_reduce_kernel(const float* __restrict__ src, const float* __restrict__ est,
                                                           int L, double* __restrict__ sums) {
  const int b = blockIdx.y;
  double d = 0, ss = 0, ee = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
    const float s = __ldg(src + (long long)b * L + i), e = __ldg(est + (long long)b * L + i);
    d += (double)s * e; ss += (double)s * s; ee += (double)e * e;
  }
  __shared__ double red[3][8];
  for (int o = 16; o > 0; o >>= 1) {
    d += __shfl_down_sync(0xffffffffu, d, o);
    ss += __shfl_down_sync(0xffffffffu, ss, o);
    ee += __shfl_down_sync(0xffffffffu, ee, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = d; red[1][threadIdx.x >> 5] = ss; red[2][threadIdx.x >> 5] = ee; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(sums + b * 3 + threadIdx.x, t);
  }
}

// model/nsvae_loss.py:L877-889: alpha = <est,src>/(|src|^2+eps), s_t = alpha src, e = est - s_t,
// snr = 10 log10(|s_t|^2/(|e|^2+eps) + eps), loss = -mean_b snr.  d_est += scale * d(loss)/d(est); loss[0] += loss.
__global__ void __launch_bounds__(256) sisnr_grad_kernel(const float* __restrict__ src, const float* __restrict__ est,
                                                         int B, int L, const double* __restrict__ sums, float scale,
                                                         float* __restrict__ d_est, double* __restrict__ loss) {
  const int b = blockIdx.y;
  const double eps = 1e-8;
  const double dot = sums[b * 3], ss = sums[b * 3 + 1], ee = sums[b * 3 + 2];
  const double alpha = dot / (ss + eps);
  const double a = alpha * alpha * ss;                              // |s_target|^2
  const double n = ee - 2 * alpha * dot + a;                        // |e_noise|^2
  const double ratio = a / (n + eps) + eps;
  // d snr / d est = k * [ (da/dest)/(n+eps) - a/(n+eps)^2 dn/dest ],  k = 10 / (ln10 * ratio)
  const double k = 10.0 / (2.302585092994046 * ratio);
  const double da_c = 2 * alpha * ss / (ss + eps);                  // da/dest = da_c * src
  const double es = dot - alpha * ss;                               // <e, src>
  const double c_src = k * (da_c / (n + eps) + a / ((n + eps) * (n + eps)) * 2 * (alpha + es / (ss + eps)));
  const double c_est = -k * a / ((n + eps) * (n + eps)) * 2;        // dn/dest = 2 (est - alpha src) - 2 es/(ss+eps) src
  const double w = -(double)scale / B;
  if (d_est)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L; i += gridDim.x * blockDim.x) {
      const long long idx = (long long)b * L + i;
      d_est[idx] += (float)(w * (c_src * __ldg(src + idx) + c_est * __ldg(est + idx)));
    }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(loss, -(10.0 * log10(ratio)) / B);
}

// two_phase_loss.multi_recon_loss (model/nsvae_loss.py:L891-913) without its SI-SNR term: per bin of pred / ori (n bins,
// (re, im) interleaved)  cpx = (pr-or)^2 + (pi-oi)^2,  mag = (sqrt(pr^2+pi^2+1e-6) - sqrt(or^2+or^2+1e-6))^2 - the
// reference's ori magnitude squares the REAL part twice (L899) and that is kept.  acc[0] += inv_bt * sum cpx,
// acc[1] += inv_bt * sum mag,  d_pred += inv_bt * (w_cpx d cpx + w_mag d mag).  HBM-bound: 16 B read (+ 8 B
// read-modify-write of d_pred) per bin.
__global__ void __launch_bounds__(256) spec_loss_kernel(const float2* __restrict__ pred, const float2* __restrict__ ori,
                                                        long long n, float w_cpx, float w_mag, float inv_bt,
                                                        float2* __restrict__ d_pred, double* __restrict__ acc) {
  double s_c = 0, s_m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float2 p = __ldg(pred + i), o = __ldg(ori + i);
    const float dr = p.x - o.x, di = p.y - o.y;
    const float pm = sqrtf(p.x * p.x + p.y * p.y + 1e-6f), om = sqrtf(o.x * o.x + o.x * o.x + 1e-6f);
    const float dm = pm - om;
    s_c += (double)(dr * dr + di * di);
    s_m += (double)(dm * dm);
    if (d_pred) {
      const float k = 2.f * w_mag * dm / pm;
      float2 g = d_pred[i];
      g.x += inv_bt * (2.f * w_cpx * dr + k * p.x);
      g.y += inv_bt * (2.f * w_cpx * di + k * p.y);
      d_pred[i] = g;
    }
  }
  __shared__ double red[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    s_c += __shfl_down_sync(0xffffffffu, s_c, o);
    s_m += __shfl_down_sync(0xffffffffu, s_m, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_c; red[1][threadIdx.x >> 5] = s_m; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(acc + threadIdx.x, t * (double)inv_bt);
  }
}

// adjoint of idv_ola_fwd: dframes[(b,t)][j] = dsig[b][hop*t - (n_fft/2 - off) + j] / env  for j < win (0 elsewhere / beyond)
__global__ void __launch_bounds__(256) ola_bwd_kernel(const float* __restrict__ dsig, const float* __restrict__ wsq, int T,
                                                      int n_fft, int hop, int win, int out_len, int frame_ld,
                                                      float* __restrict__ dframes, long long n) {
  const int off = (n_fft - win) / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % frame_ld);
    const long long bt = i / frame_ld;
    const int t = (int)(bt % T), b = (int)(bt / T);
    float v = 0.f;
    const int s = hop * t + off + j - n_fft / 2;                    // output sample this tap contributes to
    if (j < win && s >= 0 && s < out_len) {
      const int q = s + n_fft / 2 - off;
      int t_hi = q / hop, t_lo = (q - win + hop) / hop;
      if (q - win + 1 <= 0) t_lo = 0;
      if (t_hi > T - 1) t_hi = T - 1;
      float env = 0.f;
      for (int tt = t_lo; tt <= t_hi; ++tt) {
        const int jj = q - hop * tt;
        if (jj >= 0 && jj < win) env += __ldg(wsq + jj);
      }
      v = __ldg(dsig + (long long)b * out_len + s) / env;
    }
    dframes[i] = v;
  }
}

// Backward of the reconstruction head on the last decoder layer (Cout = 1).  raw (NB, F, T, 2): transposed-conv
// output before ComplexBatchNormal; zb[6]: Z, b' of the batch statistics; pre = Z raw + b', m = PReLU(pre).
//   real_imag head: S = m;   mask head (model/pvae_module.py:L2594-2609): S = X * tanh(|m|)/|m| * m.
// drows[(b*T + t)][2k + part] (+ dpred (NB, F, T, 2) if given) = dL/dS.  Writes planes (C = 1, Cp = 16) y <- raw and g <- dL/dm for the cbn_bwd kernels.
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ zb,
                                                       float slope, int mask, const float* __restrict__ stft_x,
                                                       const float* __restrict__ drows, int drows_ld,
                                                       const float* __restrict__ dpred, int NB, int F, int T,
                                                       float* __restrict__ y_planes, float* __restrict__ g_planes) {
  const long long n = (long long)NB * F * T;
  const int Tp = T + 1;
  const long long R = (long long)NB * Tp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const int f = (int)((i / T) % F);
    const int b = (int)(i / ((long long)T * F));
    const float2 y = __ldg(reinterpret_cast<const float2*>(raw + i * 2));
    const float2 gs = __ldg(reinterpret_cast<const float2*>(drows + ((long long)b * T + t) * drows_ld + 2 * f));
    float gr = gs.x, gi = gs.y;
    if (dpred) {                                                      // gradient that reached `predict` directly
      const float2 gp = __ldg(reinterpret_cast<const float2*>(dpred + i * 2));
      gr += gp.x; gi += gp.y;
    }
    if (mask) {
      const float pr = fmaf(zb[0], y.x, fmaf(zb[1], y.y, zb[4])), pi = fmaf(zb[2], y.x, fmaf(zb[3], y.y, zb[5]));
      const float mr = prelu_f(pr, slope), mi = prelu_f(pi, slope);
      const float2 X = __ldg(reinterpret_cast<const float2*>(stft_x + i * 2));
      // S = X u (complex), u = h(r) m, h = tanh(r)/r:  g_u = conj(X) g,  g_m = h g_u + h'(r)/r (g_u . m) m
      const float gur = X.x * gr + X.y * gi, gui = X.x * gi - X.y * gr;
      const float r = sqrtf(mr * mr + mi * mi);
      float h = 1.f, hp_r = 0.f;
      if (r > 1e-12f) {
        const float th = tanhf(r);
        h = th / r;
        hp_r = ((1.f - th * th) * r - th) / (r * r * r);              // h'(r) / r
      }
      const float dotm = gur * mr + gui * mi;
      gr = h * gur + hp_r * dotm * mr;
      gi = h * gui + hp_r * dotm * mi;
    }
    const long long row = ((long long)f * R + (long long)b * Tp + 1 + t) * 16;
    y_planes[row] = y.x; y_planes[row + 8] = y.y;
    g_planes[row] = gr; g_planes[row + 8] = gi;
  }
}


// ---- last decoder layer (Cout = 1, kernel (5,2), stride (2,1), freq pad 2, causal) --------------------------------------
// dy planes fp32 [Fout = 2 Fin - 1][R][16] (re at channel 0, im at 8), w10 [10 (kf*2+kt)][Ktot][2] (pack_dec5 without fold).
// dx[fi][r][c] = sum_{kf,kt} dy[2fi - 2 + kf][r + kt] . w10[kf*2+kt][k_off + c]      grid (row tiles of 64, Fin)
__global__ void __launch_bounds__(256) dec5_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w10,
                                                         int Ktot, int k_off, int Cp, int Fin, int NB, int T,
                                                         float* __restrict__ dx) {
  extern __shared__ float ws[];                                       // [10][Cp][2]
  for (int i = threadIdx.x; i < 10 * Cp * 2; i += 256) {
    const int tap = i / (Cp * 2), rem = i % (Cp * 2);
    ws[i] = __ldg(w10 + ((long long)tap * Ktot + k_off) * 2 + rem);
  }
  __syncthreads();
  const int Tp = T + 1, Fout = 2 * Fin - 1, fi = blockIdx.y;
  const long long R = (long long)NB * Tp;
  const int q = Cp >> 2, lanes = 256 / q;
  if ((int)threadIdx.x >= lanes * q) return;
  const int c4 = (threadIdx.x % q) * 4;
  long long r_end = (long long)(blockIdx.x + 1) * 64;
  if (r_end > R) r_end = R;
  for (long long r = (long long)blockIdx.x * 64 + threadIdx.x / q; r < r_end; r += lanes) {
    const int tt = (int)(r % Tp);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt != 0) {
#pragma unroll
      for (int kf = 0; kf < 5; ++kf) {
        const int fo = 2 * fi - 2 + kf;
        if (fo < 0 || fo >= Fout) continue;
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          if (kt && tt == T) continue;                                // the frame after the last one is dropped
          const float* g = dy + ((long long)fo * R + r + kt) * 16;
          const float gr = __ldg(g), gi = __ldg(g + 8);
          const float* w = ws + ((kf * 2 + kt) * Cp + c4) * 2;
          a.x = fmaf(w[0], gr, fmaf(w[1], gi, a.x));
          a.y = fmaf(w[2], gr, fmaf(w[3], gi, a.y));
          a.z = fmaf(w[4], gr, fmaf(w[5], gi, a.z));
          a.w = fmaf(w[6], gr, fmaf(w[7], gi, a.w));
        }
      }
    }
    *reinterpret_cast<float4*>(dx + ((long long)fi * R + r) * Cp + c4) = a;
  }
}

// dW[tap][k_off + c][part] += sum_{fi, r} x[fi][r][c] * dy[2fi - 2 + kf][r + kt][part]     grid (Fin, chunks), block 256
// thread = (4 channels, row lane): the 20 gradient values of a row are loaded once per 80 FMAs.
__global__ void __launch_bounds__(256) dec5_wgrad_kernel(const void* __restrict__ x, int x_split,
                                                         const float* __restrict__ dy, int Ktot, int k_off, int Cp,
                                                         int Fin, int NB, int T, int rows_per_chunk,
                                                         float* __restrict__ dW) {
  extern __shared__ float red[];                                     // [nj][Cp][20]
  const int Tp = T + 1, Fout = 2 * Fin - 1, fi = blockIdx.x;
  const long long R = (long long)NB * Tp, hl = (long long)Fin * R * Cp;
  const int q = Cp >> 2, nj = 256 / q, c4 = (threadIdx.x % q) * 4, j = threadIdx.x / q;
  long long r0 = (long long)blockIdx.y * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > R) r1 = R;
  float acc[4][20];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < 20; ++k) acc[a][k] = 0.f;
  if (j < nj)
    for (long long r = r0 + j; r < r1; r += nj) {
      const int tt = (int)(r % Tp);
      if (tt == 0) continue;
      const long long idx = ((long long)fi * R + r) * Cp + c4;
      const float4 xv = x_split ? ld_split4(reinterpret_cast<const unsigned short*>(x), hl, idx)
                                : ldg4(reinterpret_cast<const float*>(x) + idx);
#pragma unroll
      for (int kf = 0; kf < 5; ++kf) {
        const int fo = 2 * fi - 2 + kf;
        if (fo < 0 || fo >= Fout) continue;
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          if (kt && tt == T) continue;
          const float* gp = dy + ((long long)fo * R + r + kt) * 16;
          const float gr = __ldg(gp), gi = __ldg(gp + 8);
          const int k = (kf * 2 + kt) * 2;
          acc[0][k] = fmaf(xv.x, gr, acc[0][k]); acc[0][k + 1] = fmaf(xv.x, gi, acc[0][k + 1]);
          acc[1][k] = fmaf(xv.y, gr, acc[1][k]); acc[1][k + 1] = fmaf(xv.y, gi, acc[1][k + 1]);
          acc[2][k] = fmaf(xv.z, gr, acc[2][k]); acc[2][k + 1] = fmaf(xv.z, gi, acc[2][k + 1]);
          acc[3][k] = fmaf(xv.w, gr, acc[3][k]); acc[3][k + 1] = fmaf(xv.w, gi, acc[3][k + 1]);
        }
      }
    }
  if (j < nj)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int k = 0; k < 20; ++k) red[((long long)j * Cp + c4 + a) * 20 + k] = acc[a][k];
  __syncthreads();
  for (int i = threadIdx.x; i < Cp * 20; i += 256) {
    float t = 0.f;
    for (int jj = 0; jj < nj; ++jj) t += red[(long long)jj * Cp * 20 + i];
    const int c = i / 20, k = i % 20;
    atomicAdd(dW + ((long long)(k >> 1) * Ktot + k_off + c) * 2 + (k & 1), t);
  }
}

__global__ void __launch_bounds__(256) axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = *reinterpret_cast<float4*>(y + i * 4);
    const float4 u = ldg4(x + i * 4);
    v.x = fmaf(a, u.x, v.x); v.y = fmaf(a, u.y, v.y); v.z = fmaf(a, u.z, v.z); v.w = fmaf(a, u.w, v.w);
    *reinterpret_cast<float4*>(y + i * 4) = v;
  }
}

// ---- reparameterisation backward (model/pvae_module.py:L2177-2231, S = 1) ----------------------------------------------
// dlatent[(bt)][ch0 + {0, zdim, 2 zdim} + j][2] += d z / d (mu, log sigma, delta) applied to dz (NB, T, zdim, 2).
// The imaginary part of log sigma is ignored by the forward (its gradient is 0).
__global__ void __launch_bounds__(256) reparam_bwd_kernel(const float* __restrict__ latent, long long n_bt, int Htot,
                                                          int ch0, int zdim, const float* __restrict__ eps_r,
                                                          const float* __restrict__ eps_i,
                                                          const float* __restrict__ dz, float* __restrict__ dlat) {
  const float e = 1e-6f;
  const long long n = n_bt * zdim;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % zdim);
    const long long bt = i / zdim;
    const long long base = (bt * Htot + ch0 + j) * 2;
    const float ls = latent[base + 2 * zdim];
    const float d0r = latent[base + 4 * zdim], d0i = latent[base + 4 * zdim + 1];
    const float er = eps_r[i], ei = eps_i[i];
    const float gzr = dz[i * 2], gzi = dz[i * 2 + 1];
    // forward (fp32 like the kernel), keeping what the chain rule needs
    const float sig = expf(ls);
    const float ad0 = sqrtf(d0r * d0r + d0i * d0i + e);
    const bool clamp = ad0 >= sig - 1e-3f;
    const float tmp = sig * 0.99f / (ad0 + e);
    const float dr = clamp ? d0r * tmp : d0r, di = clamp ? d0i * tmp : d0i;
    const float ad2 = dr * dr + di * di + e;                          // |delta|^2 + eps after the protection
    const float den = sqrtf(2.f * (sig + dr) + e);
    const float de = den + e;
    const float num = sig + dr;
    const float rad = sig * sig - ad2 + e;
    const float sq = sqrtf(rad);
    // z_r = mu_r + num/de * er;  z_i = mu_i + di/de * er + sq/de * ei
    const float g_num = gzr * er / de;
    const float g_di_direct = gzi * er / de;
    const float g_sq = gzi * ei / de;
    const float g_de = -(gzr * er * num + gzi * (er * di + ei * sq)) / (de * de);
    const float g_rad = g_sq * 0.5f / sq;
    const float g_den = g_de;                                         // de = den + e
    const float g_two = g_den * 0.5f / den;                           // den = sqrt(2 (sig + dr) + e)
    float g_sig = g_num + 2.f * g_two + g_rad * 2.f * sig;
    float g_dr = g_num + 2.f * g_two - g_rad * 2.f * dr;
    float g_di = g_di_direct - g_rad * 2.f * di;
    float g_d0r, g_d0i;
    if (clamp) {
      // dr = d0r * tmp, di = d0i * tmp, tmp = 0.99 sig / (ad0 + e)
      const float g_tmp = g_dr * d0r + g_di * d0i;
      g_sig += g_tmp * 0.99f / (ad0 + e);
      const float g_ad0 = -g_tmp * tmp / (ad0 + e);
      g_d0r = g_dr * tmp + g_ad0 * d0r / ad0;
      g_d0i = g_di * tmp + g_ad0 * d0i / ad0;
    } else {
      g_d0r = g_dr;
      g_d0i = g_di;
    }
    dlat[base] += gzr;
    dlat[base + 1] += gzi;
    dlat[base + 2 * zdim] += g_sig * sig;                             // sig = exp(log sigma_re)
    dlat[base + 4 * zdim] += g_d0r;
    dlat[base + 4 * zdim + 1] += g_d0i;
  }
}

}  // namespace idv

extern "C" int idv_sisnr_fwd_bwd(const float* src, const float* est, int B, int L, float scale, float* d_est,
                                 double* sums, double* loss, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(src && est && sums && loss && B > 0 && B <= 65535 && L > 0, "idv_sisnr_fwd_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  IDV_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * 3 * sizeof(double), st));
  int chunks = cdiv(L, 256 * 16);
  if (chunks > 64) chunks = 64;
  dim3 grid(chunks, B);
  sisnr_reduce_kernel<<<grid, 256, 0, st>>>(src, est, L, sums);
  IDV_LAUNCH_CHECK("sisnr_reduce_kernel");
  sisnr_grad_kernel<<<grid, 256, 0, st>>>(src, est, B, L, sums, scale, d_est, loss);
  IDV_LAUNCH_CHECK("sisnr_grad_kernel");
  return IDV_OK;
}

extern "C" int idv_spec_loss_fwd_bwd(const float* pred, const float* ori, int64_t n_bins, float w_cpx, float w_mag,
                                     float inv_bt, float* d_pred, double* acc, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(pred && ori && acc && n_bins > 0, "idv_spec_loss_fwd_bwd: bad argument");
  long long blocks = (n_bins + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  spec_loss_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(pred), reinterpret_cast<const float2*>(ori), n_bins, w_cpx, w_mag, inv_bt,
      reinterpret_cast<float2*>(d_pred), acc);
  IDV_LAUNCH_CHECK("spec_loss_kernel");
  return IDV_OK;
}

extern "C" int idv_ola_bwd(const float* dsig, const float* wsq, int B, int T, int n_fft, int hop, int win, int frame_ld,
                           float* dframes, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(dsig && wsq && dframes && B > 0 && T > 1 && frame_ld >= win, "idv_ola_bwd: bad argument");
  const long long n = (long long)B * T * frame_ld;
  ola_bwd_kernel<<<grid_for(n, 16), 256, 0, (cudaStream_t)stream>>>(dsig, wsq, T, n_fft, hop, win, hop * (T - 1), frame_ld,
                                                                    dframes, n);
  IDV_LAUNCH_CHECK("ola_bwd_kernel");
  return IDV_OK;
}

extern "C" int idv_head_bwd(const float* raw, const float* zb, float slope, int mask, const float* stft_x,
                            const float* drows, int drows_ld, const float* dpred, int NB, int F, int T, float* y_planes,
                            float* g_planes, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(raw && zb && drows && y_planes && g_planes && (!mask || stft_x) && NB > 0 && F > 0 && T > 0 &&
                    drows_ld >= 2 * F,
                "idv_head_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)F * NB * (T + 1) * 16 * sizeof(float);
  IDV_CUDA(cudaMemsetAsync(y_planes, 0, bytes, st));
  IDV_CUDA(cudaMemsetAsync(g_planes, 0, bytes, st));
  head_bwd_kernel<<<grid_for((long long)NB * F * T, 16), 256, 0, st>>>(raw, zb, slope, mask, stft_x, drows, drows_ld, dpred,
                                                                       NB, F, T, y_planes, g_planes);
  IDV_LAUNCH_CHECK("head_bwd_kernel");
  return IDV_OK;
}

extern "C" int idv_dec5_dgrad(const float* dy, const float* w10, int Ktot, int k_off, int Cp, int Fin, int NB, int T,
                              float* dx, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(dy && w10 && dx && Cp >= 4 && Cp % 4 == 0 && Cp <= 1024 && k_off >= 0 && k_off + Cp <= Ktot && Fin > 0 &&
                    NB > 0 && T > 0,
                "idv_dec5_dgrad: bad argument");
  const long long R = (long long)NB * (T + 1);
  dim3 grid((unsigned)cdiv64(R, 64), Fin);
  dec5_dgrad_kernel<<<grid, 256, 10 * Cp * 2 * sizeof(float), (cudaStream_t)stream>>>(dy, w10, Ktot, k_off, Cp, Fin, NB, T, dx);
  IDV_LAUNCH_CHECK("dec5_dgrad_kernel");
  return IDV_OK;
}

extern "C" int idv_dec5_wgrad(const void* x, int x_split, const float* dy, int Ktot, int k_off, int Cp, int Fin, int NB,
                              int T, float* dW, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(x && dy && dW && Cp >= 8 && Cp <= 512 && Cp % 8 == 0 && k_off >= 0 && k_off + Cp <= Ktot && Fin > 0 &&
                    NB > 0 && T > 0,
                "idv_dec5_wgrad: bad argument (Cp a multiple of 8, <= 512)");
  const long long R = (long long)NB * (T + 1);
  int chunks = cdiv(148 * 8, Fin);
  const long long per = cdiv64(R, chunks);
  chunks = (int)cdiv64(R, per);
  dim3 grid(Fin, chunks);
  const int nj = 256 / (Cp / 4);
  const size_t smem = (size_t)nj * Cp * 20 * sizeof(float);
  if (smem > 48 * 1024)
    IDV_CUDA(cudaFuncSetAttribute(dec5_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dec5_wgrad_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, x_split, dy, Ktot, k_off, Cp, Fin, NB, T, (int)per, dW);
  IDV_LAUNCH_CHECK("dec5_wgrad_kernel");
  return IDV_OK;
}

extern "C" int idv_reparam_bwd(const float* latent, int NB, int T, int Htot, int ch0, int zdim, const float* eps_r,
                               const float* eps_i, const float* dz, float* dlatent, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(latent && eps_r && eps_i && dz && dlatent && NB > 0 && T > 0 && zdim > 0 && ch0 >= 0 &&
                    ch0 + 3 * zdim <= Htot,
                "idv_reparam_bwd: bad argument");
  const long long n = (long long)NB * T * zdim;
  reparam_bwd_kernel<<<grid_for(n, 4), 256, 0, (cudaStream_t)stream>>>(latent, (long long)NB * T, Htot, ch0, zdim, eps_r,
                                                                        eps_i, dz, dlatent);
  IDV_LAUNCH_CHECK("reparam_bwd_kernel");
  return IDV_OK;
}

extern "C" int idv_axpy(float* y, const float* x, float a, int64_t n, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(y && x && n > 0 && n % 4 == 0, "idv_axpy: n must be a positive multiple of 4");
  axpy_kernel<<<grid_for(n / 4, 16), 256, 0, (cudaStream_t)stream>>>(y, x, a, n / 4);
  IDV_LAUNCH_CHECK("axpy_kernel");
  return IDV_OK;
}
