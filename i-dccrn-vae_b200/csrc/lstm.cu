// Complex LSTM (model/complex_progress.py:L39-74): the sequential part.
//
// ComplexLSTM = two nn.LSTM modules (lstm_re, lstm_im) each run on x_re and x_im -> four independent
// streams (m, p).  The input projections for all T are tap-GEMMs (tapgemm_*.cu); this file holds
//   * the persistent recurrent kernel: one cooperative launch per layer, W_hh slices resident in
//     shared memory for the whole sequence, cell state resident in shared memory, h exchanged
//     through L2 with one grid barrier per time step;
//   * the stream combine (re = rr - ii, im = ir + ri) into the reference's (B,T,H,2) layout;
//   * the reparameterisation (model/pvae_module.py:L2177-2231) with supplied or Philox eps.
#include <curand_kernel.h>

#include "idv_common.cuh"

namespace idv {

constexpr int LS_KC = 384;        // K chunk staged in shared memory
constexpr int LS_ROWS = 32;       // rows per chunk (one per lane)

struct LstmRecParams {
  const float* g;
  int64_t g_m_off, g_p_off;
  int g_ld;
  const float* whh;               // [2][4H][H]
  int NB, T, H, Tsteps;           // T = allocated frames per utterance (row layout), Tsteps = valid steps
  float* hseq;                    // [4][R][H]
  unsigned short* hsplit;         // optional bf16 [2][4][R][H] copy (hi, lo) for the next layer's tensor-core in-proj
  unsigned int* sync;
  int NU, NR, Hs, RC;             // unit slices, row slices, units per CTA, rows per CTA
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire_u32(ctr) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// grid (NU, NR, 2); block = 32 * Hs/UT threads.  Lane = row, warp = UT hidden units.
template <int UT>
__global__ void lstm_rec_kernel(const LstmRecParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, Hs = p.Hs;
  float* Ws = smem;                                   // [Hs*4][H]   row = jl*4 + gate
  float* hs = Ws + (size_t)Hs * 4 * H;                // [32][LS_KC + 4]
  float* cs = hs + LS_ROWS * (LS_KC + 4);             // [RCpad][Hs]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthr = blockDim.x;
  const int m = blockIdx.z;
  const int u0 = blockIdx.x * Hs;
  const int nrows = 2 * p.NB;
  const int row0 = blockIdx.y * p.RC;
  const int row1 = min(nrows, row0 + p.RC);
  const int nchunks = (row1 - row0 + LS_ROWS - 1) / LS_ROWS;
  const int Tp = p.T + 1;
  const int64_t R = (int64_t)p.NB * Tp;
  const unsigned int nblk = gridDim.x * gridDim.y * gridDim.z;

  // resident W_hh slice
  const float* wm = p.whh + (size_t)m * 4 * H * H;
  for (int i = tid; i < Hs * 4 * (H / 4); i += nthr) {
    const int row = i / (H / 4), k4 = i % (H / 4);
    const int jl = row >> 2, gate = row & 3;
    *reinterpret_cast<float4*>(Ws + (size_t)row * H + k4 * 4) =
        ldg4(wm + ((size_t)gate * H + u0 + jl) * H + k4 * 4);
  }
  for (int i = tid; i < nchunks * LS_ROWS * Hs; i += nthr) cs[i] = 0.f;
  // zero initial state = the pad rows of this CTA's (rows, units)
  for (int i = tid; i < (row1 - row0) * Hs; i += nthr) {
    const int q = row0 + i / Hs, jl = i % Hs;
    const int pp = q / p.NB, b = q % p.NB;
    const int64_t zi = ((int64_t)(m * 2 + pp) * R + (int64_t)b * Tp) * H + u0 + jl;
    p.hseq[zi] = 0.f;
    if (p.hsplit) st_split1(p.hsplit, 4 * R * H, zi, 0.f);
  }
  unsigned int bar = 0;
  grid_barrier(p.sync, ++bar * nblk);

  const int jbase = warp * UT;                        // first local unit of this warp
  for (int t = 0; t < p.Tsteps; ++t) {
    for (int ch = 0; ch < nchunks; ++ch) {
      const int q = row0 + ch * LS_ROWS + lane;
      const bool valid = q < row1;
      const int pp = valid ? q / p.NB : 0, b = valid ? q % p.NB : 0;
      const int64_t rcur = (int64_t)b * Tp + 1 + t;
      // gate pre-activations from the input projection
      float acc[UT][4];
      {
        const float* gp = p.g + m * p.g_m_off + pp * p.g_p_off + rcur * p.g_ld + u0 + jbase;
#pragma unroll
        for (int u = 0; u < UT; ++u)
#pragma unroll
          for (int gt = 0; gt < 4; ++gt) acc[u][gt] = valid ? __ldg(gp + gt * H + u) : 0.f;
      }
      for (int kc0 = 0; kc0 < H; kc0 += LS_KC) {
        const int kc = min(LS_KC, H - kc0);
        // stage h(t-1) rows of this chunk (row index of t-1 is rcur-1; pad row when t == 0)
        __syncthreads();
        for (int i = tid; i < LS_ROWS * (kc / 4); i += nthr) {
          const int rl = i / (kc / 4), k4 = i % (kc / 4);
          const int qq = row0 + ch * LS_ROWS + rl;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (qq < row1) {
            const int p2 = qq / p.NB, b2 = qq % p.NB;
            v = ldcg4(p.hseq + ((int64_t)(m * 2 + p2) * R + (int64_t)b2 * Tp + t) * H + kc0 + k4 * 4);
          }
          *reinterpret_cast<float4*>(hs + rl * (LS_KC + 4) + k4 * 4) = v;
        }
        __syncthreads();
        const float* hrow = hs + lane * (LS_KC + 4);
        const float* wbase = Ws + (size_t)jbase * 4 * H + kc0;
#pragma unroll 2
        for (int k4 = 0; k4 < kc / 4; ++k4) {
          const float4 hv = *reinterpret_cast<const float4*>(hrow + k4 * 4);
#pragma unroll
          for (int u = 0; u < UT; ++u)
#pragma unroll
            for (int gt = 0; gt < 4; ++gt) {
              const float4 wv = *reinterpret_cast<const float4*>(wbase + (size_t)(u * 4 + gt) * H + k4 * 4);
              acc[u][gt] = fmaf(hv.x, wv.x, acc[u][gt]);
              acc[u][gt] = fmaf(hv.y, wv.y, acc[u][gt]);
              acc[u][gt] = fmaf(hv.z, wv.z, acc[u][gt]);
              acc[u][gt] = fmaf(hv.w, wv.w, acc[u][gt]);
            }
        }
      }
      if (valid) {
        float* hout = p.hseq + ((int64_t)(m * 2 + pp) * R + rcur) * H + u0 + jbase;
        float* cp = cs + (size_t)(ch * LS_ROWS + lane) * Hs + jbase;
#pragma unroll
        for (int u = 0; u < UT; ++u) {
          const float ig = sigmoid_f(acc[u][0]);
          const float fg = sigmoid_f(acc[u][1]);
          const float gg = tanhf(acc[u][2]);
          const float og = sigmoid_f(acc[u][3]);
          const float c = fg * cp[u] + ig * gg;
          cp[u] = c;
          const float hv = og * tanhf(c);
          hout[u] = hv;
          if (p.hsplit) st_split1(p.hsplit, 4 * R * H, (hout - p.hseq) + u, hv);
        }
      }
    }
    grid_barrier(p.sync, ++bar * nblk);
  }
}

// latent[b][t][j][part]: part 0 = h(re,x_re) - h(im,x_im), part 1 = h(re,x_im) + h(im,x_re)
__global__ void __launch_bounds__(256) lstm_combine_kernel(const float* __restrict__ hseq, int NB, int Talloc, int T,
                                                           int H, float* __restrict__ latent) {
  const int64_t n = (int64_t)NB * T * H;
  const int Tp = Talloc + 1;
  const int64_t R = (int64_t)NB * Tp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % H);
    const int64_t bt = i / H;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const int64_t row = ((int64_t)b * Tp + 1 + t) * H + j;
    const float rr = __ldg(hseq + 0 * R * H + row);   // (m=re, p=x_re)
    const float ir = __ldg(hseq + 1 * R * H + row);   // (m=re, p=x_im)
    const float ri = __ldg(hseq + 2 * R * H + row);   // (m=im, p=x_re)
    const float ii = __ldg(hseq + 3 * R * H + row);   // (m=im, p=x_im)
    *reinterpret_cast<float2*>(latent + i * 2) = make_float2(rr - ii, ir + ri);
  }
}

__global__ void __launch_bounds__(256) reparam_kernel(const float* __restrict__ latent, int NB, int T, int Htot,
                                                      int ch0, int zdim, int S, const float* __restrict__ eps_r,
                                                      const float* __restrict__ eps_i, uint64_t seed,
                                                      uint64_t offset, const unsigned long long* __restrict__ offset_dev,
                                                      int variant, float* __restrict__ z) {
  const int64_t n = (int64_t)NB * S * T * zdim;
  const float e = 1e-6f;
  if (offset_dev) offset += *offset_dev;               // device-side draw counter (CUDA-graph replays)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % zdim);
    int64_t rem = i / zdim;
    const int t = (int)(rem % T);
    rem /= T;
    const int s = (int)(rem % S);
    const int b = (int)(rem / S);
    (void)s;
    const float* lp = latent + (((int64_t)b * T + t) * Htot + ch0) * 2;
    const float2 mu = __ldg(reinterpret_cast<const float2*>(lp + (int64_t)j * 2));
    const float2 ls = __ldg(reinterpret_cast<const float2*>(lp + (int64_t)(zdim + j) * 2));
    const float2 dl = __ldg(reinterpret_cast<const float2*>(lp + (int64_t)(2 * zdim + j) * 2));
    float er, ei;
    if (eps_r) {
      er = __ldg(eps_r + i);
      ei = __ldg(eps_i + i);
    } else {
      curandStatePhilox4_32_10_t st;
      // Philox's `offset` counts single 32-bit outputs and curand_normal2 consumes two of them: one draw = one whole
      // Philox block (4 outputs), so consecutive draws (counter + 1) never share a uniform
      curand_init((unsigned long long)seed, (unsigned long long)i, 4ULL * (unsigned long long)offset, &st);
      const float2 nrm = curand_normal2(&st);
      er = nrm.x;
      ei = nrm.y;
    }
    // model/pvae_module.py:L2177-2231, same operation order; variant 1 = the *_fc_latent encoders (L2403-2450:
    // log sigma clamped to [-13, 13], square-root arguments clamped at eps, no eps in the denominators)
    const float sig = expf(variant ? fminf(fmaxf(ls.x, -13.f), 13.f) : ls.x);
    float dr = dl.x, di = dl.y;
    float ad = sqrtf(dr * dr + di * di + e);
    const float tmp = sig * 0.99f / (ad + e);
    if (ad >= sig - 1e-3f) {
      dr *= tmp;
      di *= tmp;
    }
    ad = sqrtf(dr * dr + di * di + e);
    const float num_r = sig + dr;
    float zr, zi;
    if (variant) {
      const float den = sqrtf(fmaxf(2.f * (sig + dr), e));
      const float sy = sqrtf(fmaxf(sig * sig - ad * ad, e)) / den;
      zr = mu.x + (num_r / den) * er;
      zi = mu.y + (di / den) * er + sy * ei;
    } else {
      const float den = sqrtf(2.f * (sig + dr) + e);
      const float sx = di / (den + e);
      const float sy = sqrtf(sig * sig - ad * ad + e) / (den + e);
      zr = mu.x + (num_r / (den + e)) * er;
      zi = mu.y + sx * er + sy * ei;
    }
    *reinterpret_cast<float2*>(z + i * 2) = make_float2(zr, zi);
  }
}

// reparameterisation of one element (model/pvae_module.py:L2177-2231, same operation order as reparam_kernel, variant 0)
__device__ __forceinline__ float2 reparam_one(float2 mu, float2 ls, float2 dl, float er, float ei) {
  const float e = 1e-6f;
  const float sig = expf(ls.x);
  float dr = dl.x, di = dl.y;
  float ad = sqrtf(dr * dr + di * di + e);
  const float tmp = sig * 0.99f / (ad + e);
  if (ad >= sig - 1e-3f) {
    dr *= tmp;
    di *= tmp;
  }
  ad = sqrtf(dr * dr + di * di + e);
  const float den = sqrtf(2.f * (sig + dr) + e);
  const float sx = di / (den + e);
  const float sy = sqrtf(sig * sig - ad * ad + e) / (den + e);
  return make_float2(mu.x + ((sig + dr) / (den + e)) * er, mu.y + sx * er + sy * ei);
}

// Fused latent stage: thread = (plane row r, latent channel j < zdim).  A valid row (b, t) combines the four LSTM
// streams into the 3 * latent_num latent values of channel j, writes them to `latent`, draws / reads eps, writes z of
// every sample and (latent 0) the decoder's z planes; pad rows and rows beyond the valid frames only zero the z planes.
__global__ void __launch_bounds__(256) latent_fused_kernel(const float* __restrict__ hseq, int NB, int Talloc, int Tv, int H,
                                                           int zdim, int latent_num, int S,
                                                           const float* __restrict__ eps_r0, const float* __restrict__ eps_i0,
                                                           const float* __restrict__ eps_r1, const float* __restrict__ eps_i1,
                                                           uint64_t seed, uint64_t offset,
                                                           const unsigned long long* __restrict__ offset_dev,
                                                           float* __restrict__ latent, float* __restrict__ z0,
                                                           float* __restrict__ z1, void* __restrict__ zplanes, int out_split,
                                                           int keep_pad) {
  pdl_trigger();
  pdl_wait();
  const int Tp = Talloc + 1;
  const int64_t R = (int64_t)NB * Tp;
  const int Ch = (zdim + 7) & ~7, Cp = 2 * Ch;
  const int64_t n = R * Ch;
  if (offset_dev) offset += *offset_dev;
  float* pf = reinterpret_cast<float*>(zplanes);
  unsigned short* ps = reinterpret_cast<unsigned short*>(zplanes);
  const int64_t plane_el = R * Cp;                       // elements of one sample's plane (per hi / lo half)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % Ch);
    const int64_t r = i / Ch;
    const int tt = (int)(r % Tp), b = (int)(r / Tp);
    const bool live = tt >= 1 && tt <= Tv && j < zdim;
    float2 zs0[8];                                        // z of latent 0 for up to 8 samples per pass
    for (int s0 = 0; s0 < S; s0 += 8) {
      const int ns = S - s0 < 8 ? S - s0 : 8;
      if (live) {
        const int t = tt - 1;
        for (int k = 0; k < latent_num; ++k) {
          float2 tri[3];
#pragma unroll
          for (int c3 = 0; c3 < 3; ++c3) {
            const int col = (3 * k + c3) * zdim + j;
            const int64_t idx = r * H + col;
            const float rr = __ldg(hseq + idx), ir = __ldg(hseq + R * H + idx);
            const float ri = __ldg(hseq + 2 * R * H + idx), ii = __ldg(hseq + 3 * R * H + idx);
            tri[c3] = make_float2(rr - ii, ir + ri);
            if (s0 == 0)
              *reinterpret_cast<float2*>(latent + (((int64_t)b * Tv + t) * H + col) * 2) = tri[c3];
          }
          const float* er_p = k ? eps_r1 : eps_r0;
          const float* ei_p = k ? eps_i1 : eps_i0;
          float* zk = k ? z1 : z0;
          for (int si = 0; si < ns; ++si) {
            const int s = s0 + si;
            const int64_t e_idx = (((int64_t)b * S + s) * Tv + t) * zdim + j;
            float er, ei;
            if (er_p) {
              er = __ldg(er_p + e_idx);
              ei = __ldg(ei_p + e_idx);
            } else {
              curandStatePhilox4_32_10_t st;
              curand_init((unsigned long long)seed, (unsigned long long)(e_idx + (int64_t)k * NB * S * Tv * zdim),
                          4ULL * (unsigned long long)offset, &st);
              const float2 nrm = curand_normal2(&st);
              er = nrm.x;
              ei = nrm.y;
            }
            const float2 zv = reparam_one(tri[0], tri[1], tri[2], er, ei);
            *reinterpret_cast<float2*>(zk + e_idx * 2) = zv;
            if (k == 0) zs0[si] = zv;
          }
        }
      }
      if (keep_pad && tt == 0) continue;                  // streaming: the pad row carries z of the previous step's last frame
      for (int si = 0; si < ns; ++si) {
        const float2 zv = live ? zs0[si] : make_float2(0.f, 0.f);
        const int64_t base = (int64_t)(s0 + si) * plane_el * (out_split ? 2 : 1) + r * Cp;
        if (out_split) {
          st_split1(ps, plane_el, base + j, zv.x);
          st_split1(ps, plane_el, base + Ch + j, zv.y);
        } else {
          pf[base + j] = zv.x;
          pf[base + Ch + j] = zv.y;
        }
      }
    }
  }
}

// lstm_combine_kernel + z_to_planes_kernel in one pass: thread = (plane row r, channel j)
__global__ void __launch_bounds__(256) combine_planes_kernel(const float* __restrict__ hseq, int NB, int Talloc, int Tv,
                                                             int H, float* __restrict__ latent, void* __restrict__ planes,
                                                             int out_split) {
  const int Tp = Talloc + 1;
  const int64_t R = (int64_t)NB * Tp;
  const int Ch = (H + 7) & ~7, Cp = 2 * Ch;
  const int64_t n = R * Ch;
  float* pf = reinterpret_cast<float*>(planes);
  unsigned short* ps = reinterpret_cast<unsigned short*>(planes);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % Ch);
    const int64_t r = i / Ch;
    const int tt = (int)(r % Tp), b = (int)(r / Tp);
    float2 v = make_float2(0.f, 0.f);
    if (tt >= 1 && tt <= Tv && j < H) {
      const int64_t idx = r * H + j;
      v = make_float2(__ldg(hseq + idx) - __ldg(hseq + 3 * R * H + idx), __ldg(hseq + R * H + idx) + __ldg(hseq + 2 * R * H + idx));
      *reinterpret_cast<float2*>(latent + (((int64_t)b * Tv + (tt - 1)) * H + j) * 2) = v;
    }
    if (out_split) {
      st_split1(ps, R * Cp, r * Cp + j, v.x);
      st_split1(ps, R * Cp, r * Cp + Ch + j, v.y);
    } else {
      pf[r * Cp + j] = v.x;
      pf[r * Cp + Ch + j] = v.y;
    }
  }
}

// nn.LSTM with one hidden unit: thread = utterance, all layers advanced step by step (exact expf / tanhf)
__global__ void __launch_bounds__(64) lstm_h1_kernel(const float* __restrict__ g, int g_ld, const float* __restrict__ wrec,
                                                     int L, int NB, int Talloc, int Tv, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  float w[4][12], h[4] = {0.f, 0.f, 0.f, 0.f}, c[4] = {0.f, 0.f, 0.f, 0.f};
  for (int l = 0; l < L; ++l)
    for (int k = 0; k < 12; ++k) w[l][k] = __ldg(wrec + l * 12 + k);
  const int Tp = Talloc + 1;
  for (int t = 0; t < Tv; ++t) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + ((long long)b * Tp + 1 + t) * g_ld));
    float x = 0.f;
    for (int l = 0; l < L; ++l) {
      float a[4];
      if (l == 0) {
        a[0] = g0.x; a[1] = g0.y; a[2] = g0.z; a[3] = g0.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = w[l][k] * x + w[l][8 + k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] += w[l][4 + k] * h[l];
      const float ig = sigmoid_f(a[0]), fg = sigmoid_f(a[1]), gg = tanhf(a[2]), og = sigmoid_f(a[3]);
      c[l] = fg * c[l] + ig * gg;
      h[l] = og * tanhf(c[l]);
      x = h[l];
    }
    out[(long long)b * Tv + t] = x;
  }
}

struct RecCfg {
  int NU, NR, Hs, RC, UT;
  size_t smem;
};

static bool pick_rec_cfg(int H, int NB, int sms, size_t smem_max, RecCfg* out) {
  const int nrows = 2 * NB;
  const int per_module = sms / 2;
  bool found = false;
  long best_cost = 0;
  RecCfg best{};
  for (int NU = 1; NU <= per_module && NU <= H; ++NU) {
    if (H % NU) continue;
    const int Hs = H / NU;
    int UT = 0;
    const int uts[5] = {6, 4, 3, 2, 1};
    for (int i = 0; i < 5; ++i)
      if (Hs % uts[i] == 0 && Hs / uts[i] <= 32) { UT = uts[i]; break; }
    if (!UT) continue;
    for (int NR = 1; NR * NU <= per_module && NR <= nrows; ++NR) {
      const int RC = cdiv(nrows, NR);
      if (cdiv(nrows, RC) != NR) continue;            // no empty row slices
      const int nch = cdiv(RC, LS_ROWS);
      const size_t smem = ((size_t)Hs * 4 * H + (size_t)LS_ROWS * (LS_KC + 4) + (size_t)nch * LS_ROWS * Hs) * 4;
      if (smem > smem_max) continue;
      // padded MACs per step per CTA + staging traffic penalty + warp-count penalty
      long cost = (long)nch * LS_ROWS * Hs * 4 * H + (long)nch * LS_ROWS * H * 8;
      if (Hs / UT < 4) cost += cost / 4;
      if (!found || cost < best_cost) {
        found = true;
        best_cost = cost;
        best = RecCfg{NU, NR, Hs, RC, UT, smem};
      }
    }
  }
  *out = best;
  return found;
}

template <int UT>
static int launch_rec(const LstmRecParams& p, const RecCfg& c, cudaStream_t st) {
  IDV_CUDA(cudaFuncSetAttribute(lstm_rec_kernel<UT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
  dim3 grid(c.NU, c.NR, 2), block(32 * c.Hs / UT);
  void* args[] = {(void*)&p};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)lstm_rec_kernel<UT>, grid, block, args, c.smem, st);
  if (e == cudaErrorCooperativeLaunchTooLarge) {
    set_error("idv_lstm_recurrent_fwd: cooperative grid %dx%dx2 not co-resident", c.NU, c.NR);
    return IDV_E_RESOURCE;
  }
  if (e != cudaSuccess) {
    set_error("idv_lstm_recurrent_fwd: launch failed: %s", cudaGetErrorString(e));
    return IDV_E_CUDA;
  }
  return IDV_OK;
}

}  // namespace idv

extern "C" int idv_lstm_recurrent_fwd(const float* g, int64_t g_m_off, int64_t g_p_off, int g_ld, const float* whh,
                                      int NB, int T, int H, float* hseq, void* hsplit, unsigned int* sync,
                                      int t_valid, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(g && whh && hseq && sync, "idv_lstm_recurrent_fwd: null pointer");
  IDV_CHECK_ARG(NB > 0 && T > 0 && H > 0 && H % 4 == 0, "idv_lstm_recurrent_fwd: need H %% 4 == 0 (NB=%d T=%d H=%d)", NB, T, H);
  int dev = 0, sms = 0, smem_optin = 0;
  IDV_CUDA(cudaGetDevice(&dev));
  IDV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  IDV_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  RecCfg c;
  IDV_CHECK_ARG(pick_rec_cfg(H, NB, sms, (size_t)smem_optin, &c), "idv_lstm_recurrent_fwd: no launch config for H=%d NB=%d", H, NB);
  LstmRecParams p;
  p.g = g; p.g_m_off = g_m_off; p.g_p_off = g_p_off; p.g_ld = g_ld; p.whh = whh;
  p.NB = NB; p.T = T; p.H = H; p.Tsteps = (t_valid > 0 && t_valid < T) ? t_valid : T; p.hseq = hseq; p.hsplit = reinterpret_cast<unsigned short*>(hsplit); p.sync = sync;
  p.NU = c.NU; p.NR = c.NR; p.Hs = c.Hs; p.RC = c.RC;
  cudaStream_t st = (cudaStream_t)stream;
  IDV_CUDA(cudaMemsetAsync(sync, 0, 2 * sizeof(unsigned int), st));
  switch (c.UT) {
    case 6: return launch_rec<6>(p, c, st);
    case 4: return launch_rec<4>(p, c, st);
    case 3: return launch_rec<3>(p, c, st);
    case 2: return launch_rec<2>(p, c, st);
    default: return launch_rec<1>(p, c, st);
  }
}

extern "C" int idv_lstm_combine_fwd(const float* hseq, int NB, int T, int H, float* latent, int t_valid,
                                    void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(hseq && latent && NB > 0 && T > 0 && H > 0, "idv_lstm_combine_fwd: bad argument");
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  const int64_t n = (int64_t)NB * Tv * H;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  lstm_combine_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(hseq, NB, T, Tv, H, latent);
  IDV_LAUNCH_CHECK("lstm_combine_kernel");
  return IDV_OK;
}

extern "C" int idv_reparam_fwd(const float* latent, int NB, int T, int Htot, int ch0, int zdim, int S,
                               const float* eps_r, const float* eps_i, uint64_t seed, uint64_t offset,
                               const uint64_t* offset_dev, int variant, float* z, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(latent && z && NB > 0 && T > 0 && S > 0 && zdim > 0 && (variant == 0 || variant == 1),
                "idv_reparam_fwd: bad argument");
  IDV_CHECK_ARG(ch0 >= 0 && ch0 + 3 * zdim <= Htot, "idv_reparam_fwd: latent slice out of range");
  IDV_CHECK_ARG((eps_r == nullptr) == (eps_i == nullptr), "idv_reparam_fwd: supply both eps tensors or neither");
  const int64_t n = (int64_t)NB * S * T * zdim;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  reparam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(latent, NB, T, Htot, ch0, zdim, S, eps_r, eps_i, seed,
                                                           offset, reinterpret_cast<const unsigned long long*>(offset_dev),
                                                           variant, z);
  IDV_LAUNCH_CHECK("reparam_kernel");
  return IDV_OK;
}

extern "C" int idv_latent_fwd(const float* hseq, int NB, int T, int H, int t_valid, int zdim, int latent_num, int S,
                              const float* eps_r0, const float* eps_i0, const float* eps_r1, const float* eps_i1,
                              uint64_t seed, uint64_t offset, const uint64_t* offset_dev, float* latent, float* z0,
                              float* z1, void* zplanes, int out_split, int keep_pad, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(hseq && latent && z0 && zplanes && NB > 0 && T > 0 && S > 0 && zdim > 0, "idv_latent_fwd: bad argument");
  IDV_CHECK_ARG((latent_num == 1 || latent_num == 2) && H == 3 * zdim * latent_num,
                "idv_latent_fwd: H = %d must equal 3 * zdim (%d) * latent_num (%d)", H, zdim, latent_num);
  IDV_CHECK_ARG((latent_num == 2) == (z1 != nullptr), "idv_latent_fwd: z1 is the second latent's output");
  IDV_CHECK_ARG((eps_r0 == nullptr) == (eps_i0 == nullptr) && (eps_r1 == nullptr) == (eps_i1 == nullptr) &&
                    (latent_num == 1 ? eps_r1 == nullptr : (eps_r1 == nullptr) == (eps_r0 == nullptr)),
                "idv_latent_fwd: supply eps for every latent (real and imaginary draw) or for none");
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  const int64_t n = (int64_t)NB * (T + 1) * ((zdim + 7) / 8 * 8);
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  IDV_CUDA(launch_pdl(latent_fused_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, hseq, NB, T, Tv, H, zdim,
                      latent_num, S, eps_r0, eps_i0, eps_r1, eps_i1, seed, offset,
                      reinterpret_cast<const unsigned long long*>(offset_dev), latent, z0, z1, zplanes, out_split,
                      keep_pad));
  return IDV_OK;
}

extern "C" int idv_lstm_combine_planes(const float* hseq, int NB, int T, int H, int t_valid, float* latent, void* planes,
                                       int out_split, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(hseq && latent && planes && NB > 0 && T > 0 && H > 0, "idv_lstm_combine_planes: bad argument");
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  const int64_t n = (int64_t)NB * (T + 1) * ((H + 7) / 8 * 8);
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  combine_planes_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(hseq, NB, T, Tv, H, latent, planes, out_split);
  IDV_LAUNCH_CHECK("combine_planes_kernel");
  return IDV_OK;
}

extern "C" int idv_lstm_h1_fwd(const float* g, int g_ld, const float* wrec, int num_layers, int NB, int T, int t_valid,
                               float* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(g && wrec && out && NB > 0 && T > 0, "idv_lstm_h1_fwd: bad argument");
  IDV_CHECK_ARG(num_layers >= 1 && num_layers <= 4 && g_ld >= 4 && g_ld % 4 == 0,
                "idv_lstm_h1_fwd: 1..4 layers and a gate row stride that is a multiple of 4 expected");
  const int Tv = (t_valid > 0 && t_valid < T) ? t_valid : T;
  lstm_h1_kernel<<<(NB + 63) / 64, 64, 0, (cudaStream_t)stream>>>(g, g_ld, wrec, num_layers, NB, T, Tv, out);
  IDV_LAUNCH_CHECK("lstm_h1_kernel");
  return IDV_OK;
}
