// tap-GEMM, fp32 SIMT version (exact-fp32 accumulate; used for parity bring-up, odd shapes and
// as the cross-check of the tcgen05 path).  See include/idv.h for the contract.
//
// Tiling: CTA = 128 rows x BN columns, K step 16, 256 threads, 8 x (BN/16) register tile per
// thread, double-buffered shared memory with register prefetch.  A rows are the activation rows
// (r - dt) of plane f_in: a 128 x 16 tile is 128 x 64 B row segments (the implicit-GEMM gather
// needs no im2col because the layout is [F][R][C] and the time tap is a row shift).
#include "idv_common.cuh"

namespace idv {

constexpr int TG_BM = 128;
constexpr int TG_BK = 16;
constexpr int TG_THREADS = 256;

struct TapGemmParams {
  const float* a[2];
  int a_ld[2];
  int64_t a_plane[2];
  int R, Tp, t_valid, keep_pad;
  const float* w;
  const float* bias;
  int N;
  const idv_unit_t* units;
  const idv_tap_t* taps;
  float* out;
  int out_ld;
  int64_t out_plane;
  int apply_prelu;
  float slope;
};

template <int BN>
__global__ void __launch_bounds__(TG_THREADS, 2) tapgemm_f32_kernel(const TapGemmParams p) {
  constexpr int TN = BN / 16;          // columns per thread (4 or 8)
  constexpr int NB4 = BN / 4;          // float4 per B row
  constexpr int B_LOADS = (TG_BK * NB4) / TG_THREADS;   // 1 (BN=64) or 2 (BN=128)
  __shared__ __align__(16) float As[2][TG_BK][TG_BM];
  __shared__ __align__(16) float Bs[2][TG_BK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int r0 = blockIdx.x * TG_BM;
  const int n0 = blockIdx.y * BN;
  const idv_unit_t unit = p.units[blockIdx.z];
  const idv_tap_t* taps = p.taps + unit.tap_begin;

  // A-load role: one row, two 16-byte k-quads
  const int a_row = tid & 127;
  const int a_kq = tid >> 7;            // 0/1 -> quads {a_kq, a_kq+2}
  // B-load role
  int b_k[B_LOADS], b_nq[B_LOADS];
#pragma unroll
  for (int i = 0; i < B_LOADS; ++i) {
    int idx = tid + i * TG_THREADS;
    b_k[i] = idx / NB4;
    b_nq[i] = idx % NB4;
  }

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // flattened K iteration state
  int tap_i = 0, k0 = 0;
  idv_tap_t tap = taps[0];
  int total_steps = 0;
  for (int t = 0; t < unit.n_taps; ++t) total_steps += (taps[t].kc + TG_BK - 1) / TG_BK;

  float4 a_reg[2], b_reg[B_LOADS];

  auto load_tiles = [&](void) {
    // A
    const int ra = r0 + a_row - tap.dt;
    const bool ok = (ra >= 0) && (ra < p.R) && (r0 + a_row < p.R);
    const bool s1 = tap.src != 0;
    const float* abase = s1 ? p.a[1] : p.a[0];
    const int64_t aplane = s1 ? p.a_plane[1] : p.a_plane[0];
    const int ald = s1 ? p.a_ld[1] : p.a_ld[0];
    const float* ap = abase + (int64_t)tap.f_in * aplane + (int64_t)ra * ald + tap.ch_off + k0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool kok = k0 + (a_kq + 2 * i) * 4 < tap.kc;      // kc is a multiple of 4, not of BK
      a_reg[i] = (ok && kok) ? ldg4(ap + (a_kq + 2 * i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // B
    const float* wp = p.w + tap.w_off + (int64_t)k0 * p.N + n0;
#pragma unroll
    for (int i = 0; i < B_LOADS; ++i) {
      const int n = n0 + b_nq[i] * 4;
      b_reg[i] = (n < p.N && k0 + b_k[i] < tap.kc) ? ldg4(wp + (int64_t)b_k[i] * p.N + b_nq[i] * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int kq = (a_kq + 2 * i) * 4;
      As[buf][kq + 0][a_row] = a_reg[i].x;
      As[buf][kq + 1][a_row] = a_reg[i].y;
      As[buf][kq + 2][a_row] = a_reg[i].z;
      As[buf][kq + 3][a_row] = a_reg[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LOADS; ++i)
      *reinterpret_cast<float4*>(&Bs[buf][b_k[i]][b_nq[i] * 4]) = b_reg[i];
  };
  auto advance = [&](void) {
    k0 += TG_BK;
    if (k0 >= tap.kc) {
      k0 = 0;
      ++tap_i;
      if (tap_i < unit.n_taps) tap = taps[tap_i];
    }
  };

  if (total_steps > 0) {
    load_tiles();
    store_tiles(0);
    advance();
  }
  __syncthreads();

  for (int s = 0; s < total_steps; ++s) {
    const int buf = s & 1;
    const bool more = (s + 1 < total_steps);
    if (more) load_tiles();
#pragma unroll
    for (int k = 0; k < TG_BK; ++k) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      if (TN == 8) {
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][(BN / 2) + tx * 4]);
        b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tiles(buf ^ 1);
      advance();
    }
    __syncthreads();
  }

  // epilogue: bias + PReLU, pad rows forced to zero
  float* outp = p.out + (int64_t)unit.out_f * p.out_plane + unit.out_ch_off;
  const float* bias = p.bias + unit.bias_off;
#pragma unroll
  for (int cg = 0; cg < TN / 4; ++cg) {
    const int n = n0 + cg * (BN / 2) + tx * 4;
    if (n >= p.N) continue;
    const float4 bv = ldg4(bias + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (r >= p.R) continue;
      float4 v;
      v.x = acc[i][cg * 4 + 0] + bv.x;
      v.y = acc[i][cg * 4 + 1] + bv.y;
      v.z = acc[i][cg * 4 + 2] + bv.z;
      v.w = acc[i][cg * 4 + 3] + bv.w;
      if (p.apply_prelu) {
        v.x = prelu_f(v.x, p.slope); v.y = prelu_f(v.y, p.slope);
        v.z = prelu_f(v.z, p.slope); v.w = prelu_f(v.w, p.slope);
      }
      if (p.Tp > 0) {                              // causal pad row, or frame beyond the valid length
        const int tt = r % p.Tp;
        if (tt == 0 && p.keep_pad) continue;       // streaming: the pad row carries x[t-1] of the last step
        if (tt == 0 || (p.t_valid > 0 && tt > p.t_valid)) v = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      *reinterpret_cast<float4*>(outp + (int64_t)r * p.out_ld + n) = v;
    }
  }
}

}  // namespace idv

extern "C" int idv_tapgemm_f32(const float* a0, int a0_ld, int64_t a0_plane, const float* a1, int a1_ld,
                               int64_t a1_plane, int R, int Tp, const float* w, const float* bias, int N,
                               const idv_unit_t* units, const idv_tap_t* taps, int n_units, float* out,
                               int out_ld, int64_t out_plane, int apply_prelu, float prelu_slope,
                               int t_valid, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(a0 && w && bias && units && taps && out, "idv_tapgemm_f32: null pointer");
  IDV_CHECK_ARG(R > 0 && N > 0 && n_units > 0, "idv_tapgemm_f32: empty problem R=%d N=%d units=%d", R, N, n_units);
  IDV_CHECK_ARG(N % 4 == 0 && a0_ld % 4 == 0 && out_ld % 4 == 0 && (a1 == nullptr || a1_ld % 4 == 0),
                "idv_tapgemm_f32: N/ld must be multiples of 4 (N=%d a0_ld=%d out_ld=%d)", N, a0_ld, out_ld);
  IDV_CHECK_ARG(n_units <= 65535, "idv_tapgemm_f32: too many units (%d)", n_units);
  TapGemmParams p;
  p.a[0] = a0; p.a[1] = a1 ? a1 : a0;
  p.a_ld[0] = a0_ld; p.a_ld[1] = a1 ? a1_ld : a0_ld;
  p.a_plane[0] = a0_plane; p.a_plane[1] = a1 ? a1_plane : a0_plane;
  p.R = R; p.Tp = Tp < 0 ? -Tp : Tp; p.keep_pad = Tp < 0; p.t_valid = t_valid; p.w = w; p.bias = bias; p.N = N; p.units = units; p.taps = taps;
  p.out = out; p.out_ld = out_ld; p.out_plane = out_plane; p.apply_prelu = apply_prelu; p.slope = prelu_slope;
  cudaStream_t st = (cudaStream_t)stream;
  if (N % 128 == 0) {
    dim3 grid(cdiv(R, TG_BM), N / 128, n_units);
    tapgemm_f32_kernel<128><<<grid, TG_THREADS, 0, st>>>(p);
  } else {
    dim3 grid(cdiv(R, TG_BM), cdiv(N, 64), n_units);
    tapgemm_f32_kernel<64><<<grid, TG_THREADS, 0, st>>>(p);
  }
  IDV_LAUNCH_CHECK("tapgemm_f32_kernel");
  return IDV_OK;
}
