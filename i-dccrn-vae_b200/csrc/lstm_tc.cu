// Complex-LSTM recurrence on the tensor cores (one nn.LSTM layer of both modules per launch).
//
// Per time step and module m the recurrent GEMM is  gates[128 rows, 4H] += h(t-1)[128, H] . W_hh^T  with the 128
// rows = (input part p in {x_re, x_im}) x (64 utterances) — the two streams of a module share W_hh.  The 4H gate
// columns are split over NC CTAs: CTA c owns hidden units [c*Hs, (c+1)*Hs) x 4 gates (N = 4*Hs columns) and keeps
// its W_hh slice (bf16 hi/lo, K-major, 128B-swizzled) RESIDENT in shared memory for the whole sequence.
// Every step:
//   producer warp : waits until all NC CTAs of the module have published h(t-1), then streams it (bf16 hi/lo,
//                   [128 rows][H]) through a 32 KB-stage TMA ring;
//   MMA warp      : 3 x tcgen05.mma (h_lo*W_hi + h_hi*W_lo + h_hi*W_hi) per K step into one TMEM accumulator;
//   4 epilogue warps (thread = row): TMEM -> registers, add the pre-computed input projection (fp32, prefetched),
//                   gate non-linearities, cell update (cell state lives in registers for all T), write h(t) as
//                   bf16 hi/lo into the exchange buffer (+ the sequence outputs), then release the step counter.
// The exchange buffer hx[parity][m][hi/lo][128][H] (≈ 0.8 MB) stays in L2; the per-module step counter is the only
// inter-CTA synchronisation (NC CTAs, not the whole grid).  Cooperative launch guarantees co-residency.
#include <stdlib.h>

#include "tc_common.cuh"

namespace idv {
namespace tc {

constexpr int L_THREADS = 192;
constexpr int L_EPI_WARP0 = 2;
constexpr int L_ROWS = 128;                 // rows per row group: 2 parts x 64 utterances
constexpr int L_HTILE = L_ROWS * BK * 2;    // one 128 x 64 bf16 tile = 16 KB

struct LstmTcParams {
  const float* g;
  long long g_m_off, g_p_off;
  int g_ld;
  int NB, T, H, NC, KC, stages, Tsteps;     // KC = H / 64; stages = depth of the h TMA ring (<= 8); Tsteps <= T valid steps
  float* hseq;                              // optional fp32 [4][R][H]
  unsigned short* hsplit;                   // optional bf16 [2][4][R][H]
  unsigned short* hx;                       // bf16 [n_rg][2 parity][2 m][2 hl][128][H]
  unsigned int* sync;                       // [n_rg][2 m] step counters (zeroed by the host)
  unsigned long long* dbg;                  // optional phase timestamps (IDV_LSTM_DBG=1), CTA (0,0,0), steps [200,208)
  int rg0;                                  // first row group of this launch (batches whose row groups do not all fit the
                                            // device at once run as consecutive launches)
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define LSTM_DBG(slot)                                                                            \
  do {                                                                                            \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && t >= 200 && t < 208)   \
      p.dbg[(t - 200) * 16 + (slot)] = gtime();                                                   \
  } while (0)

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  // tanh(x) = 1 - 2/(exp(2x)+1): no cancellation for large |x|, abs error ~1e-7 near 0
  return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f);
}

template <int N>        // N = 4 * Hs gate columns per CTA
__global__ void __launch_bounds__(L_THREADS, 1)
lstm_rec_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH, const LstmTcParams p) {
  constexpr int HS = N / 4;
  constexpr int TMEM_COLS = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  constexpr int W_TILE = N * BK * 2;                       // one (hl, k-chunk) weight tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KC = p.KC;
  const int w_bytes = 2 * KC * W_TILE;
  const int stages = p.stages;              // as many 32 KB stages as fit next to the resident weights
  uint8_t* ring = smem + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + stages * 2 * L_HTILE);
  const uint32_t wfull = smem_u32(bars), hfull0 = wfull + 8, hempty0 = hfull0 + 8 * 8, accfull = hempty0 + 8 * 8,
                 accempty = accfull + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  const uint32_t smem_w = smem_u32(smem), smem_ring = smem_u32(ring);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x, m = blockIdx.y, rg = p.rg0 + blockIdx.z;
  const int NC = p.NC, H = p.H, T = p.Tsteps;
  const int Tp = p.T + 1;
  const long long R = (long long)p.NB * Tp;
  unsigned int* ctr = p.sync + rg * 2 + m;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmH);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(wfull, 1);
    for (int s = 0; s < 8; ++s) {
      mbar_init(hfull0 + 8 * s, 1);
      mbar_init(hempty0 + 8 * s, 1);
    }
    mbar_init(accfull, 1);
    mbar_init(accempty, 4);
    fence_barrier_init();
  }
  if (warp == L_EPI_WARP0) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      // resident W_hh slice: [hl][kc][N rows][64]
      mbar_expect_tx(wfull, (uint32_t)w_bytes);
      for (int hl = 0; hl < 2; ++hl)
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d(&tmW, wfull, smem_w + (hl * KC + kc) * W_TILE, kc * BK, ((hl * 2 + m) * NC + c) * N);
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < T; ++t) {
        // wait until every CTA of this (row group, module) has published h(t-1)
        const unsigned int target = (unsigned int)NC * (unsigned int)t;
        if (t > 0) {
          long long t0 = 0;
          unsigned int spins = 0;
          while (ld_acquire_gpu(ctr) < target) {
            if ((++spins & 255u) == 0) {
              const long long now = clock64();
              if (t0 == 0) t0 = now;
              else if (now - t0 > WAIT_TIMEOUT_CYCLES) __trap();
            }
          }
          fence_proxy_async_global();          // generic-proxy writes of the peers -> visible to the TMA reads
        }
        LSTM_DBG(0);
        const int par = t & 1;
        const int row_base = (((rg * 2 + par) * 2 + m) * 2) * L_ROWS;
        for (int kc0 = 0; kc0 < KC; ++kc0) {
          const int kc = (kc0 + c) % KC;       // CTAs walk the K chunks in rotated order (spreads the L2 requests)
          mbar_wait(hempty0 + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_ring + stage * 2 * L_HTILE;
          mbar_expect_tx(hfull0 + 8 * stage, 2 * L_HTILE);
          tma_load_2d(&tmH, hfull0 + 8 * stage, sa, kc * BK, row_base);
          tma_load_2d(&tmH, hfull0 + 8 * stage, sa + L_HTILE, kc * BK, row_base + L_ROWS);
          if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
        }
        LSTM_DBG(1);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(N);
      mbar_wait(wfull, 0);
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < T; ++t) {
        mbar_wait(accempty, (t & 1) ^ 1);
        tc_fence_after();
        for (int kc0 = 0; kc0 < KC; ++kc0) {
          const int kc = (kc0 + c) % KC;
          mbar_wait(hfull0 + 8 * stage, phase);
          tc_fence_after();
          if (kc0 == 0) LSTM_DBG(2);
          const uint32_t sa = smem_ring + stage * 2 * L_HTILE;
          const uint64_t a_hi = make_desc_sw128(sa), a_lo = make_desc_sw128(sa + L_HTILE);
          const uint64_t b_hi = make_desc_sw128(smem_w + kc * W_TILE);
          const uint64_t b_lo = make_desc_sw128(smem_w + (KC + kc) * W_TILE);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
            umma_bf16(tmem_base, a_lo + koff, b_hi + koff, idesc, (kc0 | k) != 0);
            umma_bf16(tmem_base, a_hi + koff, b_lo + koff, idesc, 1);
            umma_bf16(tmem_base, a_hi + koff, b_hi + koff, idesc, 1);
          }
          umma_commit(hempty0 + 8 * stage);
          if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(accfull);
        LSTM_DBG(3);
      }
    }
  } else {
    // ================================ gates / cell update (thread = row) ================================
    const int q = warp & 3;
    const int r = q * 32 + lane;                      // row in the group: part = r / 64, utterance = r % 64
    const int part = r >> 6;
    const int b = rg * 64 + (r & 63);
    const bool valid = b < p.NB;
    const int u0 = c * HS;
    const float* gbase = p.g + m * p.g_m_off + part * p.g_p_off + u0;
    float cst[HS];
#pragma unroll
    for (int j = 0; j < HS; ++j) cst[j] = 0.f;
    for (int t = 0; t < T; ++t) {
      const long long rcur = (long long)b * Tp + 1 + t;
      // prefetch the input projection of this step (independent of the recurrence)
      float gin[N];
      if (valid) {
        const float* gp = gbase + rcur * p.g_ld;
#pragma unroll
        for (int gt = 0; gt < 4; ++gt)
#pragma unroll
          for (int j = 0; j < HS; j += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(gp + gt * H + j));
            gin[gt * HS + j] = v.x; gin[gt * HS + j + 1] = v.y; gin[gt * HS + j + 2] = v.z; gin[gt * HS + j + 3] = v.w;
          }
      } else {
#pragma unroll
        for (int j = 0; j < N; ++j) gin[j] = 0.f;
      }
      mbar_wait(accfull, t & 1);
      tc_fence_after();
      if (warp == L_EPI_WARP0 && lane == 0) LSTM_DBG(4);
      uint32_t v[N];
#pragma unroll
      for (int c0 = 0; c0 < N; c0 += 16) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v + c0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accempty);           // accumulator drained -> next step's MMAs may start
      if (warp == L_EPI_WARP0 && lane == 0) LSTM_DBG(5);
      float hn[HS];
#pragma unroll
      for (int j = 0; j < HS; ++j) {
        const float ig = fast_sigmoid(__uint_as_float(v[j]) + gin[j]);
        const float fg = fast_sigmoid(__uint_as_float(v[HS + j]) + gin[HS + j]);
        const float gg = fast_tanh(__uint_as_float(v[2 * HS + j]) + gin[2 * HS + j]);
        const float og = fast_sigmoid(__uint_as_float(v[3 * HS + j]) + gin[3 * HS + j]);
        cst[j] = fg * cst[j] + ig * gg;
        hn[j] = og * fast_tanh(cst[j]);
      }
      if (valid) {
        // exchange buffer for step t+1: parity (t+1)&1
        const int par = (t + 1) & 1;
        unsigned short* hx = p.hx + ((((long long)(rg * 2 + par) * 2 + m) * 2) * L_ROWS + r) * H + u0;
        const long long hx_hl = (long long)L_ROWS * H;
#pragma unroll
        for (int j = 0; j < HS; j += 4)
          st_split4(hx, hx_hl, j, make_float4(hn[j], hn[j + 1], hn[j + 2], hn[j + 3]));
      }
      if (warp == L_EPI_WARP0 && lane == 0) LSTM_DBG(6);
      // publish h(t): only the exchange-buffer stores are on the critical path of the other CTAs
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == L_EPI_WARP0 && lane == 0) {
        LSTM_DBG(7);
        atomicAdd(ctr, 1u);
        LSTM_DBG(8);
      }
      if (valid) {                                    // sequence outputs, off the critical path
        const long long oidx = ((long long)(m * 2 + part) * R + rcur) * H + u0;
#pragma unroll
        for (int j = 0; j < HS; j += 4) {
          const float4 hv = make_float4(hn[j], hn[j + 1], hn[j + 2], hn[j + 3]);
          if (p.hsplit) st_split4(p.hsplit, 4 * R * H, oidx + j, hv);
          if (p.hseq) *reinterpret_cast<float4*>(p.hseq + oidx + j) = hv;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == L_EPI_WARP0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int N>
static int launch_lstm_tc(const CUtensorMap& mw, const CUtensorMap& mh, const LstmTcParams& p, int n_rg, size_t smem,
                          cudaStream_t st) {
  IDV_CUDA(cudaFuncSetAttribute(lstm_rec_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(p.NC, 2, n_rg), block(L_THREADS);
  void* args[] = {(void*)&mw, (void*)&mh, (void*)&p};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)lstm_rec_tc_kernel<N>, grid, block, args, smem, st);
  if (e == cudaErrorCooperativeLaunchTooLarge) {
    set_error("idv_lstm_recurrent_tc: cooperative grid %dx2x%d is not co-resident", p.NC, n_rg);
    return IDV_E_RESOURCE;
  }
  if (e != cudaSuccess) {
    set_error("idv_lstm_recurrent_tc: launch failed: %s", cudaGetErrorString(e));
    return IDV_E_CUDA;
  }
  return IDV_OK;
}

}  // namespace tc
}  // namespace idv

extern "C" int idv_lstm_tc_config(int H, int* n_cols, int* n_ctas) {
  // gate columns per CTA: W slice (N*H*4 B) + TMA ring must fit 227 KB; N % 16 == 0; H % (N/4) == 0
  using namespace idv;
  IDV_CHECK_ARG(n_cols && n_ctas, "idv_lstm_tc_config: null pointer");
  int N = 0;
  if (H % 64 == 0) {
    // N = 32 (8 hidden units per CTA) keeps the resident slice small so the TMA ring can hold most of h(t-1)
    // (the step is latency-bound on bytes in flight); H = 768 needs N = 48 to stay within 148 co-resident CTAs
    // option "lstm_ncols" = 64: fewer, fatter CTAs (48 per layer at H = 384) — leaves SMs to kernels of other streams
    if (idv::option_lstm_ncols() == 64 && H <= 384 && H % 16 == 0) N = 64;
    else if (H <= 512 && H % 8 == 0) N = 32;
    else if (H <= 768 && H % 12 == 0) N = 48;
  }
  IDV_CHECK_ARG(N > 0, "idv_lstm_tc_config: hidden size %d is not supported by the tensor-core recurrence", H);
  *n_cols = N;
  *n_ctas = H / (N / 4);
  return IDV_OK;
}

extern "C" int idv_lstm_recurrent_tc(const float* g, int64_t g_m_off, int64_t g_p_off, int g_ld, const void* wpack,
                                     int NB, int T, int H, float* hseq, void* hsplit, void* hx, unsigned int* sync,
                                     int t_valid, void* stream) {
  using namespace idv;
  using namespace idv::tc;
  IDV_CHECK_ARG(g && wpack && hx && sync && (hseq || hsplit), "idv_lstm_recurrent_tc: null pointer");
  IDV_CHECK_ARG(NB > 0 && T > 0, "idv_lstm_recurrent_tc: empty problem");
  int N = 0, NC = 0;
  int rc = idv_lstm_tc_config(H, &N, &NC);
  if (rc) return rc;
  const int n_rg = cdiv(NB, 64);
  const int KC = H / 64;
  int dev = 0, sms = 0, smem_optin = 0;
  IDV_CUDA(cudaGetDevice(&dev));
  IDV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  IDV_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t w_bytes = (size_t)2 * KC * N * BK * 2;
  int stages = (int)(((size_t)smem_optin - w_bytes - 1024 - 256) / (2 * L_HTILE));
  if (stages > 8) stages = 8;
  if (stages > KC) stages = KC;              // one step never has more than KC tiles in flight
  IDV_CHECK_ARG(stages >= 2 || (stages >= 1 && KC == 1), "idv_lstm_recurrent_tc: not enough shared memory for H=%d", H);
  const size_t smem = w_bytes + (size_t)stages * 2 * L_HTILE + 1024 + 256;
  IDV_CHECK_ARG((int)smem <= smem_optin, "idv_lstm_recurrent_tc: needs %zu B of shared memory", smem);
  IDV_CHECK_ARG(NC * 2 <= sms, "idv_lstm_recurrent_tc: %d CTAs per row group exceed the %d SMs", NC * 2, sms);
  const int rg_per_launch = sms / (NC * 2);           // row groups (64 utterances each) that are co-resident
  CUtensorMap mW, mH;
  rc = encode_map_2d(&mW, wpack, H, (uint64_t)2 * 2 * NC * N, BK, N);
  if (rc) return rc;
  rc = encode_map_2d(&mH, hx, H, (uint64_t)n_rg * 2 * 2 * 2 * L_ROWS, BK, L_ROWS);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  IDV_CUDA(cudaMemsetAsync(hx, 0, (size_t)n_rg * 2 * 2 * 2 * L_ROWS * H * 2, st));
  IDV_CUDA(cudaMemsetAsync(sync, 0, (size_t)n_rg * 2 * sizeof(unsigned int), st));
  LstmTcParams p;
  p.g = g; p.g_m_off = g_m_off; p.g_p_off = g_p_off; p.g_ld = g_ld;
  p.NB = NB; p.T = T; p.H = H; p.NC = NC; p.KC = KC; p.stages = stages;
  p.Tsteps = (t_valid > 0 && t_valid < T) ? t_valid : T;
  p.hseq = hseq; p.hsplit = reinterpret_cast<unsigned short*>(hsplit);
  p.hx = reinterpret_cast<unsigned short*>(hx); p.sync = sync;
  p.dbg = nullptr;
  const bool dbg = getenv("IDV_LSTM_DBG") != nullptr && p.Tsteps > 208;
  if (dbg) {
    IDV_CUDA(cudaMalloc(&p.dbg, 8 * 16 * sizeof(unsigned long long)));
    IDV_CUDA(cudaMemsetAsync(p.dbg, 0, 8 * 16 * sizeof(unsigned long long), st));
  }
  for (int rg0 = 0; rg0 < n_rg && rc == IDV_OK; rg0 += rg_per_launch) {
    p.rg0 = rg0;
    const int nz = n_rg - rg0 < rg_per_launch ? n_rg - rg0 : rg_per_launch;
    switch (N) {
      case 64: rc = launch_lstm_tc<64>(mW, mH, p, nz, smem, st); break;
      case 48: rc = launch_lstm_tc<48>(mW, mH, p, nz, smem, st); break;
      default: rc = launch_lstm_tc<32>(mW, mH, p, nz, smem, st); break;
    }
  }
  if (dbg && rc == IDV_OK) {
    unsigned long long h[8 * 16];
    IDV_CUDA(cudaStreamSynchronize(st));
    IDV_CUDA(cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    // slots: 0 step counter seen, 1 TMA issued, 2 first h tile landed, 3 last MMA issued, 4 accumulator ready,
    //        5 TMEM drained, 6 gates+stores done, 7 fence+bar done, 8 counter published
    for (int i = 0; i < 8; ++i) {
      fprintf(stderr, "[lstm_tc dbg] step %d:", 200 + i);
      for (int sl = 0; sl < 9; ++sl) fprintf(stderr, " %lld", (long long)(h[i * 16 + sl] - h[0]));
      fprintf(stderr, "\n");
    }
  }
  return rc;
}
