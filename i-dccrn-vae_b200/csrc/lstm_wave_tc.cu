// Two-layer complex-LSTM recurrence as ONE wavefront kernel (tcgen05): layer 0, the layer-1 input projection and
// layer 1 run concurrently, skewed by one time step each, so the sequential depth is T+2 steps instead of 2T and
// the layer-1 gate pre-activations never go to HBM.
//
// CTA roles (grid = NC x (2 modules * 3 roles), one CTA per SM, cooperative launch):
//   L0 : h0(t) = cell(G0(t) + h0(t-1) W_hh0^T)            reads hxA, writes hxA            publishes counter A
//   IP : G1(t) = h0(t) W_ih1^T + (b_ih1 + b_hh1)          reads hxA, writes G1x (fp32)     publishes counter B
//   L1 : h1(t) = cell(G1(t) + h1(t-1) W_hh1^T)            reads hxC + G1x, writes hxC, hseq publishes counter C
// Every role is the same pipeline as csrc/lstm_tc.cu: its N = 4*Hs gate columns of the weight matrix resident in
// shared memory (bf16 hi/lo), the 128 x H input rows streamed by TMA from an L2-resident exchange buffer, 3 MMAs
// per K step into TMEM, thread = row epilogue.  Exchange buffers are 4 deep in time (hxA, G1x) / 2 deep (hxC) and
// the per-(module, role) step counters (one increment per CTA and step) carry both the data dependencies and the
// buffer-reuse back-pressure:
//   L0(t) waits A >= NC*t           and B >= NC*(t-3)  (slot t%4 of hxA was last read by IP(t-4))
//   IP(t) waits A >= NC*(t+1)       and C >= NC*(t-3)  (slot t%4 of G1x was last read by L1(t-4))
//   L1(t) waits C >= NC*t           and B >= NC*(t+1)
// PAIR (default): neighbouring CTAs of one (module, role) run as a CTA pair (cluster of 2, tcgen05 cta_group::2, M = 128
// = 64 rows of each CTA, N = 128 = the 64 gate columns of each CTA): a CTA streams only ITS 64 rows of h (98 KB instead
// of 196 KB per step - the streaming phase is bound by the ~68 GB/s of TMA ingest one SM gets) and its MMAs read 2 KB of
// A instead of 4 KB.  The even CTA issues the MMAs; a CTA's TMEM then holds its 64 rows x 128 columns as lanes 0-63 =
// columns 0-63 (the even CTA's hidden units), lanes 64-127 = columns 64-127 (the odd CTA's), so epilogue thread L works
// on row 64*rank + L%64 and the units of CTA (c & ~1) + L/64.
#include <stdlib.h>

#include <type_traits>

#include "tc_common.cuh"

namespace idv {
namespace tc {

// epilogue warps: a thread = (row, HS / (warps / 4) of the CTA's hidden units): 8 warps = 8 units per thread at N = 64, 6 at
// N = 48.  Measured on B200 with 16 warps at N = 64 (-DIDV_WAVE_EPI_WARPS_N64=16): the gate math + stores of a step got
// shorter (1.28 -> 1.02 us) but the publish (named barrier of 512 threads + fence + counter) longer (1.15 -> 1.9 us): step
// 7.2 -> 8.2 us, config-2 LSTM 5.9 -> 6.5 ms (profiles/r02_lstm_dbg_s_16_epilogue_warps.log).  Stays 8.
constexpr int W_EPI_WARP0 = 2;
#ifndef IDV_WAVE_EPI_WARPS_N64
#define IDV_WAVE_EPI_WARPS_N64 8
#endif
constexpr int W_EPI_WARPS_N64 = IDV_WAVE_EPI_WARPS_N64;
__host__ __device__ constexpr int wave_epi_warps(int n_cols) { return n_cols == 64 ? W_EPI_WARPS_N64 : 8; }
__host__ __device__ constexpr int wave_threads(int n_cols) { return 64 + 32 * wave_epi_warps(n_cols); }
constexpr int W_ROWS = 128;
constexpr int W_HTILE = W_ROWS * BK * 2;
constexpr int W_SYNC_STRIDE = 32;            // uint32 between the step counters (one 128-byte line each)
constexpr int W_REP = 1;                     // replicas of the h exchange buffers (measured on B200: replication does not help)

struct WaveParams {
  const float* g0;
  long long g_m_off, g_p_off;
  int g_ld;
  const float* bias1;                       // [2 m][NC][N] CTA-major (gate*Hs + j)
  int NB, T, H, NC, KC, stages, Tsteps;     // T = allocated frames (row layout), Tsteps = valid steps
  // A launch works on ONE or TWO chunks of <= 64 utterances: chunk ch = utterances [b0[ch], b0[ch] + NBc[ch]).  Two chunks
  // are two independent recurrences INTERLEAVED in the same CTAs (same resident weights; own exchange buffers, step
  // counters, TMEM accumulator and cell state): while the h(t) of one chunk travels through the L2 (publish +
  // propagation: 3.5 of the 8 us of a step) the CTA runs step t of the other chunk.
  int nch, b0[2], NBc[2];
  long long hxA_ch, hxC_ch, g1x_ch;         // elements between the two chunks' exchange buffers
  int blkA_ch, blkC_ch;                     // ... and the same distance in blocks of the tensor maps
  int sync_mode;                            // 0: acquire polls, fence + atomic; 1: relaxed polls + one fence, red.release; 2: acquire polls,
                                            // red.release (option lstm_sync_mode; 1 and 2 measured slower than 0)
  int n_roles;                              // 3 = two-layer wavefront (L0 | IP | L1), 1 = ONE nn.LSTM layer per launch (role L0 only)
  unsigned short* hsplit;                   // single-layer mode: optional bf16 [2][4][R][H] output (next layer's in-proj input)
  float* hseq0;                             // single-layer mode: optional fp32 [4][R][H] output
  float* hseq1;                             // fp32 [4][R][H] layer-1 output
  unsigned short* hxA;                      // bf16 [W_REP][4 slot][2 m][2 hl][128][H]   h0
  unsigned short* hxC;                      // bf16 [W_REP][2 slot][2 m][2 hl][128][H]   h1
  float* g1x;                               // fp32 [4 slot][2 m][128][4H]        layer-1 gate pre-activations
  unsigned int* sync;                       // [2 m][3] counters A, B, C, W_SYNC_STRIDE uint32 apart
  // single-layer mode (n_roles = 1), optional: one step counter per 64-wide K CHUNK of h instead of one per module,
  // [2 chunks of utterances][2 m][KC] x W_SYNC_STRIDE.  A CTA increments the counter(s) of the chunk(s) its hidden units
  // fall into; a consumer polls all KC counters at once (one lane each) and loads chunk k as soon as ITS ~6 producers
  // have published, instead of waiting for all 64 CTAs of the module: the 64 same-address atomics no longer serialise
  // (H = 768: published -> seen by the next step took 3.6 of 11.2 us) and the ingest overlaps the stragglers' publish.
  unsigned int* kcsync;
  // how h(t) leaves the CTA (option "lstm_tma_publish"): 0 = every epilogue thread stores its units of its row (4-byte
  // stores at N = 48: 1 536 partial-sector writes per CTA and step, whose drain the publishing fence waits for); 2 = the
  // block is staged in shared memory and leaves with 16-byte stores (384 per CTA and step): step 10.9 -> 9.7 us at H = 768
  // (config 2b 46.1 -> 44.4 ms), but 7.2 -> 8.2 us at N = 64 / H = 384, whose direct stores are 16 bytes already (the extra
  // barrier costs more than it saves) - so auto = 2 for the 48-column one-layer kernel only; 1 = staged + ONE tensor store
  // per CTA and step, the publisher waiting for the bulk group: a box of 128 rows x 48 bytes takes the TMA unit ~4.5 us
  // (step 11.1 us at H = 768, 8.5 us at H = 384) - measured, kept as a switch.  profiles/r02_lstm_dbg_*_{tma_publish,
  // staged_stores}.log
  int tma_pub;
  unsigned long long* dbg;                  // optional phase timestamps (IDV_LSTM_DBG): CTA 0 of each role, module 0
};

__device__ __forceinline__ unsigned long long wgtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// slots per (role, step): 0 deps satisfied, 1 loads issued, 2 first tile landed, 3 MMAs issued, 4 accumulator ready,
// 5 TMEM drained, 6 stores done, 7 published
#define WAVE_DBG(slot)                                                                      \
  do {                                                                                      \
    if (p.dbg && blockIdx.x == 0 && m == 0 && ch == 0 && t >= 300 && t < 304)               \
      p.dbg[(role * 4 + (t - 300)) * 8 + (slot)] = wgtime();                                \
  } while (0)

__device__ __forceinline__ unsigned int ldacq(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ldrlx(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// relaxed: spin with relaxed loads and acquire once at the end (experiment: an acquire load per spin iteration is a
// load + fence)
__device__ __forceinline__ void wait_counter(const unsigned int* ctr, long long target, bool relaxed = false) {
  if (target <= 0) return;
  long long t0 = 0;
  unsigned int spins = 0;
  while ((long long)(relaxed ? ldrlx(ctr) : ldacq(ctr)) < target) {
    if ((++spins & 255u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > WAIT_TIMEOUT_CYCLES) __trap();
    }
  }
  if (relaxed) asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// publish one step: release the CTA's h stores (the caller's warps met at a named barrier) and count the CTA in
__device__ __forceinline__ void publish_step(unsigned int* ctr, int mode) {
  if (mode == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
  } else {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
  }
}
__device__ __forceinline__ float wsig(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float wtanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

template <int N, bool PAIR>
__global__ void __launch_bounds__(wave_threads(N), 1)
lstm_wave_tc_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmWi,
                    const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmHA,
                    const __grid_constant__ CUtensorMap tmHC, const __grid_constant__ CUtensorMap tmSA,
                    const __grid_constant__ CUtensorMap tmSC, const WaveParams p) {
  constexpr int HS = N / 4;
  constexpr int W_EPI_WARPS = wave_epi_warps(N);
  // PAIR: the split runs as TWO MMAs per K step, A_hi x [W_hi | W_lo] (width 2 * 2N: the hi and lo weight tiles of a K
  // chunk are adjacent in shared memory) and A_lo x W_hi on top of its first half; the epilogue adds the two halves
  constexpr int ACC_COLS = PAIR ? 2 * N : N;
  constexpr int ACC_STRIDE = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : 128);     // TMEM columns between the chunks' accumulators
  constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static_assert(ACC_COLS <= 128, "two accumulators must fit the 512 TMEM columns (and the allocation a power of two)");
  constexpr int W_TILE = N * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KC = p.KC, stages = p.stages;
  const int w_bytes = 2 * KC * W_TILE;
  uint8_t* ring = smem + w_bytes;
  constexpr int ROWS_C = PAIR ? W_ROWS / 2 : W_ROWS;          // rows of h this CTA streams
  constexpr int HT = ROWS_C * BK * 2;                         // bytes of the hi (or lo) rows of one K chunk
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  constexpr int UW = (PAIR ? 2 : 1) * HS;                     // hidden units this CTA stores per row (its pair's)
  constexpr int STG = 2 * ROWS_C * UW * 2;                    // staging of one chunk's h block: [hi | lo][rows][UW] bf16
  uint8_t* stg = ring + stages * 2 * HT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + (p.tma_pub ? 2 * STG : 0));
  const uint32_t wfull = smem_u32(bars), hfull0 = wfull + 8, hempty0 = hfull0 + 8 * 8, accfull0 = hempty0 + 8 * 8,
                 accempty0 = accfull0 + 16;                   // accumulator barriers: one pair per chunk
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  const uint32_t smem_w = smem_u32(smem), smem_ring = smem_u32(ring);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x;
  const int m = blockIdx.y / p.n_roles, role = blockIdx.y % p.n_roles;        // 0 = L0, 1 = IP, 2 = L1
  const int NC = p.NC, H = p.H, T = p.Tsteps;
  const int Tp = p.T + 1;
  const long long R = (long long)p.NB * Tp;
  // every counter on a 128-byte line of its own: the pollers of one counter (ld.acquire of ~48 producer threads) and
  // the increments of another must not queue up on the same L2 line
  // counters of chunk ch: p.sync + ch * 6 * W_SYNC_STRIDE + (m * 3 + {A, B, C}) * W_SYNC_STRIDE
  unsigned int* const sync_m = p.sync + m * 3 * W_SYNC_STRIDE;
  constexpr int CH_SYNC = 6 * W_SYNC_STRIDE;
  const int nch = p.nch;
  const bool rlx = p.sync_mode == 1;
  const CUtensorMap* tmW = role == 0 ? &tmW0 : (role == 1 ? &tmWi : &tmW1);
  const CUtensorMap* tmH = role == 2 ? &tmHC : &tmHA;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(tmW);
    prefetch_tmap(tmH);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(wfull, 1);
    for (int s = 0; s < 8; ++s) {
      mbar_init(hfull0 + 8 * s, 1);
      mbar_init(hempty0 + 8 * s, 1);
    }
    for (int ch = 0; ch < 2; ++ch) {
      mbar_init(accfull0 + 8 * ch, 1);
      mbar_init(accempty0 + 8 * ch, PAIR ? 2 * W_EPI_WARPS : W_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == W_EPI_WARP0) {
    if (PAIR) {
      tmem_alloc_2sm(smem_u32(tmem_slot), TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // (whole warp in the loop, TMA instructions under elect.sync: see the MMA issuer)
    {
      if (elect_one_sync()) {
        // PAIR: the barriers the MMA issuer waits on live in the even CTA and collect the bytes of both CTAs' loads
        if (!PAIR || rank == 0) mbar_expect_tx(wfull, (uint32_t)w_bytes * (PAIR ? 2 : 1));
        for (int hl = 0; hl < 2; ++hl)
          for (int kc = 0; kc < KC; ++kc) {
            if (PAIR)      // [kc][hi | lo]
              tma_load_2d_2sm(tmW, wfull & PEER_BIT_MASK, smem_w + (kc * 2 + hl) * W_TILE, kc * BK, ((hl * 2 + m) * NC + c) * N);
            else
              tma_load_2d(tmW, wfull, smem_w + (hl * KC + kc) * W_TILE, kc * BK, ((hl * 2 + m) * NC + c) * N);
          }
      }
      __syncwarp();
      // (per-chunk counters) CTAs whose published units fall into K chunk `lane`: a CTA publishes the units of its pair
      int kc_cnt = 0;
      if (p.n_roles == 1 && p.kcsync != nullptr && lane < KC)
        for (int cc = 0; cc < NC; ++cc) {
          const int ulo = (PAIR ? (cc & ~1) : cc) * HS, uhi = ulo + (PAIR ? 2 : 1) * HS - 1;
          if ((ulo >> 6) <= lane && lane <= (uhi >> 6)) ++kc_cnt;
        }
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < T; ++t)
      for (int ch = 0; ch < nch; ++ch) {
        const unsigned int* cA = sync_m + ch * CH_SYNC;
        const unsigned int* cB = cA + W_SYNC_STRIDE;
        const unsigned int* cC = cB + W_SYNC_STRIDE;
        int slot;
        const bool by_chunk = p.n_roles == 1 && p.kcsync != nullptr;
        if (lane == 0 && !by_chunk) {
          if (role == 0) {            // input h0(t-1): slot (t)%4 holds h0(t-1) (h0(t) is written to slot (t+1)%4)
            wait_counter(cA, (long long)NC * t, rlx);
            if (p.n_roles == 3) wait_counter(cB, (long long)NC * (t - 3), rlx);   // (single layer: all readers of a slot are L0 CTAs,
          } else if (role == 1) {     // input h0(t): slot (t+1)%4                //  at most one step apart)
            wait_counter(cA, (long long)NC * (t + 1), rlx);
            wait_counter(cC, (long long)NC * (t - 3), rlx);
          } else {                    // input h1(t-1): slot t%2
            wait_counter(cC, (long long)NC * t, rlx);
            wait_counter(cB, (long long)NC * (t + 1), rlx);
          }
          WAVE_DBG(0);
        }
        __syncwarp();
        fence_proxy_async_global();
        slot = role == 0 ? (t & 3) : (role == 1 ? ((t + 1) & 3) : (t & 1));
        const int nslot = role == 2 ? 2 : 4;
        const int blk = ((((c % W_REP) * nslot + slot) * 2 + m) * 2) +     // block index of the hi tile (lo = +1)
                        ch * (role == 2 ? p.blkC_ch : p.blkA_ch);
        // by_chunk: lane k polls the counter of K chunk k; chunk k is loaded as soon as it and every chunk before it is ready
        const unsigned int* kb = by_chunk ? p.kcsync + (long long)((ch * 2 + m) * KC) * W_SYNC_STRIDE : nullptr;
        const long long kc_target = (long long)kc_cnt * t;
        uint32_t ready = by_chunk ? 0u : 0xffffffffu;
        long long poll_t0 = 0;
        unsigned int polls = 0;
        for (int kc0 = 0; kc0 < KC; ++kc0) {
          const int kc = kc0;                  // same order in every CTA (measured: rotating the order does not help)
          while (!((ready >> kc) & 1u)) {
            bool ok = true;
            if (lane < KC && !((ready >> lane) & 1u)) ok = kc_target <= 0 || (long long)ldacq(kb + lane * W_SYNC_STRIDE) >= kc_target;
            ready = __ballot_sync(0xffffffffu, ok);
            if ((ready >> kc) & 1u) {
              fence_proxy_async_global();
              if (lane == 0 && kc == 0) WAVE_DBG(0);
            } else if ((++polls & 63u) == 0) {
              const long long now = clock64();
              if (poll_t0 == 0) poll_t0 = now;
              else if (now - poll_t0 > WAIT_TIMEOUT_CYCLES) __trap();
            }
          }
          mbar_wait(hempty0 + 8 * stage, phase ^ 1);
          if (elect_one_sync()) {
            const uint32_t sa = smem_ring + stage * 2 * HT;
            if (PAIR) {
              if (rank == 0) mbar_expect_tx(hfull0 + 8 * stage, 4 * HT);
              tma_load_3d_2sm(tmH, (hfull0 + 8 * stage) & PEER_BIT_MASK, sa, kc * BK, (int)rank * ROWS_C, blk);   // my 64 rows
            } else {
              mbar_expect_tx(hfull0 + 8 * stage, 2 * HT);
              tma_load_3d(tmH, hfull0 + 8 * stage, sa, kc * BK, 0, blk);      // {64 k, 128 rows, hi+lo} = 32 KB
            }
            if (kc0 + 1 == KC) WAVE_DBG(1);
          }
          __syncwarp();
          if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // the whole warp runs the loop, the tcgen05 instructions are issued under elect.sync (from a `lane == 0` branch ptxas
    // wraps every tcgen05.mma in an elect / R2UR / branch loop: ~37 ns of issue per MMA, measured in lstm_cluster_tc.cu)
    if (rank == 0) {
      constexpr uint32_t idesc = PAIR ? make_idesc_mn(128, 2 * N) : make_idesc(N);
      constexpr uint32_t idesc_cat = make_idesc_mn(128, 4 * N);
      mbar_wait(wfull, 0);
      uint32_t stage = 0, phase = 0;
      for (int t = 0; t < T; ++t)
      for (int ch = 0; ch < nch; ++ch) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(ch * ACC_STRIDE);
        mbar_wait(accempty0 + 8 * ch, (t & 1) ^ 1);
        tc_fence_after();
        for (int kc0 = 0; kc0 < KC; ++kc0) {
          const int kc = kc0;
          mbar_wait(hfull0 + 8 * stage, phase);
          tc_fence_after();
          if (elect_one_sync()) {
            if (kc0 == 0) WAVE_DBG(2);
            if (p.dbg && blockIdx.x == 0 && m == 0 && role == 0 && ch == 0 && t >= 300 && t < 304 && kc0 < 8)
              p.dbg[96 + (t - 300) * 8 + kc0] = wgtime();
            const uint32_t sa = smem_ring + stage * 2 * HT;
            const uint64_t a_hi = make_desc_sw128(sa), a_lo = make_desc_sw128(sa + HT);
            const uint64_t b_hi = make_desc_sw128(smem_w + (PAIR ? 2 * kc : kc) * W_TILE);
            const uint64_t b_lo = make_desc_sw128(smem_w + (KC + kc) * W_TILE);          // (not used by PAIR)
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
              if (PAIR) {
                umma_bf16_2sm(d_tmem, a_hi + koff, b_hi + koff, idesc_cat, (kc0 | k) != 0);   // [hi*hi | hi*lo]
                umma_bf16_2sm(d_tmem, a_lo + koff, b_hi + koff, idesc, 1);                      // + lo*hi
              } else {
                umma_bf16(d_tmem, a_lo + koff, b_hi + koff, idesc, (kc0 | k) != 0);
                umma_bf16(d_tmem, a_hi + koff, b_lo + koff, idesc, 1);
                umma_bf16(d_tmem, a_hi + koff, b_hi + koff, idesc, 1);
              }
            }
            if (PAIR) umma_commit_2sm(hempty0 + 8 * stage, 3);
            else umma_commit(hempty0 + 8 * stage);
            if (kc0 + 1 == KC) {
              if (PAIR) umma_commit_2sm(accfull0 + 8 * ch, 3);
              else umma_commit(accfull0 + 8 * ch);
              WAVE_DBG(3);
            }
          }
          __syncwarp();
          if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================================ epilogue: 8 warps, thread = (row, half of the CTA's units) ===========
    const int q = warp & 3;                           // TMEM lane quarter (rows q*32 .. q*32+31)
    const int half = (warp - W_EPI_WARP0) >> 2;       // which part of the HS hidden units
    constexpr int HU = HS / (W_EPI_WARPS / 4);        // units per thread
    constexpr int VW = (HU % 4 == 0) ? 4 : 2;         // vector width of the global accesses (N = 48: 6 units per thread)
    static_assert(HU % VW == 0 && HU <= 8, "units per epilogue thread");
    // TMEM lane -> (row, owner of the gate columns): one CTA per tile: lane = row, own columns; PAIR: see the header
    const int tl = q * 32 + lane;
    const int r = PAIR ? (int)rank * 64 + (tl & 63) : tl;
    const int c_own = PAIR ? (c & ~1) + (tl >> 6) : c;
    const int part = r >> 6;
    const int u0 = c_own * HS + half * HU;            // first hidden unit of this thread
    float cst[2][HU];                                 // cell state of this thread's units, per chunk
#pragma unroll
    for (int j = 0; j < HU; ++j) cst[0][j] = cst[1][j] = 0.f;
    float bias[4 * HU];
    if (role == 1) {
#pragma unroll
      for (int gt = 0; gt < 4; ++gt)
#pragma unroll
        for (int j = 0; j < HU; ++j)
          bias[gt * HU + j] = __ldg(p.bias1 + ((long long)m * NC + c_own) * N + gt * HS + half * HU + j);
    }
    // one time step of chunk CH (compile-time index: the per-chunk cell state stays in registers)
    auto step = [&](const int t, auto CH) {
      constexpr int ch = decltype(CH)::value;
      const bool valid = (r & 63) < p.NBc[ch];
      const int b = p.b0[ch] + (r & 63);              // utterance of the whole batch (row layout, outputs)
      unsigned int* const cB = sync_m + ch * CH_SYNC + W_SYNC_STRIDE;
      unsigned int* const my_ctr = sync_m + ch * CH_SYNC + role * W_SYNC_STRIDE;
      const uint32_t accfull = accfull0 + 8 * ch, accempty = accempty0 + 8 * ch;
      const uint32_t tacc = tmem_base + (uint32_t)(ch * ACC_STRIDE);
      float* const g1x = p.g1x + ch * p.g1x_ch;
      const long long rcur = (long long)b * Tp + 1 + t;
      float gin[4 * HU];
      if (role == 0) {
        if (valid) {
          const float* gp = p.g0 + m * p.g_m_off + part * p.g_p_off + u0 + rcur * p.g_ld;
#pragma unroll
          for (int gt = 0; gt < 4; ++gt)
#pragma unroll
            for (int j = 0; j < HU; j += VW) {
              if (VW == 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(gp + gt * H + j));
                gin[gt * HU + j] = v.x; gin[gt * HU + j + 1] = v.y; gin[gt * HU + j + 2] = v.z; gin[gt * HU + j + 3] = v.w;
              } else {
                const float2 v = __ldg(reinterpret_cast<const float2*>(gp + gt * H + j));
                gin[gt * HU + j] = v.x; gin[gt * HU + j + 1] = v.y;
              }
            }
        } else {
#pragma unroll
          for (int j = 0; j < 4 * HU; ++j) gin[j] = 0.f;
        }
      } else if (role == 2) {
        // G1(t) is produced inside this kernel: acquire counter B, then coherent (L2) loads
        if (lane == 0) wait_counter(cB, (long long)NC * (t + 1), rlx);
        __syncwarp();
        if constexpr (VW == 4) {                      // (the wavefront roles only exist for N = 64)
          const float* gp = g1x + ((((long long)(t & 3) * 2 + m) * W_ROWS + r) * 4) * H + u0;
#pragma unroll
          for (int gt = 0; gt < 4; ++gt)
#pragma unroll
            for (int j = 0; j < HU; j += 4) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(gp + gt * H + j));
              gin[gt * HU + j] = v.x; gin[gt * HU + j + 1] = v.y; gin[gt * HU + j + 2] = v.z; gin[gt * HU + j + 3] = v.w;
            }
        }
      }
      mbar_wait(accfull, t & 1);
      tc_fence_after();
      if (warp == W_EPI_WARP0 && lane == 0) WAVE_DBG(4);
      uint32_t v[4 * HU];                             // [gate][unit]: accumulator columns gate*HS + half*HU + j
#pragma unroll
      for (int gt = 0; gt < 4; ++gt) {
        if (HU == 8) {
          tmem_ld8(tacc + ((uint32_t)(q * 32) << 16) + gt * HS + half * HU, v + gt * HU);
        } else if (HU == 4) {
          tmem_ld4(tacc + ((uint32_t)(q * 32) << 16) + gt * HS + half * HU, v + gt * HU);
        } else {                                      // 8 columns are read, the first HU kept (all inside the allocation)
          uint32_t w8[8];
          tmem_ld8(tacc + ((uint32_t)(q * 32) << 16) + gt * HS + half * HU, w8);
#pragma unroll
          for (int j = 0; j < HU; ++j) v[gt * HU + j] = w8[j];
        }
      }
      if (PAIR) {
        uint32_t v2[4 * HU];
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) {
          if (HU == 8) {
            tmem_ld8(tacc + ((uint32_t)(q * 32) << 16) + N + gt * HS + half * HU, v2 + gt * HU);
          } else if (HU == 4) {
            tmem_ld4(tacc + ((uint32_t)(q * 32) << 16) + N + gt * HS + half * HU, v2 + gt * HU);
          } else {
            uint32_t w8[8];
            tmem_ld8(tacc + ((uint32_t)(q * 32) << 16) + N + gt * HS + half * HU, w8);
#pragma unroll
            for (int j = 0; j < HU; ++j) v2[gt * HU + j] = w8[j];
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4 * HU; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_cluster(accempty & PEER_BIT_MASK); else mbar_arrive(accempty); }
      if (warp == W_EPI_WARP0 && lane == 0) WAVE_DBG(5);
      if (role == 1) {
        // ---- layer-1 input projection: G1(t) rows -> exchange buffer slot t%4
        if constexpr (VW == 4) {
          float* gp = g1x + ((((long long)(t & 3) * 2 + m) * W_ROWS + r) * 4) * H + u0;
#pragma unroll
          for (int gt = 0; gt < 4; ++gt)
#pragma unroll
            for (int j = 0; j < HU; j += 4)
              *reinterpret_cast<float4*>(gp + gt * H + j) =
                  make_float4(__uint_as_float(v[gt * HU + j]) + bias[gt * HU + j],
                              __uint_as_float(v[gt * HU + j + 1]) + bias[gt * HU + j + 1],
                              __uint_as_float(v[gt * HU + j + 2]) + bias[gt * HU + j + 2],
                              __uint_as_float(v[gt * HU + j + 3]) + bias[gt * HU + j + 3]);
        }
        // publish: ONE counter increment per CTA and step (same-address atomics serialise in the L2, ~27 clk each:
        // 8 per CTA x 24 CTAs took 2.4 us to drain): the epilogue warps meet at a named barrier, then one thread
        // fences - cumulative over the stores it has synchronised with, the pattern of a cooperative-groups grid
        // sync - and increments.  (Every thread fencing BEFORE the barrier measured the same step time.)
        asm volatile("bar.sync 1, %0;" ::"n"(32 * W_EPI_WARPS) : "memory");
        if (warp == W_EPI_WARP0 && lane == 0) {
          publish_step(my_ctr, p.sync_mode);
        }
        return;
      }
      float hn[HU];
#pragma unroll
      for (int j = 0; j < HU; ++j) {
        const float ig = wsig(__uint_as_float(v[j]) + gin[j]);
        const float fg = wsig(__uint_as_float(v[HU + j]) + gin[HU + j]);
        const float gg = wtanh(__uint_as_float(v[2 * HU + j]) + gin[2 * HU + j]);
        const float og = wsig(__uint_as_float(v[3 * HU + j]) + gin[3 * HU + j]);
        cst[ch][j] = fg * cst[ch][j] + ig * gg;
        hn[j] = og * wtanh(cst[ch][j]);
      }
      // h(t) goes to the slot the consumers of step t+1 read: (t+1)%4 for h0, (t+1)%2 for h1
      const int wslot = role == 0 ? ((t + 1) & 3) : ((t + 1) & 1);
      if (p.tma_pub) {
        // block of this CTA in shared memory: [hi | lo][its rows][the UW units of its pair] (all rows: the unused ones
        // hold finite values nobody reads), written with one tensor store by the publishing thread below
        const int row_l = PAIR ? (tl & 63) : tl;
        const int uoff = (PAIR ? (tl >> 6) * HS : 0) + half * HU;
        uint8_t* sp = stg + ch * STG + (row_l * UW + uoff) * 2;
        unsigned short hi[HU], lo[HU];
#pragma unroll
        for (int j = 0; j < HU; ++j) split_bf16(hn[j], hi[j], lo[j]);
        if constexpr (HU % 8 == 0) {
#pragma unroll
          for (int j = 0; j < HU; j += 8) {
            *reinterpret_cast<uint4*>(sp + j * 2) = make_uint4((unsigned)hi[j] | ((unsigned)hi[j + 1] << 16), (unsigned)hi[j + 2] | ((unsigned)hi[j + 3] << 16),
                                                              (unsigned)hi[j + 4] | ((unsigned)hi[j + 5] << 16), (unsigned)hi[j + 6] | ((unsigned)hi[j + 7] << 16));
            *reinterpret_cast<uint4*>(sp + ROWS_C * UW * 2 + j * 2) =
                make_uint4((unsigned)lo[j] | ((unsigned)lo[j + 1] << 16), (unsigned)lo[j + 2] | ((unsigned)lo[j + 3] << 16),
                           (unsigned)lo[j + 4] | ((unsigned)lo[j + 5] << 16), (unsigned)lo[j + 6] | ((unsigned)lo[j + 7] << 16));
          }
        } else {
#pragma unroll
          for (int j = 0; j < HU; j += 2) {
            *reinterpret_cast<unsigned*>(sp + j * 2) = (unsigned)hi[j] | ((unsigned)hi[j + 1] << 16);
            *reinterpret_cast<unsigned*>(sp + ROWS_C * UW * 2 + j * 2) = (unsigned)lo[j] | ((unsigned)lo[j + 1] << 16);
          }
        }
        if (p.tma_pub == 1) fence_proxy_async();
      } else if (valid) {
        const int nslot = role == 0 ? 4 : 2;
#pragma unroll
        for (int rep = 0; rep < W_REP; ++rep) {
          unsigned short* hx = (role == 0 ? p.hxA + ch * p.hxA_ch : p.hxC + ch * p.hxC_ch) +
                               (((((long long)rep * nslot + wslot) * 2 + m) * 2) * W_ROWS + r) * H + u0;
#pragma unroll
          for (int j = 0; j < HU; j += VW) {
            if constexpr (VW == 4) st_split4(hx, (long long)W_ROWS * H, j, make_float4(hn[j], hn[j + 1], hn[j + 2], hn[j + 3]));
            else st_split2(hx, (long long)W_ROWS * H, j, make_float2(hn[j], hn[j + 1]));
          }
        }
      }
      if (p.tma_pub == 2) {
        // mode 2: the staged block leaves with 16-byte stores (a row's strip of UW units = UW / 8 vectors per plane): a
        // quarter of the store transactions of the 4-byte stores above
        asm volatile("bar.sync 1, %0;" ::"n"(32 * W_EPI_WARPS) : "memory");
        constexpr int VPR = UW * 2 / 16;                      // vectors per row strip
        const int blk = ((wslot * 2 + m) * 2) + ch * (role == 2 ? p.blkC_ch : p.blkA_ch);
        unsigned short* hx0 = role == 0 ? p.hxA : p.hxC;      // block 0 of the exchange buffer ([block][128 rows][H])
        const int ubase = (PAIR ? (c & ~1) : c) * HS;
        for (int i = threadIdx.x - 32 * W_EPI_WARP0; i < 2 * ROWS_C * VPR; i += 32 * W_EPI_WARPS) {
          const int pl = i / (ROWS_C * VPR), rr = (i / VPR) % ROWS_C, vv = i % VPR;
          const uint4 x = *reinterpret_cast<const uint4*>(stg + ch * STG + ((pl * ROWS_C + rr) * UW) * 2 + vv * 16);
          *reinterpret_cast<uint4*>(hx0 + ((long long)(blk + pl) * W_ROWS + (PAIR ? (int)rank * ROWS_C : 0) + rr) * H + ubase + vv * 8) = x;
        }
      }
      if (warp == W_EPI_WARP0 && lane == 0) WAVE_DBG(6);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * W_EPI_WARPS) : "memory");
      if (warp == W_EPI_WARP0 && lane == 0) {
        if (p.tma_pub == 1) {
          const int blk = ((wslot * 2 + m) * 2) + ch * (role == 2 ? p.blkC_ch : p.blkA_ch);
          tma_store_3d(role == 2 ? &tmSC : &tmSA, smem_u32(stg + ch * STG), (PAIR ? (c & ~1) : c) * HS, PAIR ? (int)rank * ROWS_C : 0, blk);
          bulk_commit_group();
          bulk_wait_all();                 // the block is written (not merely read out of shared memory)
          fence_proxy_async_global();
        }
        if (p.n_roles == 1 && p.kcsync != nullptr) {
          // one increment per K chunk this CTA's stores belong to (the units of its pair: 1 or 2 chunks)
          unsigned int* kb = p.kcsync + (long long)((ch * 2 + m) * KC) * W_SYNC_STRIDE;
          const int ulo = (PAIR ? (c & ~1) : c) * HS, uhi = ulo + (PAIR ? 2 : 1) * HS - 1;
          if (p.sync_mode == 0) {
            __threadfence();
            atomicAdd(kb + (ulo >> 6) * W_SYNC_STRIDE, 1u);
            if ((uhi >> 6) != (ulo >> 6)) atomicAdd(kb + (uhi >> 6) * W_SYNC_STRIDE, 1u);
          } else {
            publish_step(kb + (ulo >> 6) * W_SYNC_STRIDE, p.sync_mode);
            if ((uhi >> 6) != (ulo >> 6)) publish_step(kb + (uhi >> 6) * W_SYNC_STRIDE, p.sync_mode);
          }
        } else {
          publish_step(my_ctr, p.sync_mode);
        }
        WAVE_DBG(7);
      }
      if (role == 2 && valid) {
        if constexpr (VW == 4) {
          const long long oidx = ((long long)(m * 2 + part) * R + rcur) * H + u0;
#pragma unroll
          for (int j = 0; j < HU; j += 4)
            *reinterpret_cast<float4*>(p.hseq1 + oidx + j) = make_float4(hn[j], hn[j + 1], hn[j + 2], hn[j + 3]);
        }
      }
      if (p.n_roles == 1 && valid) {                  // one layer per launch: the sequence outputs, off the critical path
        const long long oidx = ((long long)(m * 2 + part) * R + rcur) * H + u0;
#pragma unroll
        for (int j = 0; j < HU; j += VW) {
          if constexpr (VW == 4) {
            const float4 hv = make_float4(hn[j], hn[j + 1], hn[j + 2], hn[j + 3]);
            if (p.hsplit) st_split4(p.hsplit, 4 * R * H, oidx + j, hv);
            if (p.hseq0) *reinterpret_cast<float4*>(p.hseq0 + oidx + j) = hv;
          } else {
            const float2 hv = make_float2(hn[j], hn[j + 1]);
            if (p.hsplit) st_split2(p.hsplit, 4 * R * H, oidx + j, hv);
            if (p.hseq0) *reinterpret_cast<float2*>(p.hseq0 + oidx + j) = hv;
          }
        }
      }
    };
    for (int t = 0; t < T; ++t) {
      step(t, std::integral_constant<int, 0>{});
      if (nch > 1) step(t, std::integral_constant<int, 1>{});
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == W_EPI_WARP0) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int N, bool PAIR>
static int launch_wave(const CUtensorMap* maps, const WaveParams& p, size_t smem, cudaStream_t st) {
  IDV_CUDA(cudaFuncSetAttribute(lstm_wave_tc_kernel<N, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(p.NC, 2 * p.n_roles, 1), block(wave_threads(N));
  cudaError_t e;
  if (PAIR) {
    // clusters of 2 along x.  All 6 * NC CTAs wait on one another and must be co-resident; that is checked against
    // cudaOccupancyMaxActiveClusters (the caller falls back to the one-CTA-per-tile cooperative kernel otherwise)
    // instead of asked for with the cooperative attribute: Nsight Compute cannot replay a launch that carries both the
    // cooperative and the cluster attribute (LaunchFailed), and every CTA that has not started yet only ever waits for
    // SMs held by OTHER kernels.  Every spin in the kernel has a timeout that traps.
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    IDV_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, lstm_wave_tc_kernel<N, PAIR>, &cfg));
    if (2 * max_clusters < (int)(grid.x * grid.y)) return IDV_E_RESOURCE;
    e = cudaLaunchKernelEx(&cfg, lstm_wave_tc_kernel<N, PAIR>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], p);
  } else {
    void* args[] = {(void*)&maps[0], (void*)&maps[1], (void*)&maps[2], (void*)&maps[3], (void*)&maps[4], (void*)&maps[5],
                    (void*)&maps[6], (void*)&p};
    e = cudaLaunchCooperativeKernel((const void*)lstm_wave_tc_kernel<N, PAIR>, grid, block, args, smem, st);
  }
  if (e == cudaErrorCooperativeLaunchTooLarge) {
    set_error("idv_lstm2_wave_tc: cooperative grid %dx6 is not co-resident", p.NC);
    return IDV_E_RESOURCE;
  }
  if (e != cudaSuccess) {
    set_error("idv_lstm2_wave_tc: launch failed: %s", cudaGetErrorString(e));
    return IDV_E_CUDA;
  }
  return IDV_OK;
}

// utterances [b0, b0 + per_launch) of the batch as one or two chunks of <= 64
static void set_chunks(WaveParams& p, int b0, int NB, int per_launch) {
  const int n = NB - b0 < per_launch ? NB - b0 : per_launch;
  p.nch = n > 64 ? 2 : 1;
  p.b0[0] = b0; p.NBc[0] = n < 64 ? n : 64;
  p.b0[1] = b0 + 64; p.NBc[1] = n > 64 ? n - 64 : 0;
}

static int wave_cols(int H) {
  // gate columns per CTA such that 6 * (H / Hs) CTAs fit the device (148 SMs): N = 64 for H = 384, 128
  if (H % 64 != 0) return 0;
  if (H % 16 == 0 && 6 * (H / 16) <= 148) return 64;
  return 0;
}

}  // namespace tc
}  // namespace idv

extern "C" int idv_lstm2_wave_config(int H, int* n_cols, int* n_ctas, int64_t* work_bytes) {
  using namespace idv;
  IDV_CHECK_ARG(n_cols && n_ctas && work_bytes, "idv_lstm2_wave_config: null pointer");
  const int N = tc::wave_cols(H);
  IDV_CHECK_ARG(N > 0, "idv_lstm2_wave_config: hidden size %d is not supported by the wavefront kernel", H);
  *n_cols = N;
  *n_ctas = H / (N / 4);
  // per chunk: hxA (4 slots) + hxC (2 slots) bf16 [slot][2][2][128][H]  +  g1x fp32 [4][2][128][4H]; two chunks per launch
  *work_bytes = 2 * ((int64_t)tc::W_REP * (4 + 2) * 2 * 2 * 128 * H * 2 + (int64_t)4 * 2 * 128 * 4 * H * 4);
  return IDV_OK;
}

static int wave_run(bool pair, const float* g0, int64_t g_m_off, int64_t g_p_off, int g_ld, const void* w_hh0,
                    const void* w_ih1, const void* w_hh1, const float* bias1, int NB, int T, int H, float* hseq1,
                    void* work, unsigned int* sync, int t_valid, void* stream) {
  using namespace idv;
  using namespace idv::tc;
  IDV_CHECK_ARG(g0 && w_hh0 && w_ih1 && w_hh1 && bias1 && hseq1 && work && sync, "idv_lstm2_wave_tc: null pointer");
  IDV_CHECK_ARG(NB > 0 && T > 0, "idv_lstm2_wave_tc: empty problem (NB %d, T %d)", NB, T);
  int N = 0, NC = 0;
  int64_t work_bytes = 0;
  int rc = idv_lstm2_wave_config(H, &N, &NC, &work_bytes);
  if (rc) return rc;
  const int KC = H / 64;
  int dev = 0, sms = 0, smem_optin = 0;
  IDV_CUDA(cudaGetDevice(&dev));
  IDV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  IDV_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  IDV_CHECK_ARG(6 * NC <= sms, "idv_lstm2_wave_tc: %d CTAs exceed the %d SMs", 6 * NC, sms);
  const size_t w_bytes = (size_t)2 * KC * N * BK * 2;
  pair = pair && NC % 2 == 0;
  const size_t stage_bytes = (size_t)2 * W_HTILE / (pair ? 2 : 1);
  // tensor-store publish: staging of two chunks' h blocks [hi | lo][rows of the CTA][units of the CTA / pair]
  const int uw = (pair ? 2 : 1) * (N / 4), rows_c = pair ? W_ROWS / 2 : W_ROWS;
  // how h(t) leaves the CTA (option "lstm_tma_publish"; auto = direct stores here: the staged forms measured slower at N = 64)
  const int tma_pub = (uw * 2) % 16 == 0 && option_lstm_tma_publish() > 0 ? option_lstm_tma_publish() : 0;
  const size_t stg_bytes = tma_pub ? (size_t)2 * 2 * rows_c * uw * 2 : 0;
  int stages = (int)(((size_t)smem_optin - w_bytes - stg_bytes - 1024 - 256) / stage_bytes);
  if (stages > (pair ? 7 : 8)) stages = pair ? 7 : 8;
  if (stages > KC) stages = KC;
  IDV_CHECK_ARG(stages >= 1, "idv_lstm2_wave_tc: not enough shared memory for H=%d", H);
  const size_t smem = w_bytes + (size_t)stages * stage_bytes + stg_bytes + 1024 + 256;
  uint8_t* wk = reinterpret_cast<uint8_t*>(work);
  // workspace: [hxA chunk 0 | hxA chunk 1][hxC chunk 0 | hxC chunk 1][g1x chunk 0 | g1x chunk 1]
  const size_t hxA_bytes = (size_t)W_REP * 4 * 2 * 2 * 128 * H * 2, hxC_bytes = (size_t)W_REP * 2 * 2 * 2 * 128 * H * 2;
  const size_t g1x_bytes = (size_t)4 * 2 * 128 * 4 * H * 4;
  CUtensorMap maps[7];
  const void* wp[3] = {w_hh0, w_ih1, w_hh1};
  for (int i = 0; i < 3; ++i) {
    rc = encode_map_2d(&maps[i], wp[i], H, (uint64_t)2 * 2 * NC * N, BK, N);
    if (rc) return rc;
  }
  rc = encode_map_3d(&maps[3], wk, H, W_ROWS, (uint64_t)2 * W_REP * 4 * 2 * 2, BK, pair ? W_ROWS / 2 : W_ROWS, 2);
  if (rc) return rc;
  rc = encode_map_3d(&maps[4], wk + 2 * hxA_bytes, H, W_ROWS, (uint64_t)2 * W_REP * 2 * 2 * 2, BK, pair ? W_ROWS / 2 : W_ROWS, 2);
  if (rc) return rc;
  maps[5] = maps[3]; maps[6] = maps[4];
  if (tma_pub == 1) {   // the same buffers for the tensor stores: boxes {units of the CTA / pair, its rows, hi | lo}, no swizzle
    rc = encode_map_3d(&maps[5], wk, H, W_ROWS, (uint64_t)2 * W_REP * 4 * 2 * 2, uw, rows_c, 2, false);
    if (rc) return rc;
    rc = encode_map_3d(&maps[6], wk + 2 * hxA_bytes, H, W_ROWS, (uint64_t)2 * W_REP * 2 * 2 * 2, uw, rows_c, 2, false);
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  WaveParams p;
  p.tma_pub = tma_pub;
  p.g0 = g0; p.g_m_off = g_m_off; p.g_p_off = g_p_off; p.g_ld = g_ld; p.bias1 = bias1;
  p.NB = NB; p.T = T; p.H = H; p.NC = NC; p.KC = KC; p.stages = stages;
  p.Tsteps = (t_valid > 0 && t_valid < T) ? t_valid : T;
  p.hseq1 = hseq1;
  p.hxA = reinterpret_cast<unsigned short*>(wk);
  p.hxC = reinterpret_cast<unsigned short*>(wk + 2 * hxA_bytes);
  p.g1x = reinterpret_cast<float*>(wk + 2 * hxA_bytes + 2 * hxC_bytes);
  p.hxA_ch = (long long)(hxA_bytes / 2); p.hxC_ch = (long long)(hxC_bytes / 2); p.g1x_ch = (long long)(g1x_bytes / 4);
  p.blkA_ch = W_REP * 4 * 2 * 2; p.blkC_ch = W_REP * 2 * 2 * 2;
  p.sync = sync; p.kcsync = nullptr;
  p.n_roles = 3; p.hsplit = nullptr; p.hseq0 = nullptr;
  p.sync_mode = option_lstm_sync_mode();
  p.dbg = nullptr;
  const bool dbg = getenv("IDV_LSTM_DBG") != nullptr && p.Tsteps > 304;
  if (dbg) {
    IDV_CUDA(cudaMalloc(&p.dbg, (96 + 32) * sizeof(unsigned long long)));
    IDV_CUDA(cudaMemsetAsync(p.dbg, 0, (96 + 32) * sizeof(unsigned long long), st));
  }
  // a chunk holds 64 utterances (M = 128 rows = 2 input parts x 64); a launch interleaves up to two chunks (128
  // utterances); larger batches run as consecutive launches on the same stream, each with freshly zeroed exchange
  // buffers / counters (h(-1) = 0)
  const int per_launch = option_lstm_interleave() ? 128 : 64;
  for (int b0 = 0; b0 < NB && rc == IDV_OK; b0 += per_launch) {
    set_chunks(p, b0, NB, per_launch);
    IDV_CUDA(cudaMemsetAsync(work, 0, (size_t)work_bytes, st));
    IDV_CUDA(cudaMemsetAsync(sync, 0, 2 * 6 * W_SYNC_STRIDE * sizeof(unsigned int), st));
    rc = pair ? launch_wave<64, true>(maps, p, smem, st) : launch_wave<64, false>(maps, p, smem, st);
  }
  if (dbg && rc == IDV_OK) {
    unsigned long long h[96 + 32];
    IDV_CUDA(cudaStreamSynchronize(st));
    IDV_CUDA(cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    const char* names[3] = {"L0", "IP", "L1"};
    for (int ro = 0; ro < 3; ++ro)
      for (int i = 0; i < 4; ++i) {
        fprintf(stderr, "[wave dbg] %s t=%d:", names[ro], 300 + i);
        for (int sl = 0; sl < 8; ++sl) fprintf(stderr, " %lld", (long long)(h[(ro * 4 + i) * 8 + sl] - h[0]));
        fprintf(stderr, "\n");
      }
    for (int i = 0; i < 4; ++i) {
      fprintf(stderr, "[wave dbg] L0 tile arrivals t=%d:", 300 + i);
      for (int k = 0; k < 6; ++k) fprintf(stderr, " %lld", (long long)(h[96 + i * 8 + k] - h[0]));
      fprintf(stderr, "\n");
    }
  }
  return rc;
}

extern "C" int idv_lstm2_wave_tc(const float* g0, int64_t g_m_off, int64_t g_p_off, int g_ld, const void* w_hh0,
                                 const void* w_ih1, const void* w_hh1, const float* bias1, int NB, int T, int H,
                                 float* hseq1, void* work, unsigned int* sync, int t_valid, void* stream) {
  int rc = IDV_E_RESOURCE;
  // pairs are launched without the cooperative guarantee (see launch_wave): not when kernels of several streams share
  // the GPU ("gemm_dynamic_tiles" is the multi-stream switch) - two half-resident grids could wait for each other's SMs
  if (idv::option_lstm_wave_pairs() && !idv::option_dynamic_tiles())
    rc = wave_run(true, g0, g_m_off, g_p_off, g_ld, w_hh0, w_ih1, w_hh1, bias1, NB, T, H, hseq1, work, sync, t_valid, stream);
  if (rc == IDV_E_RESOURCE)        // the CTA pairs do not all fit the device at once: one CTA per tile, cooperative launch
    rc = wave_run(false, g0, g_m_off, g_p_off, g_ld, w_hh0, w_ih1, w_hh1, bias1, NB, T, H, hseq1, work, sync, t_valid, stream);
  return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// ONE nn.LSTM layer of both modules per launch on the same kernel (role L0 only): the CTA-pair recurrence for hidden
// sizes whose two layers do not fit the device at once (H = 768: the three weight matrices of the wavefront are 56.6 MB
// of bf16 hi/lo against 33.6 MB of shared memory on 148 SMs - at most one layer's W_hh, 18.9 MB, can be resident) and
// for the training forward, which needs the per-layer sequences.  Against idv_lstm_recurrent_tc (one CTA per tile): a
// CTA streams 64 instead of 128 rows of h per step, two MMAs per K step, one counter increment per CTA and step.
// ---------------------------------------------------------------------------------------------------------------------
static int layer_cols(int H) {
  if (H % 64 != 0) return 0;
  if (H <= 512 && H % 16 == 0 && (H / 16) % 2 == 0) return 64;
  if (H <= 768 && H % 12 == 0 && (H / 12) % 2 == 0) return 48;
  return 0;
}

extern "C" int idv_lstm_layer_pair_config(int H, int* n_cols, int* n_ctas, int64_t* work_bytes) {
  using namespace idv;
  IDV_CHECK_ARG(n_cols && n_ctas && work_bytes, "idv_lstm_layer_pair_config: null pointer");
  const int N = layer_cols(H);
  IDV_CHECK_ARG(N > 0, "idv_lstm_layer_pair_config: hidden size %d is not supported by the CTA-pair recurrence", H);
  *n_cols = N;
  *n_ctas = H / (N / 4);
  // h exchange buffers: bf16 [2 chunks][4 slots][2 m][2 hl][128][H], then the per-K-chunk step counters [2][2][H/64] lines
  *work_bytes = 2 * (int64_t)4 * 2 * 2 * 128 * H * 2 + (int64_t)2 * 2 * (H / 64) * tc::W_SYNC_STRIDE * 4;
  return IDV_OK;
}

extern "C" int idv_lstm_layer_pair_tc(const float* g, int64_t g_m_off, int64_t g_p_off, int g_ld, const void* wpack, int NB,
                                      int T, int H, float* hseq, void* hsplit, void* work, unsigned int* sync, int t_valid,
                                      void* stream) {
  using namespace idv;
  using namespace idv::tc;
  IDV_CHECK_ARG(g && wpack && work && sync && (hseq || hsplit), "idv_lstm_layer_pair_tc: null pointer");
  IDV_CHECK_ARG(NB > 0 && T > 0, "idv_lstm_layer_pair_tc: empty problem");
  int N = 0, NC = 0;
  int64_t work_bytes = 0;
  int rc = idv_lstm_layer_pair_config(H, &N, &NC, &work_bytes);
  if (rc) return rc;
  const int KC = H / 64;
  int dev = 0, sms = 0, smem_optin = 0;
  IDV_CUDA(cudaGetDevice(&dev));
  IDV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  IDV_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  IDV_CHECK_ARG(2 * NC <= sms, "idv_lstm_layer_pair_tc: %d CTAs exceed the %d SMs", 2 * NC, sms);
  const size_t w_bytes = (size_t)2 * KC * N * BK * 2;
  // (no cooperative guarantee for clusters: not when kernels of several streams share the GPU, see idv_lstm2_wave_tc)
  bool pair = option_lstm_wave_pairs() && !option_dynamic_tiles();
  cudaStream_t st = (cudaStream_t)stream;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const size_t stage_bytes = (size_t)2 * W_HTILE / (pair ? 2 : 1);
    const int uw = (pair ? 2 : 1) * (N / 4), rows_c = pair ? W_ROWS / 2 : W_ROWS;
    // auto: staged 16-byte stores for the 48-column kernel (H = 768: 4-byte stores of 6 units per thread otherwise), direct
    // stores for the 64-column one (16-byte stores already)
    const int tma_opt = option_lstm_tma_publish() < 0 ? (N == 48 ? 2 : 0) : option_lstm_tma_publish();
    const int tma_pub = (uw * 2) % 16 == 0 ? tma_opt : 0;
    const size_t stg_bytes = tma_pub ? (size_t)2 * 2 * rows_c * uw * 2 : 0;
    int stages = (int)(((size_t)smem_optin - w_bytes - stg_bytes - 1024 - 256) / stage_bytes);
    if (stages > (pair ? 7 : 8)) stages = pair ? 7 : 8;
    if (stages > KC) stages = KC;
    IDV_CHECK_ARG(stages >= 2 || (stages >= 1 && KC == 1), "idv_lstm_layer_pair_tc: not enough shared memory for H=%d", H);
    const size_t smem = w_bytes + (size_t)stages * stage_bytes + stg_bytes + 1024 + 256;
    CUtensorMap maps[7];
    rc = encode_map_2d(&maps[0], wpack, H, (uint64_t)2 * 2 * NC * N, BK, N);
    if (rc) return rc;
    rc = encode_map_3d(&maps[3], work, H, W_ROWS, (uint64_t)2 * 4 * 2 * 2, BK, pair ? W_ROWS / 2 : W_ROWS, 2);
    if (rc) return rc;
    maps[1] = maps[0]; maps[2] = maps[0]; maps[4] = maps[3];
    maps[5] = maps[3];
    if (tma_pub == 1) {
      rc = encode_map_3d(&maps[5], work, H, W_ROWS, (uint64_t)2 * 4 * 2 * 2, uw, rows_c, 2, false);
      if (rc) return rc;
    }
    maps[6] = maps[5];
    WaveParams p;
    p.tma_pub = tma_pub;
    p.g0 = g; p.g_m_off = g_m_off; p.g_p_off = g_p_off; p.g_ld = g_ld; p.bias1 = nullptr;
    p.NB = NB; p.T = T; p.H = H; p.NC = NC; p.KC = KC; p.stages = stages;
    p.Tsteps = (t_valid > 0 && t_valid < T) ? t_valid : T;
    p.hseq1 = nullptr;
    p.hxA = reinterpret_cast<unsigned short*>(work); p.hxC = nullptr; p.g1x = nullptr;
    p.hxA_ch = (long long)4 * 2 * 2 * 128 * H; p.hxC_ch = 0; p.g1x_ch = 0;
    p.blkA_ch = 4 * 2 * 2; p.blkC_ch = 0;
    p.sync = sync; p.dbg = nullptr;
    p.kcsync = option_lstm_chunk_sync()
                   ? reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(work) + (size_t)2 * 4 * 2 * 2 * 128 * H * 2)
                   : nullptr;
    p.n_roles = 1; p.hsplit = reinterpret_cast<unsigned short*>(hsplit); p.hseq0 = hseq;
    p.sync_mode = option_lstm_sync_mode();
    const bool dbg = getenv("IDV_LSTM_DBG") != nullptr && p.Tsteps > 304;
    if (dbg) {
      IDV_CUDA(cudaMalloc(&p.dbg, (96 + 32) * sizeof(unsigned long long)));
      IDV_CUDA(cudaMemsetAsync(p.dbg, 0, (96 + 32) * sizeof(unsigned long long), st));
    }
    rc = IDV_OK;
    const int per_launch = option_lstm_interleave() ? 128 : 64;
    for (int b0 = 0; b0 < NB && rc == IDV_OK; b0 += per_launch) {
      set_chunks(p, b0, NB, per_launch);
      IDV_CUDA(cudaMemsetAsync(work, 0, (size_t)work_bytes, st));
      IDV_CUDA(cudaMemsetAsync(sync, 0, 2 * 6 * W_SYNC_STRIDE * sizeof(unsigned int), st));
      if (N == 64) rc = pair ? launch_wave<64, true>(maps, p, smem, st) : launch_wave<64, false>(maps, p, smem, st);
      else rc = pair ? launch_wave<48, true>(maps, p, smem, st) : launch_wave<48, false>(maps, p, smem, st);
      if (rc == IDV_E_RESOURCE && b0 == 0) break;
    }
    if (dbg) {
      if (rc == IDV_OK) {
        unsigned long long h[96 + 32];
        IDV_CUDA(cudaStreamSynchronize(st));
        IDV_CUDA(cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost));
        // slots: 0 deps satisfied, 1 loads issued, 2 first tile landed, 3 MMAs issued, 4 accumulator ready, 5 TMEM drained,
        // 6 stores done, 7 published
        for (int i = 0; i < 4; ++i) {
          fprintf(stderr, "[layer dbg] H=%d N=%d t=%d:", H, N, 300 + i);
          for (int sl = 0; sl < 8; ++sl) fprintf(stderr, " %lld", (long long)(h[i * 8 + sl] - h[0]));
          fprintf(stderr, "\n");
        }
        for (int i = 0; i < 2; ++i) {
          fprintf(stderr, "[layer dbg] tile arrivals t=%d:", 300 + i);
          for (int k = 0; k < 8; ++k) fprintf(stderr, " %lld", (long long)(h[96 + i * 8 + k] - h[0]));
          fprintf(stderr, "\n");
        }
      }
      cudaFree(p.dbg);
      p.dbg = nullptr;
    }
    if (rc != IDV_E_RESOURCE || !pair) break;
    pair = false;                 // the CTA pairs are not all co-resident: one CTA per tile, cooperative launch
  }
  return rc;
}
