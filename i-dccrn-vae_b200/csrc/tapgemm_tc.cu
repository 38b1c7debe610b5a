// tap-GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), error-compensated bf16:
//
//   every fp32 operand x is held as two bf16 planes  hi = bf16(x), lo = bf16(x - hi)  (same bytes as
//   fp32) and every product is issued as three MMAs  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  accumulated
//   in fp32 in TMEM (single-pass TF32/BF16 misses the 1e-4 parity budget, this split
//   measures 1.6e-5 end to end).
//
// Implicit GEMM without im2col: activations are planes [hl][F][R][Cp] (bf16), so the A tile of tap
// (f_in, dt, ch_off, k0) for output rows [r0, r0+128) is the box {64 ch, 128 rows, 1 plane, 2 (hi,lo)} at
// (ch_off+k0, r0-dt, f_in, 0) of a 4-D tensor map; rows outside [0,R) (the r = -1 causal tap of the first
// tile, the tail of the last) are zero-filled by TMA.  Weights are [hl][slot][N][kc_max] (K-major), one
// box {64, BN, 1, 2} per K step.  Both land 128B-swizzled, which is the UMMA K-major SW128 canonical
// layout, so the MMA descriptors need no data movement.
//
// Persistent, warp-specialised CTA (one per SM):  warp 0 = TMA producer, warp 1 = MMA issuer (one
// elected thread), warps 2-5 = epilogue (TMEM -> registers -> bias + PReLU + pad-row mask -> split bf16
// or fp32 -> global).  STAGES-deep smem ring (full/empty mbarriers) and two TMEM accumulator stages
// (tmem_full/tmem_empty) so the epilogue of tile i overlaps the main loop of tile i+1.
#include "tc_common.cuh"

namespace idv {
namespace tc {

constexpr int EPI_WARP0 = 2;
// Epilogue warps per TMEM lane quarter.  Measured on B200 with 8 warps (two per quarter, each taking every other
// 64-column slab) for BN <= 128: the narrow layers got SLOWER (enc1 1.07 -> 1.29 ms, dec4 1.21 -> 1.34 ms) because the
// doubled staging buffers cost a ring stage - those layers are bound by operand ingest (L2 -> shared memory: 80 KB per
// K chunk and CTA pair for 0.33 us of MMAs), not by their epilogue; the kernel keeps the generality, the count stays 4.
template <int BN> struct EpiWarps { static constexpr int value = 4; };
constexpr int MAX_THREADS = 64 + 32 * 8;

struct Params {
  int R, Tp, N, t_valid, keep_pad;
  int n_units, n_row_tiles, n_col_tiles;
  const idv_unit_t* units;
  const idv_tap_t* taps;
  const float* bias;
  const float* bias2;         // optional: bias of the FIRST frame of every utterance (rows with r % Tp == 1) - a layer
                              // composed with the affine map in front of it (dense + first decoder layer) loses the
                              // bias terms of the time taps that read the zero pad row
  void* out;
  int out_ld;
  long long out_plane;        // elements between output planes
  long long out_hl;           // elements between the hi and the lo plane set (split mode)
  int out_split;              // 1: bf16 hi/lo planes, 0: fp32
  int apply_prelu;
  float slope;
  // head mode (last decoder layer, Cout = 1): a unit holds unit.out_ch_off (<= 16) consecutive output bins starting at
  // fo = unit.out_f, bin e in accumulator columns (2e, 2e+1) = (re, im); the epilogue applies bias + PReLU (+ mask
  // head) and writes `predict` (NBtot, head_fout, T, 2) directly.  0 = off, 1 = real/imag, 2 = mask.
  int head;
  int head_fout, head_bmul, head_boff;
  const float* stft_x;
  float* predict;
  unsigned int* sched;        // [0] next tile, [1] CTAs finished (dynamic tile scheduler; self-resetting)
  int order;                  // tile order: 0 = (unit, row tile, N tile), 1 = (row tile, unit, N tile) - see decode_tile
  int tma_out;                // 1: split-bf16 tiles leave through shared memory + TMA stores (tmO), see the epilogue
  // split-K (small problems: a frame-streaming step has 10-40 output tiles for 148 SMs, and ONE CTA streaming the whole K
  // range of a tile - megabytes of weights and activation rows - is bound by the L2 -> shared-memory rate of a single
  // SM): a tile's 64-wide K steps are divided over `ksplit` CTAs; every CTA stores its partial accumulator in the fp32
  // workspace, waits until all parts of its tile are there and then reduces + finishes every ksplit-th 32-column slab
  // (fixed summation order: deterministic).  Tile index = (output tile) * ksplit + part; ksplit <= BN / 32.
  // (First version, measured: partial sums added with red.global.add.v4.f32 and the last CTA finishing the tile - the
  // 1.1 M vector reductions of a layer took 100-150 us, four times the unsplit kernel.)
  int ksplit;
  float* ws;                  // [output tiles][ksplit][128 rows][BN] partial accumulators
  unsigned int* ws_cnt;       // [output tiles][2] parts stored / parts finished, zero between launches
};

// Tile index -> (unit, row tile, N tile).  The CTAs that run at the same time work on CONSECUTIVE tile indices, so the
// order decides what they share in the L2:
//   order 0: unit outermost - co-resident CTAs walk the rows of ONE output plane; the input planes it reads (5 for a
//            stride-2 conv, 84 MB each at batch 64) are evicted from the 126 MB L2 before the next output plane, which
//            shares 3 of them, comes round: every input plane is read from DRAM 2-3 times (ncu: 39.8 GB per step for
//            23 GB of compulsory traffic);
//   order 1: row tile outermost, unit inner - co-resident CTAs compute ALL output planes of the same few row tiles,
//            so an input box is fetched from DRAM once and served from the L2 to the 2-5 units that read it.
__device__ __forceinline__ void decode_tile(const Params& p, int t, int n_row_tiles, int& u, int& rt, int& nt) {
  nt = t % p.n_col_tiles;
  const int q = t / p.n_col_tiles;
  if (p.order) {
    u = q % p.n_units;
    rt = q / p.n_units;
  } else {
    rt = q % n_row_tiles;
    u = q / n_row_tiles;
  }
}

// TWO: the kernel runs as CTA PAIRS (thread-block clusters of 2, tcgen05 cta_group::2).  A pair computes a 256 x BN
// tile: each CTA stages its own 128 activation rows and HALF of the weight tile (BN/2 rows), one thread of the even
// CTA issues M = 256 MMAs that read both CTAs' shared memory and write both CTAs' TMEM.  Per MMA a CTA's shared
// memory serves 4 KB (A) + BN/2 x 32 B (B) instead of 4 KB + BN x 32 B, and TMA writes half the weight bytes: the
// shared-memory traffic (TMA writes + UMMA reads) that bounds the 1-CTA kernel drops below the math time.
template <int BN, bool TWO = false>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;              // one of hi / lo
  static constexpr int B_BYTES = (TWO ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int EPI_WARPS = EpiWarps<BN>::value;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  // epilogue staging for the TMA stores: per epilogue warp [hi | lo][32 rows][64 bf16] = 8 KB, 1024-byte aligned
  static constexpr int OUT_STAGE_BYTES = BN >= 64 ? EPI_WARPS * 8192 : 0;      // (N = 32 tiles never leave through TMA)
  // ring depth: as many stages as fit next to the staging buffers (227 KB per SM)
  static constexpr int STAGES = TWO ? (BN >= 256 ? 3 : (BN >= 64 ? 4 : 5)) : (BN >= 256 ? 2 : (BN >= 128 ? 3 : 4));
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;   // two accumulator stages, power of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 1024 /*barriers, padded*/ + OUT_STAGE_BYTES;
  static_assert(SMEM_BYTES <= 232448, "stage ring + output staging exceed the shared memory of one SM");
};

// DYN = false: tile i of CTA b is b + i*gridDim.x (lock-step CTAs, best when the kernel owns the GPU);
// DYN = true : tiles are claimed from a global counter (kernels of other streams may hold SMs).
template <int BN, bool DYN, bool TWO>
__global__ void __launch_bounds__(MAX_THREADS, 1)
tapgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO, const Params p) {
  static_assert(!(TWO && DYN), "the CTA-pair kernel uses the static tile schedule");
  using C = Cfg<BN, TWO>;
  // TWO: rank of this CTA in its pair, index / count of pairs; a "tile" is then a PAIR of row tiles (2*rt + rank)
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;
  const int wid = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nwork = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tile-queue full[4] / empty[4], then the TMEM
  // base address and the 4-entry tile queue
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * C::STAGES;
  const uint32_t tfull0 = empty0 + 8 * C::STAGES, tempty0 = tfull0 + 16;
  const uint32_t qfull0 = tempty0 + 16, qempty0 = qfull0 + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 12);
  volatile int* tq = reinterpret_cast<volatile int*>(tmem_slot + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t out_stage0 = smem_base + C::STAGES * C::STAGE_BYTES + 1024;      // (stage ring is a multiple of 1024 B)
  // Layers with a short K loop (the first encoder layer on the STFT rows: 2 chunks; dense + first decoder layer: 8) are
  // bound by their epilogue, not by the main loop: they give up one ring stage, and the freed shared memory holds extra
  // output staging buffers so that several tensor stores per epilogue warp are in flight instead of one.
  const int nst = (p.tma_out && C::STAGES > 2 && p.units[0].reserved <= 8) ? C::STAGES - 1 : C::STAGES;
  const int n_obufs = C::OUT_STAGE_BYTES ? 1 + ((C::STAGES - nst) * C::STAGE_BYTES) / C::OUT_STAGE_BYTES : 1;   // 1, 2 or 3
  const uint32_t out_stage_x = smem_base + nst * C::STAGE_BYTES;                 // extra staging buffers (freed stage)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_row_tiles = TWO ? (p.n_row_tiles + 1) / 2 : p.n_row_tiles;
  const int total_tiles = p.n_units * n_row_tiles * p.n_col_tiles * p.ksplit;

  pdl_trigger();          // the next kernel of the stream may be scheduled; it waits for this grid at its own pdl_wait()
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmW);
    if (p.tma_out) prefetch_tmap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, (TWO ? 2 : 1) * C::EPI_WARPS);      // one arrive per epilogue warp (of both CTAs of a pair)
    }
    for (int qi = 0; qi < 4; ++qi) {
      mbar_init(qfull0 + 8 * qi, 1);
      mbar_init(qempty0 + 8 * qi, 1 + C::EPI_WARPS);     // MMA thread + the epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == EPI_WARP0) {
    if (TWO) {
      tmem_alloc_2sm(smem_u32(tmem_slot), C::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();            // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the predecessor's tail; its outputs (this kernel's
  // activation planes) may only be read, and this kernel's outputs written, from here on
  pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      // dynamic tile scheduler: tiles are claimed from a global counter so that any number of co-resident CTAs
      // (other kernels may hold SMs) share the work evenly; the claimed ids are handed to the MMA / epilogue
      // warps through a 4-entry shared-memory queue (-1 terminates)
      int t_next = DYN ? (int)atomicAdd(p.sched, 1u) : wid;
      for (uint32_t qi = 0;; ++qi) {
        int t = t_next;
        if (t >= total_tiles) t = -1;
        if (DYN) {
          const uint32_t qs = qi & 3, qph = (qi >> 2) & 1;
          mbar_wait(qempty0 + 8 * qs, qph ^ 1);
          tq[qs] = t;
          mbar_arrive(qfull0 + 8 * qs);
        }
        if (t < 0) break;
        // DYN: claimed one tile ahead, the atomic's latency hides behind this tile's loads
        t_next = DYN ? (int)atomicAdd(p.sched, 1u) : t + nwork;
        int ui, rt, nt;
        decode_tile(p, t / p.ksplit, n_row_tiles, ui, rt, nt);
        if (TWO) rt = 2 * rt + (int)rank;
        const idv_unit_t unit = p.units[ui];
        // split-K: this CTA's part of the unit's K steps (counted over the taps in table order)
        const int part = t % p.ksplit;
        const int c_begin = unit.reserved * part / p.ksplit, c_end = unit.reserved * (part + 1) / p.ksplit;
        int ci = 0;
        // co-resident CTAs work on the same unit: start each at a different tap so they do not all request the
        // same weight tile (same L2 lines) at the same moment; the accumulation order is per-CTA but fixed
        // (both CTAs of a pair walk the taps in the same order)
        const int rot = p.ksplit > 1 ? 0 : (int)((unsigned)wid % (unsigned)unit.n_taps);
        for (int ti0 = 0; ti0 < unit.n_taps; ++ti0) {
          const int ti = ti0 + rot < unit.n_taps ? ti0 + rot : ti0 + rot - unit.n_taps;
          const idv_tap_t tap = p.taps[unit.tap_begin + ti];
          const CUtensorMap* am = tap.src ? &tmA1 : &tmA0;
          for (int k0 = 0; k0 < tap.kc; k0 += BK, ++ci) {
            if (ci < c_begin || ci >= c_end) continue;
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
            if (TWO) {
              // the full barrier of a stage lives in the even CTA and collects the bytes of BOTH CTAs' loads
              const uint32_t fb = (full0 + 8 * stage) & PEER_BIT_MASK;
              if (rank == 0) mbar_expect_tx(full0 + 8 * stage, 2 * C::STAGE_BYTES);
              tma_load_4d_2sm(am, fb, sa, tap.ch_off + k0, rt * BM - tap.dt, tap.f_in, 0);
              tma_load_4d_2sm(&tmW, fb, sa + 2 * C::A_BYTES, k0, nt * BN + (int)rank * (BN / 2), tap.w_off, 0);
            } else {
              mbar_expect_tx(full0 + 8 * stage, C::STAGE_BYTES);
              tma_load_4d(am, full0 + 8 * stage, sa, tap.ch_off + k0, rt * BM - tap.dt, tap.f_in, 0);
              tma_load_4d(&tmW, full0 + 8 * stage, sa + 2 * C::A_BYTES, k0, nt * BN, tap.w_off, 0);
            }
            if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The whole warp runs the loop and the tcgen05 instructions are issued under elect.sync.  (Issued from a
    // `lane == 0` branch ptxas cannot know that one thread is active and wraps EVERY tcgen05.mma in an elect / R2UR /
    // branch loop - ~50 clocks of issue per MMA, more than a 128 x 64 MMA takes: the narrow tiles were issue-bound.)
    if (rank == 0) {
      constexpr uint32_t idesc = TWO ? make_idesc_m256(BN) : make_idesc(BN);
      uint32_t stage = 0, phase = 0;
      for (uint32_t local = 0;; ++local) {
        int t;
        if (DYN) {
          const uint32_t qs = local & 3, qph = (local >> 2) & 1;
          mbar_wait(qfull0 + 8 * qs, qph);
          t = tq[qs];
          __syncwarp();
          if (lane == 0) mbar_arrive(qempty0 + 8 * qs);
        } else {
          t = (int)(wid + local * nwork);
          if (t >= total_tiles) t = -1;
        }
        if (t < 0) break;
        int ui, rt_, nt_;
        decode_tile(p, t / p.ksplit, n_row_tiles, ui, rt_, nt_);
        const int ks_unit = p.units[ui].reserved, part = t % p.ksplit;
        const int ksteps = ks_unit * (part + 1) / p.ksplit - ks_unit * part / p.ksplit;
        const uint32_t acc = local & 1, aphase = (local >> 1) & 1;
        mbar_wait(tempty0 + 8 * acc, aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
            const uint64_t a_hi = make_desc_sw128(sa), a_lo = make_desc_sw128(sa + C::A_BYTES);
            const uint64_t b_hi = make_desc_sw128(sa + 2 * C::A_BYTES);
            const uint64_t b_lo = make_desc_sw128(sa + 2 * C::A_BYTES + C::B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);      // +32 B per K step inside the swizzle row
              if (TWO) {
                umma_bf16_2sm(d_tmem, a_lo + koff, b_hi + koff, idesc, (ks | k) != 0);
                umma_bf16_2sm(d_tmem, a_hi + koff, b_lo + koff, idesc, 1);
                umma_bf16_2sm(d_tmem, a_hi + koff, b_hi + koff, idesc, 1);
              } else {
                umma_bf16(d_tmem, a_lo + koff, b_hi + koff, idesc, (ks | k) != 0);
                umma_bf16(d_tmem, a_hi + koff, b_lo + koff, idesc, 1);
                umma_bf16(d_tmem, a_hi + koff, b_hi + koff, idesc, 1);
              }
            }
            if (TWO) umma_commit_2sm(empty0 + 8 * stage, 3);   // frees the slot in BOTH CTAs when these MMAs retire
            else umma_commit(empty0 + 8 * stage);       // frees the smem slot when these MMAs retire
            if (ks + 1 == ksteps) {
              if (TWO) umma_commit_2sm(tfull0 + 8 * acc, 3);
              else umma_commit(tfull0 + 8 * acc);         // accumulator complete -> epilogue
            }
          }
          __syncwarp();
          if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1; }
        }
        if (ksteps == 0) {                              // (a unit without taps: the epilogue still gets its signal)
          if (elect_one_sync()) {
            if (TWO) umma_commit_2sm(tfull0 + 8 * acc, 3);
            else umma_commit(tfull0 + 8 * acc);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ epilogue (4 or 8 warps) ================================
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int ew = warp - EPI_WARP0;                  // epilogue warp index (staging buffer)
    constexpr int NH = C::EPI_WARPS / 4;              // warps per lane quarter: each takes every NH-th column slab
    const int half = ew >> 2;
    uint32_t oslab = 0;                               // tensor stores issued by this warp (staging-buffer rotation)
    for (uint32_t local = 0;; ++local) {
      int t;
      if (DYN) {
        const uint32_t qs = local & 3, qph = (local >> 2) & 1;
        mbar_wait(qfull0 + 8 * qs, qph);
        t = tq[qs];
        __syncwarp();
        if (lane == 0) mbar_arrive(qempty0 + 8 * qs);
      } else {
        t = (int)(wid + local * nwork);
        if (t >= total_tiles) t = -1;
      }
      if (t < 0) break;
      int ui, rt, nt;
      decode_tile(p, t / p.ksplit, n_row_tiles, ui, rt, nt);
      if (TWO) rt = 2 * rt + (int)rank;
      const idv_unit_t unit = p.units[ui];
      // the accumulator-empty barrier the MMA issuer waits on (TWO: the even CTA's, 8 warps arrive)
      const uint32_t tempty_bar = TWO ? ((tempty0 + 8 * (local & 1)) & PEER_BIT_MASK) : (tempty0 + 8 * (local & 1));
      const uint32_t acc = local & 1, aphase = (local >> 1) & 1;
      mbar_wait(tfull0 + 8 * acc, aphase);
      tc_fence_after();
      const int r = rt * BM + q * 32 + lane;
      const bool row_ok = r < p.R;
      const int tt_row = p.Tp > 0 ? r % p.Tp : 1;
      const bool pad_row = p.Tp > 0 && p.head != 3 && (tt_row == 0 || (p.t_valid > 0 && tt_row > p.t_valid));
      const float* bias = ((p.bias2 != nullptr && tt_row == 1) ? p.bias2 : p.bias) + unit.bias_off + nt * BN;
      if (p.head == 3) {
        // ---- STFT epilogue: rows are (b, t) frames (Tp = frames per utterance, no pad rows), column pair
        //      (2k, 2k+1) = (re, im) of bin k; written to the reference layout (B, nbins, T, 2)
        const int T = p.Tp;
        const int b = r / T, t = r % T;
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 32 * NH) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c0, v);
          tmem_ld_wait();
          if (row_ok) {
            // optional second copy for the first encoder layer's tap-GEMM: split-bf16 activation rows of ONE plane,
            // row b*(T+1)+1+t (causal pad row in front of every utterance), column head_boff + 2*bin + part
            unsigned short* rows = reinterpret_cast<unsigned short*>(p.out);
            const long long ridx = ((long long)r + b + 1) * p.out_ld + p.head_boff;
            const int n0 = nt * BN + c0;
            // 32 consecutive columns of a row = 64 bytes of hi and of lo: 16-byte vectors when the column offset allows
            // (columns past 2*head_fout hold exact zeros - zero basis rows, no bias - which is what the padding needs)
            const bool vec = rows != nullptr && (p.head_boff & 7) == 0 && (p.out_ld & 7) == 0 && p.head_boff + n0 + 32 <= p.out_ld;
            if (vec) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  unsigned short h0, l0, h1, l1;
                  split_bf16(__uint_as_float(v[j + 2 * e]), h0, l0);
                  split_bf16(__uint_as_float(v[j + 2 * e + 1]), h1, l1);
                  hw[e] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                  lw[e] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                }
                *reinterpret_cast<uint4*>(rows + ridx + n0 + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(rows + p.out_hl + ridx + n0 + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int bin = (n0 + j) >> 1;
              if (bin < p.head_fout) {
                const float xr = __uint_as_float(v[j]), xi = __uint_as_float(v[j + 1]);
                *reinterpret_cast<float2*>(p.predict + ((long long)(b * p.head_fout + bin) * T + t) * 2) = make_float2(xr, xi);
                if (rows && !vec) {
                  unsigned short h0, l0, h1, l1;
                  split_bf16(xr, h0, l0);
                  split_bf16(xi, h1, l1);
                  *reinterpret_cast<unsigned int*>(rows + ridx + 2 * bin) = (unsigned)h0 | ((unsigned)h1 << 16);
                  *reinterpret_cast<unsigned int*>(rows + p.out_hl + ridx + 2 * bin) = (unsigned)l0 | ((unsigned)l1 << 16);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (TWO) mbar_arrive_cluster(tempty_bar); else mbar_arrive(tempty_bar); }
        continue;
      }
      if (p.head) {
        // ---- fused reconstruction head: up to 16 output bins per unit, written to the reference layout
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (TWO) mbar_arrive_cluster(tempty_bar); else mbar_arrive(tempty_bar); }
        if (row_ok && !pad_row) {
          const int T = p.Tp - 1;
          const int b = r / p.Tp, t = r % p.Tp - 1;
          const float b_r = __ldg(bias), b_i = __ldg(bias + 1);
          const int nbins = unit.out_ch_off;          // head units: out_f = first output bin, out_ch_off = bins (<= 16)
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int fo = unit.out_f + e;
            if (e >= nbins || fo >= p.head_fout || (e * NH) / 16 != half) continue;      // bins split over the warps of a quarter
            float yr = prelu_f(__uint_as_float(v[2 * e]) + b_r, p.slope);
            float yi = prelu_f(__uint_as_float(v[2 * e + 1]) + b_i, p.slope);
            if (p.head == 2) {
              // model/pvae_module.py:L2594-2609 — operation order kept
              const float mag = tanhf(sqrtf(yr * yr + yi * yi));
              const float ph = atan2f(yi / (mag + 1e-8f), yr / (mag + 1e-8f));
              const float2 X = __ldg(reinterpret_cast<const float2*>(p.stft_x + ((long long)(b * p.head_fout + fo) * T + t) * 2));
              const float in_mag = sqrtf(X.x * X.x + X.y * X.y);
              const float in_ph = atan2f(X.y, X.x);
              float sn, cs;
              sincosf(in_ph + ph, &sn, &cs);
              const float gm = in_mag * mag;
              yr = gm * cs;
              yi = gm * sn;
            }
            const int bo = b * p.head_bmul + p.head_boff;
            *reinterpret_cast<float2*>(p.predict + ((long long)(bo * p.head_fout + fo) * T + t) * 2) = make_float2(yr, yi);
            if (p.out) {
              // the same value as the K-major split-bf16 spectrum row the iSTFT's DFT GEMM reads (idv_spec_rows_split's
              // layout: row bo*T + t, columns 2*bin + part): no separate transpose pass over `predict`
              unsigned short h0, l0, h1, l1;
              split_bf16(yr, h0, l0);
              split_bf16(yi, h1, l1);
              unsigned short* rows = reinterpret_cast<unsigned short*>(p.out);
              const long long idx = ((long long)bo * T + t) * p.out_ld + 2 * fo;
              *reinterpret_cast<unsigned int*>(rows + idx) = (unsigned)h0 | ((unsigned)h1 << 16);
              *reinterpret_cast<unsigned int*>(rows + p.out_hl + idx) = (unsigned)l0 | ((unsigned)l1 << 16);
            }
          }
        }
        continue;
      }
      if (BN >= 64 && p.tma_out && (unit.out_ch_off & 63) == 0) {
        // ---- split-bf16 tile through shared memory and TMA stores.  A thread owns a ROW of the accumulator, so direct
        // stores scatter every 16-byte vector of a warp over 32 rows (32 L1 lines per store instruction: the narrow
        // layers, whose main loop is short, were bound by this epilogue).  Instead each epilogue warp stages its 32 rows
        // x 64 columns (hi and lo) in a 128B-swizzled buffer and one lane issues a tensor store of the box
        // {64 ch, 32 rows, 1 plane, hi|lo}; rows past R are clipped by the tensor map, pad rows carry zeros.
        const int r0 = rt * BM + q * 32;
        const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll 1
        for (int c0 = half * 64; c0 < BN; c0 += 64 * NH, ++oslab) {
          const uint32_t ob = oslab % (uint32_t)n_obufs;
          const uint32_t stg = (ob == 0 ? out_stage0 : out_stage_x + (ob - 1) * (uint32_t)C::OUT_STAGE_BYTES) + (uint32_t)ew * 8192u;
          if (lane == 0) {                                   // the box that used this staging buffer last has left it
            if (n_obufs == 1) bulk_wait_read<0>();
            else if (n_obufs == 2) bulk_wait_read<1>();
            else bulk_wait_read<2>();
          }
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c0 + 32 * h, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float x0 = __uint_as_float(v[j + 2 * e]) + __ldg(bias + c0 + 32 * h + j + 2 * e);
                float x1 = __uint_as_float(v[j + 2 * e + 1]) + __ldg(bias + c0 + 32 * h + j + 2 * e + 1);
                if (p.apply_prelu) { x0 = prelu_f(x0, p.slope); x1 = prelu_f(x1, p.slope); }
                if (pad_row) { x0 = 0.f; x1 = 0.f; }
                const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
              }
              const uint32_t chunk = (uint32_t)(4 * h + (j >> 3));                    // 16-byte chunk of the 128-byte row
              const uint32_t a = stg + (uint32_t)lane * 128u + ((chunk ^ sw) << 4);
              st_shared_v4(a, hw[0], hw[1], hw[2], hw[3]);
              st_shared_v4(a + 4096u, lw[0], lw[1], lw[2], lw[3]);
            }
          }
          fence_proxy_async();                               // generic-proxy writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0 && r0 < p.R) {
            const int n0 = nt * BN + c0;
            int plane = unit.out_f, col = unit.out_ch_off + n0;
            if (p.N > p.out_ld) {                            // columns wrap into consecutive output planes
              const int pa = n0 / p.out_ld;
              plane += pa;
              col = n0 - pa * p.out_ld;
            }
            if ((long long)(plane + 1) * p.out_plane <= p.out_hl) tma_store_4d(&tmO, stg, col, r0, plane, 0);
            bulk_commit_group();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (TWO) mbar_arrive_cluster(tempty_bar); else mbar_arrive(tempty_bar); }
        continue;
      }
      const long long obase0 = (long long)unit.out_f * p.out_plane + (long long)r * p.out_ld + unit.out_ch_off + nt * BN;
      const float* ws_row = nullptr;
      long long ws_part = 0;              // elements between the partial tiles of one output tile
      int my_part = 0;
      if (p.ksplit > 1) {
        // ---- split-K: every CTA stores its partial accumulator (plain stores), the ksplit CTAs of an output tile wait for
        // one another (they are co-resident: the grid is at most one CTA per SM), then CTA `part` reduces and finishes
        // the 32-column slabs j with j % ksplit == part - a reduce-scatter through the L2, summed in a fixed order
        const int tb = t / p.ksplit;
        my_part = t % p.ksplit;
        ws_part = (long long)BM * BN;
        // partial tile layout [BN / 4 column vectors][128 rows] float4: the 32 lanes (rows) of a warp store / load 512
        // consecutive bytes per instruction (row-major, one row per thread, scattered every 16-byte piece over its own line)
        float* ws_tile = p.ws + (long long)tb * p.ksplit * ws_part;
        float* mine = ws_tile + my_part * ws_part + (long long)(q * 32 + lane) * 4;
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 32 * NH) {
          if ((c0 >> 5) % p.ksplit == my_part) continue;          // my own slabs stay in TMEM
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(mine + (long long)((c0 + j) >> 2) * (BM * 4)) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * C::EPI_WARPS) : "memory");
        if (ew == 0 && lane == 0) {
          atomicAdd(p.ws_cnt + 2 * tb, 1u);
          long long t0 = 0;
          unsigned int spins = 0;
          while (*reinterpret_cast<volatile unsigned int*>(p.ws_cnt + 2 * tb) < (unsigned int)p.ksplit) {
            if ((++spins & 255u) == 0) {
              const long long now = clock64();
              if (t0 == 0) t0 = now;
              else if (now - t0 > WAIT_TIMEOUT_CYCLES) __trap();
            }
          }
          __threadfence();
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * C::EPI_WARPS) : "memory");
        ws_row = ws_tile + (long long)(q * 32 + lane) * 4;
      }
#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += 32 * NH) {
        // N > out_ld: the unit's columns wrap into consecutive output planes, out_ld columns each (two output planes of a
        // narrow transposed conv computed as ONE tile from the input planes they share); planes past the tensor are dropped
        long long obase = obase0;
        bool plane_ok = true;
        if (p.N > p.out_ld) {
          const int n0 = nt * BN + c0, pa = n0 / p.out_ld;
          obase = (long long)(unit.out_f + pa) * p.out_plane + (long long)r * p.out_ld + (n0 - pa * p.out_ld) - c0;
          plane_ok = p.out_hl <= 0 || (long long)(unit.out_f + pa + 1) * p.out_plane <= p.out_hl;
        }
        if (ws_row && (c0 >> 5) % p.ksplit != my_part) continue;   // split-K: another CTA of the tile finishes this slab
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c0, v);
        tmem_ld_wait();
        if (ws_row) {
          // own part from TMEM, the others from the workspace in part order 0, 1, ... (the own one takes its place: the
          // sum does not depend on which CTA computes it)
          float sum[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] = 0.f;
#pragma unroll 1
          for (int pp = 0; pp < p.ksplit; ++pp) {
            if (pp == my_part) {
#pragma unroll
              for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(v[j]);
            } else {
              const float* src = ws_row + pp * ws_part + (long long)(c0 >> 2) * (BM * 4);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 sv = __ldcg(reinterpret_cast<const float4*>(src + (long long)(j >> 2) * (BM * 4)));
                sum[j] += sv.x; sum[j + 1] += sv.y; sum[j + 2] += sv.z; sum[j + 3] += sv.w;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(sum[j]);
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]) + __ldg(bias + c0 + j);
          if (p.apply_prelu) x = prelu_f(x, p.slope);
          f[j] = pad_row ? 0.f : x;
        }
        if (row_ok && plane_ok && !(p.keep_pad && tt_row == 0)) {   // keep_pad: the pad row carries x[t-1] of the last step
          if (p.out_split) {
            __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(p.out) + obase + c0;
            __nv_bfloat16* ol = oh + p.out_hl;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat16 h0 = __float2bfloat16_rn(f[j + 2 * e]), h1 = __float2bfloat16_rn(f[j + 2 * e + 1]);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(f[j + 2 * e] - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(f[j + 2 * e + 1] - __bfloat162float(h1));
                hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
              }
              *reinterpret_cast<uint4*>(oh + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(ol + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          } else {
            float* of = reinterpret_cast<float*>(p.out) + obase + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(of + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (TWO) mbar_arrive_cluster(tempty_bar); else mbar_arrive(tempty_bar); }
      if (ws_row) {
        // the CTA of the tile that finishes last re-arms the tile's two counters for the next launch
        asm volatile("bar.sync 1, %0;" ::"n"(32 * C::EPI_WARPS) : "memory");
        if (ew == 0 && lane == 0) {
          const int tb = t / p.ksplit;
          if (atomicAdd(p.ws_cnt + 2 * tb + 1, 1u) == (unsigned int)(p.ksplit - 1)) {
            p.ws_cnt[2 * tb] = 0u;
            p.ws_cnt[2 * tb + 1] = 0u;
          }
        }
      }
    }
  }

  if (warp >= EPI_WARP0 && lane == 0) bulk_wait_all();      // this warp's tensor stores are complete
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();            // neither CTA leaves (or frees TMEM) while its peer may still touch it
  if (warp == EPI_WARP0) {
    tc_fence_after();
    if (TWO) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
  if (DYN && threadIdx.x == 0) {
    // last CTA out re-arms the scheduler slot for its next user
    if (atomicAdd(p.sched + 1, 1u) == gridDim.x - 1) {
      p.sched[0] = 0;
      p.sched[1] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------ host
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_map_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return IDV_E_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d) failed: CUresult %d (cols %llu rows %llu box %u %u)", (int)rc,
              (unsigned long long)cols, (unsigned long long)rows, box_cols, box_rows);
    return IDV_E_CUDA;
  }
  return IDV_OK;
}

int encode_map_3d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t blocks, uint32_t box_cols,
                  uint32_t box_rows, uint32_t box_blocks, bool swizzle128) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return IDV_E_CUDA;
  }
  cuuint64_t dims[3] = {cols, rows, blocks};
  cuuint64_t strides[2] = {cols * 2, cols * rows * 2};
  cuuint32_t box[3] = {box_cols, box_rows, box_blocks};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed: CUresult %d", (int)rc);
    return IDV_E_CUDA;
  }
  return IDV_OK;
}

static int encode_map_4d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t hl_stride_el,
                         uint32_t b0, uint32_t b1, uint64_t plane_stride_el = 0) {
  // dims (elements): {d0 = channels/k, d1 = rows/n, d2 = planes/slots, 2 = hi/lo}
  cuuint64_t dims[4] = {d0, d1, d2, 2};
  cuuint64_t strides[3] = {d0 * 2, (plane_stride_el ? plane_stride_el : d0 * d1) * 2, hl_stride_el * 2};
  cuuint32_t box[4] = {b0, b1, 1, 2};
  cuuint32_t es[4] = {1, 1, 1, 1};
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return IDV_E_CUDA;
  }
  CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                                       box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (dims %llu %llu %llu box %u %u)", (int)rc, (unsigned long long)d0,
              (unsigned long long)d1, (unsigned long long)d2, b0, b1);
    return IDV_E_CUDA;
  }
  return IDV_OK;
}

// Pool of self-resetting scheduler slots (2 x u32 each), one pool per device, slots handed out round-robin so that
// launches in flight on different streams never share one.
constexpr int SCHED_SLOTS = 2048;
static unsigned int* sched_slot() {
  static unsigned int* pool[64] = {nullptr};
  static unsigned int next[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!pool[dev]) {
    if (cudaMalloc(&pool[dev], SCHED_SLOTS * 2 * sizeof(unsigned int)) != cudaSuccess) return nullptr;
    if (cudaMemset(pool[dev], 0, SCHED_SLOTS * 2 * sizeof(unsigned int)) != cudaSuccess) return nullptr;
  }
  const unsigned int i = __sync_fetch_and_add(&next[dev], 1u) % SCHED_SLOTS;
  return pool[dev] + 2 * i;
}

// split-K workspace, one per device: the partial accumulators of up to KS_TILES CTAs (128 x 256 fp32 each) + two counters
// per output tile (zeroed once; the CTA that finishes a tile last re-arms them).  Launches of one stream are ordered (a
// dependent launch touches the workspace after its pdl_wait()); several streams: split-K is off (gemm_dynamic_tiles).
constexpr int KS_TILES = 160;
static int ks_workspace(float** ws, unsigned int** cnt) {
  static float* pool[64] = {nullptr};
  int dev = 0;
  IDV_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!pool[dev]) {
    const size_t bytes = (size_t)KS_TILES * BM * 256 * sizeof(float) + 2 * KS_TILES * sizeof(unsigned int);
    float* p = nullptr;
    IDV_CUDA(cudaMalloc(&p, bytes));
    IDV_CUDA(cudaMemset(p, 0, bytes));
    pool[dev] = p;
  }
  *ws = pool[dev];
  *cnt = reinterpret_cast<unsigned int*>(pool[dev] + (size_t)KS_TILES * BM * 256);
  return IDV_OK;
}

template <int BN, bool DYN>
static int launch2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& o, const Params& p,
                   int sms, cudaStream_t st) {
  using C = Cfg<BN>;
  IDV_CUDA(cudaFuncSetAttribute(tapgemm_tc_kernel<BN, DYN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                C::SMEM_BYTES));
  const int total = p.n_units * p.n_row_tiles * p.n_col_tiles * p.ksplit;
  const int grid = total < sms ? total : sms;
  IDV_CUDA(launch_pdl(tapgemm_tc_kernel<BN, DYN, false>, dim3(grid), dim3(C::THREADS), (size_t)C::SMEM_BYTES, st, a0, a1, w, o, p));
  return IDV_OK;
}

// CTA-pair kernel: clusters of 2, as many as fit the device at once (persistent)
template <int BN>
static int launch_pairs(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& o, Params& p,
                        cudaStream_t st) {
  using C = Cfg<BN, true>;
  auto kern = tapgemm_tc_kernel<BN, false, true>;
  IDV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = 1;
  static int max_pairs[64] = {0};
  int dev = 0;
  IDV_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!max_pairs[dev]) {
    cfg.gridDim = dim3(2);
    int n = 0;
    IDV_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    IDV_CHECK_ARG(n > 0, "idv_tapgemm_tc: no CTA pair fits the device");
    max_pairs[dev] = n;
  }
  const int total = p.n_units * ((p.n_row_tiles + 1) / 2) * p.n_col_tiles;
  const int pairs = total < max_pairs[dev] ? total : max_pairs[dev];
  cfg.gridDim = dim3(2 * pairs);
  cfg.numAttrs = option_launch_pdl() ? 2 : 1;
  p.sched = nullptr;
  IDV_CUDA(cudaLaunchKernelEx(&cfg, kern, a0, a1, w, o, p));
  return IDV_OK;
}

template <int BN>
static int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& o, Params& p, int sms,
                  cudaStream_t st) {
  if (option_dynamic_tiles()) {
    p.sched = sched_slot();
    IDV_CHECK_ARG(p.sched != nullptr, "idv_tapgemm_tc: could not allocate the tile-scheduler pool");
    return launch2<BN, true>(a0, a1, w, o, p, sms, st);
  }
  p.sched = nullptr;
  return launch2<BN, false>(a0, a1, w, o, p, sms, st);
}

}  // namespace tc
}  // namespace idv

extern "C" int idv_tapgemm_tc(const void* a0, int a0_cp, int a0_planes, const void* a1, int a1_cp, int a1_planes,
                              int R, int Tp, const void* wt, int kc_max, int n_slots, const float* bias, int N,
                              const idv_unit_t* units, const idv_tap_t* taps, int n_units, void* out, int out_ld,
                              int64_t out_plane, int64_t out_hl, int out_split, int apply_prelu, float prelu_slope,
                              int t_valid, void* stream) {
  return idv_tapgemm_tc_head(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, N, units,
                             taps, n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, prelu_slope, 0, 0,
                             1, 0, nullptr, nullptr, t_valid, stream);
}

static int tapgemm_tc_impl(const void* a0, int a0_cp, int a0_planes, const void* a1, int a1_cp, int a1_planes,
                           int R, int Tp, const void* wt, int kc_max, int n_slots, const float* bias, const float* bias2,
                           int N, const idv_unit_t* units, const idv_tap_t* taps, int n_units, void* out, int out_ld,
                           int64_t out_plane, int64_t out_hl, int out_split, int apply_prelu,
                           float prelu_slope, int head, int head_fout, int head_bmul, int head_boff,
                           const float* stft_x, float* predict, int t_valid, int min_ksteps, void* stream);

extern "C" int idv_tapgemm_tc_head(const void* a0, int a0_cp, int a0_planes, const void* a1, int a1_cp, int a1_planes,
                                   int R, int Tp, const void* wt, int kc_max, int n_slots, const float* bias, int N,
                                   const idv_unit_t* units, const idv_tap_t* taps, int n_units, void* out, int out_ld,
                                   int64_t out_plane, int64_t out_hl, int out_split, int apply_prelu,
                                   float prelu_slope, int head, int head_fout, int head_bmul, int head_boff,
                                   const float* stft_x, float* predict, int t_valid, void* stream) {
  return tapgemm_tc_impl(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, nullptr, N, units,
                         taps, n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, prelu_slope, head,
                         head_fout, head_bmul, head_boff, stft_x, predict, t_valid, 0, stream);
}

extern "C" int idv_tapgemm_tc_splitk(const void* a0, int a0_cp, int a0_planes, const void* a1, int a1_cp, int a1_planes,
                                     int R, int Tp, const void* wt, int kc_max, int n_slots, const float* bias,
                                     const float* bias_first, int N, const idv_unit_t* units, const idv_tap_t* taps,
                                     int n_units, void* out, int out_ld, int64_t out_plane, int64_t out_hl, int out_split,
                                     int apply_prelu, float prelu_slope, int t_valid, int min_ksteps, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(!bias_first || Tp > 1 || Tp < -1, "idv_tapgemm_tc_splitk: a first-frame bias needs |Tp| = frames per utterance + 1");
  IDV_CHECK_ARG(min_ksteps >= 0, "idv_tapgemm_tc_splitk: min_ksteps = the smallest unit.reserved of the table (0: no split)");
  return tapgemm_tc_impl(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, bias_first, N,
                         units, taps, n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, prelu_slope, 0, 0,
                         1, 0, nullptr, nullptr, t_valid, min_ksteps, stream);
}

extern "C" int idv_tapgemm_tc_b2(const void* a0, int a0_cp, int a0_planes, const void* a1, int a1_cp, int a1_planes,
                                 int R, int Tp, const void* wt, int kc_max, int n_slots, const float* bias,
                                 const float* bias_first, int N, const idv_unit_t* units, const idv_tap_t* taps,
                                 int n_units, void* out, int out_ld, int64_t out_plane, int64_t out_hl, int out_split,
                                 int apply_prelu, float prelu_slope, int t_valid, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(bias_first && (Tp > 1 || Tp < -1), "idv_tapgemm_tc_b2: needs the first-frame bias and |Tp| = frames per utterance + 1");
  return tapgemm_tc_impl(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, bias_first, N,
                         units, taps, n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, prelu_slope, 0, 0,
                         1, 0, nullptr, nullptr, t_valid, 0, stream);
}

static int tapgemm_tc_impl(const void* a0, int a0_cp, int a0_planes, const void* a1, int a1_cp, int a1_planes,
                           int R, int Tp, const void* wt, int kc_max, int n_slots, const float* bias, const float* bias2,
                           int N, const idv_unit_t* units, const idv_tap_t* taps, int n_units, void* out, int out_ld,
                           int64_t out_plane, int64_t out_hl, int out_split, int apply_prelu,
                           float prelu_slope, int head, int head_fout, int head_bmul, int head_boff,
                           const float* stft_x, float* predict, int t_valid, int min_ksteps, void* stream) {
  using namespace idv;
  using namespace idv::tc;
  IDV_CHECK_ARG(a0 && wt && bias && units && taps && (out || head), "idv_tapgemm_tc: null pointer");
  IDV_CHECK_ARG(head >= 0 && head <= 3, "idv_tapgemm_tc: epilogue mode must be 0..3");
  if (head == 3) {
    IDV_CHECK_ARG(Tp > 0 && predict && head_fout > 0 && R % Tp == 0, "idv_tapgemm_tc: STFT epilogue needs Tp = frames per utterance");
    IDV_CHECK_ARG(!out || (head_boff >= 0 && head_boff % 2 == 0 && out_ld >= head_boff + 2 * head_fout && out_ld % 2 == 0 && out_hl > 0),
                  "idv_tapgemm_tc: STFT epilogue with activation rows needs out_ld >= head_boff + 2 * head_fout and out_hl");
  } else if (head) {
    IDV_CHECK_ARG(N == 32 && Tp > 1 && predict && head_fout > 0 && head_bmul > 0 && head_boff >= 0 && (head != 2 || stft_x),
                  "idv_tapgemm_tc: head mode needs N == 32, Tp, predict (and stft_x for the mask head)");
    IDV_CHECK_ARG(!out || (out_ld >= 2 * head_fout && out_ld % 2 == 0 && out_hl > 0),
                  "idv_tapgemm_tc: head mode with spectrum rows needs out_ld >= 2 * head_fout and the hi/lo stride out_hl");
  }
  IDV_CHECK_ARG(R > 0 && n_units > 0 && a0_planes > 0 && n_slots > 0, "idv_tapgemm_tc: empty problem");
  IDV_CHECK_ARG(N >= 32 && N % 32 == 0 && (N % 64 == 0 || N == 32), "idv_tapgemm_tc: N=%d must be 32 or a multiple of 64", N);
  IDV_CHECK_ARG(head || N <= out_ld || (out_ld % 32 == 0 && N % out_ld == 0),
                "idv_tapgemm_tc: N=%d > out_ld=%d wraps into consecutive planes and needs out_ld %% 32 == 0, N %% out_ld == 0", N, out_ld);
  IDV_CHECK_ARG(a0_cp % 8 == 0 && kc_max % 64 == 0 && (head || out_ld % 8 == 0) && (!a1 || a1_cp % 8 == 0),
                "idv_tapgemm_tc: channel counts must be multiples of 8 and kc_max of 64");
  int dev = 0, sms = 0;
  IDV_CUDA(cudaGetDevice(&dev));
  IDV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // Widest N tile by default (fewest A re-reads).  Small problems (streaming steps, short batches) would leave most
  // SMs idle and, with one CTA streaming a whole weight slice through a 2-stage ring, run at the latency of single
  // TMA round trips: narrow the tile (more CTAs, 3-4 stage ring) until every SM has a tile.
  int BN = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32));
  // split-K (Params::ksplit) instead of narrow tiles when the caller knows the K steps of its units: the WIDEST tile
  // (fewest re-reads of the activation rows) whose K range divides over enough CTAs to fill the device
  int ksplit = 1;
  if (head == 0 && min_ksteps >= 2 && option_splitk() && !option_dynamic_tiles() && BN >= 64) {
    const long long base = (long long)n_units * cdiv(R, BM) * (N / BN);
    // (measured: with 64-column tiles on more than half of the SMs the unsplit kernel is as fast - the reduction costs a
    // few microseconds per launch)
    const long long narrow = (long long)n_units * cdiv(R, BM) * (N / 64);
    if (base * 2 <= sms && narrow * 2 <= sms) {
      long long s = sms / base;                    // (base * s CTAs <= sms <= KS_TILES partial tiles)
      if (s > min_ksteps) s = min_ksteps;
      if (s > BN / 32) s = BN / 32;                // a CTA finishes whole 32-column slabs
      ksplit = (int)s;
    }
  }
  if ((head == 0 || head == 3) && ksplit == 1)
    while (BN > 64 && N % (BN / 2) == 0 && (long long)n_units * cdiv(R, BM) * (N / BN) < sms) BN /= 2;
  CUtensorMap mA0, mA1, mW;
  int rc = encode_map_4d(&mA0, a0, a0_cp, R, a0_planes, (uint64_t)a0_planes * R * a0_cp, BK, BM);
  if (rc) return rc;
  if (a1) {
    rc = encode_map_4d(&mA1, a1, a1_cp, R, a1_planes, (uint64_t)a1_planes * R * a1_cp, BK, BM);
    if (rc) return rc;
  } else {
    mA1 = mA0;
  }
  // CTA pairs for the wide tiles (every CTA stages half of the weight tile) unless tiles are claimed dynamically
  // (large problems only: a streaming step with a handful of row tiles keeps one CTA per tile - measured: 9 row tiles
  // as 5 pairs were 10 % slower - and a small odd tile count would waste a tenth of the pairs)
  const int n_rt = cdiv(R, BM);
  const bool pairs = ksplit == 1 && option_gemm_pairs() && !option_dynamic_tiles() && n_rt >= 8 && (n_rt % 2 == 0 || n_rt >= 32);
  rc = encode_map_4d(&mW, wt, kc_max, N, n_slots, (uint64_t)n_slots * N * kc_max, BK, pairs ? BN / 2 : BN);
  if (rc) return rc;
  Params p;
  p.R = R; p.Tp = Tp < 0 ? -Tp : Tp; p.keep_pad = Tp < 0; p.N = N; p.n_units = n_units; p.t_valid = t_valid;
  p.n_row_tiles = cdiv(R, BM); p.n_col_tiles = N / BN;
  p.units = units; p.taps = taps; p.bias = bias; p.bias2 = bias2; p.out = out; p.out_ld = out_ld;
  p.out_plane = out_plane; p.out_hl = out_hl; p.out_split = out_split; p.apply_prelu = apply_prelu; p.slope = prelu_slope;
  p.head = head; p.head_fout = head_fout; p.head_bmul = head_bmul; p.head_boff = head_boff;
  p.stft_x = stft_x; p.predict = predict;
  p.order = option_tile_order();
  p.ksplit = ksplit; p.ws = nullptr; p.ws_cnt = nullptr;
  if (ksplit > 1) {
    rc = ks_workspace(&p.ws, &p.ws_cnt);
    if (rc) return rc;
  }
  // TMA-store epilogue: split-bf16 outputs whose rows are 128-byte multiples, not the streaming steps (their pad rows
  // must stay untouched: a tensor store writes whole boxes); units with a column offset that is not a multiple of 64
  // fall back to direct stores inside the kernel
  CUtensorMap mO = mA0;
  p.tma_out = 0;
  if (option_tma_store() && ksplit == 1 && head == 0 && out_split && Tp >= 0 && BN >= 64 && out_ld % 64 == 0 && out_plane > 0 && out_hl > 0 &&
      out_hl % out_plane == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    rc = encode_map_4d(&mO, out, (uint64_t)out_ld, (uint64_t)R, (uint64_t)(out_hl / out_plane), (uint64_t)out_hl, BK, 32,
                       (uint64_t)out_plane);
    if (rc) return rc;
    p.tma_out = 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (pairs) {
    switch (BN) {
      case 256: return launch_pairs<256>(mA0, mA1, mW, mO, p, st);
      case 128: return launch_pairs<128>(mA0, mA1, mW, mO, p, st);
      case 64: return launch_pairs<64>(mA0, mA1, mW, mO, p, st);
      default: return launch_pairs<32>(mA0, mA1, mW, mO, p, st);
    }
  }
  switch (BN) {
    case 256: return launch<256>(mA0, mA1, mW, mO, p, sms, st);
    case 128: return launch<128>(mA0, mA1, mW, mO, p, sms, st);
    case 64: return launch<64>(mA0, mA1, mW, mO, p, sms, st);
    default: return launch<32>(mA0, mA1, mW, mO, p, sms, st);
  }
}
