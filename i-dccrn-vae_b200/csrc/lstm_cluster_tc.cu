// Two-layer complex-LSTM recurrence for FEW utterances per chunk (8 or 16): one thread-block CLUSTER per (module, role) and
// chunk, the weights resident in tensor memory, the hidden state exchanged between the CTAs of a cluster through distributed
// shared memory.  Larger batches run as independent chunks of 16 utterances in the same launch (see ClusterParams).
//
// Why a second kernel: the wavefront kernel (csrc/lstm_wave_tc.cu) moves h(t) through the L2 - stores, fence, counter,
// poll, TMA load: 4.5 of its 7-8 us per step are gate math + publish + propagation, whatever the batch.  With <= 8 utterances
// a step's h is NR = 16 rows x H (32 rows for <= 16 utterances), i.e. 24 KB as split bf16 (H = 384): small enough that EVERY
// CTA of a (module, role) can receive all of it in its own shared memory each step.  So here
//   * the problem is transposed: D[gate columns (M = 128 TMEM lanes)][rows (N = NR)] = W[gate cols][K = H] x h^T[K][rows];
//     W (bf16 hi AND lo, this CTA's 4 * UPC gate columns) is the A operand and stays in TENSOR MEMORY for the whole
//     sequence (tcgen05.mma with A from TMEM: lane = gate column, two K elements per 32-bit column: H/2 columns each for
//     hi and lo next to the 2 * NR accumulator columns) - with W in shared memory (first version) the 48 small MMAs of a
//     step took 1.9 us; h^T is the B operand in the no-swizzle K-major canonical layout ([k-core][hi | lo][8-row group]
//     [8 rows][8 k]), in which the UPC hidden units a CTA produces are ONE contiguous block of UPC * 4 * NR bytes;
//   * after the gates a CTA pushes its block into the B buffer of every CTA of the cluster with one
//     cp.async.bulk.shared::cluster per destination, completing on the DESTINATION's mbarrier: the consumer's MMA warp
//     simply waits for H * 4 * NR bytes of transactions - no fence, no counter, no poll on the recurrence chain;
//   * the split product runs as 2 MMAs per K step: W_hi x [h_hi | h_lo] (N = 2 NR: the hi and lo row groups of a k-core
//     are adjacent) and W_lo x h_hi on top of its first half; the epilogue adds the two halves.  TMEM addresses are
//     compile-time constants and the issue runs under elect.sync (48 MMAs in 0.48 us; 1.76 us from a `lane == 0` branch);
//   * epilogue thread = (gate column = TMEM lane 4 * unit + gate, RPT = NR / EPW rows), EPW = 4 warps per lane quarter;
//     the four gates of a unit sit in four neighbouring lanes and are exchanged with 3 warp shuffles per kept row, after
//     which lane `gate` owns RPT / 4 rows of the unit (cell state in registers).
// Roles as in the wavefront kernel: L0 | IP (layer-1 input projection) | L1, one cluster each per module = 6 clusters per
// chunk.  Between clusters the data goes through global memory once (L0 -> IP: h0(t) as a bulk store of the staged block +
// a release counter; IP -> L1: G1(t) fp32 + counter), T deep, so there is no back-pressure between the roles and the
// latency of these hops only fills the pipeline.  The 6 clusters of a chunk should be co-resident (checked with
// cudaOccupancyMaxActiveClusters; the caller falls back to the wavefront kernel otherwise).  Measurements: DESIGN.md 4.2b.
#include <stdlib.h>

#include "tc_common.cuh"

namespace idv {
namespace tc {

constexpr int CL_EPI_WARP0 = 4;           // warp 0 loader, 1 MMA issuer, 2 publisher, 3 idle, 4 .. 4 + 4 * EPW epilogue
constexpr int CL_SYNC_STRIDE = 32;        // uint32 between the two counters of a module (one 128-byte line each)
                                          // TMEM: [accumulator 2 * NR][W_hi H/2][W_lo H/2] columns

struct ClusterParams {
  const float* g0;
  long long g_m_off, g_p_off;
  int g_ld;
  const float* bias1;                     // fp32 [2 m][CS][128 lanes]
  const unsigned short* w[3];             // W_hh0, W_ih1, W_hh1: bf16 [2 hl][2 m][CS][4*UPC][H]
  int NB, T, Tsteps, H, CS;
  // a launch runs ceil(NB / (NR/2)) independent CHUNKS of NR/2 utterances: blockIdx.y = chunk * 6 + (module, role); every
  // chunk has its own six clusters, counters and exchange buffers.  The dependencies inside a chunk are acyclic (L0 -> IP ->
  // L1, T-deep buffers, no back-pressure), so any number of chunks is deadlock-free whatever the block scheduler does; the
  // chunks that are co-resident (clusters do not span GPCs: 5 chunks at H = 128, 1 at H = 384 on a B200) run concurrently.
  long long hx0_chunk, g1x_chunk;         // elements between the chunks' buffers
  float* hseq1;                           // fp32 [4][R][H]
  unsigned short* hx0;                    // bf16 [T][2 m][H/8 k-cores][2 hi,lo][NRG][8 rows][8 k]
  float* g1x;                             // fp32 [T][2 m][CS][128 lanes][NR rows]
  unsigned int* sync;                     // [chunk][2 m][A, B] step counters
  unsigned long long* dbg;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// shared memory of this CTA -> shared memory of a CTA of the cluster; completes on the destination CTA's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ void bulk_load_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void bulk_store_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// UMMA shared-memory descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes;
// LBO = bytes between the two core matrices of a K = 16 step, SBO = bytes between 8-row groups
// (cute/atom/mma_traits_sm100.hpp: INTERLEAVE ((8,n),2):((1,SBO),LBO) in 16-byte units)
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 8 columns (16 bf16 of K per lane) at a_tmem
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp writes lane (lane base + t)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned long long cgtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void cl_wait_counter(const unsigned int* ctr, long long target) {
  if (target <= 0) return;
  long long t0 = 0;
  unsigned int spins = 0;
  while (true) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if ((long long)v >= target) break;
    if ((++spins & 255u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > WAIT_TIMEOUT_CYCLES) __trap();
    }
  }
}
__device__ __forceinline__ float sel4(int i, float a, float b, float c, float d) {
  return i == 0 ? a : (i == 1 ? b : (i == 2 ? c : d));
}
// slots per (role, step): 0 operand landed, 1 MMAs issued, 2 accumulator seen by the epilogue, 3 block staged,
// 4 block pushed to the cluster, 5 published to the next role
#define CL_DBG(slot)                                                                         \
  do {                                                                                       \
    if (p.dbg && rank == 0 && m == 0 && chunk == 0 && t >= 300 && t < 304) p.dbg[(role * 4 + (t - 300)) * 8 + (slot)] = cgtime(); \
  } while (0)

// EPW epilogue warps per TMEM lane quarter: a thread owns RPT = 16 / EPW of the 16 rows of its gate column (a single warp per
// quarter took 1.76 us for its ~640 dependent instructions per step - issue-latency bound, measured)
template <int UPC, int EPW, int NR>
__global__ void __launch_bounds__(32 * (CL_EPI_WARP0 + 4 * EPW), 1)
lstm_cluster_tc_kernel(const ClusterParams p) {
  constexpr int MROWS = 4 * UPC;                       // gate columns of this CTA = valid TMEM lanes
  constexpr int NRG = NR / 8;                          // 8-row groups
  constexpr int ACC_COLS = 2 * NR;
  constexpr int RPT = NR / EPW;                        // rows per epilogue thread
  constexpr int KEEP = RPT / 4;                        // rows a lane keeps after the quad exchange
  constexpr int N_EPI_WARPS = 4 * EPW;
  static_assert(RPT == 4 || RPT == 8 || RPT == 16, "rows per epilogue thread");
  constexpr int SB = UPC * 4 * NR;                     // bytes of the block of h this CTA produces per step
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int H = p.H, CS = p.CS, T = p.Tsteps;
  const int BB = H * 4 * NR;                           // bytes of one B buffer (all of h, hi and lo)
  uint8_t* bbuf = smem;                                // 2 B buffers
  uint8_t* stg = bbuf + 2 * BB;                        // 2 staging blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + 2 * SB);
  const uint32_t wfull = smem_u32(bars), full0 = wfull + 8, empty0 = full0 + 16, accfull = empty0 + 16,
                 accempty = accfull + 8, staged0 = accempty + 8, pubdone0 = staged0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const uint32_t smem_b = smem_u32(bbuf), smem_s = smem_u32(stg);
  // ALL of tensor memory is allocated, so the allocation starts at column 0, lane 0: the MMA issuer then works with
  // compile-time TMEM addresses in uniform registers (an address read back from shared memory costs an R2UR + elect loop
  // per tcgen05.mma: 37 ns per MMA, measured - 1.76 of the 4.6 us of a step)
  constexpr uint32_t tmem_cols = 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();             // = blockIdx.x (clusters span x)
  const int chunk = blockIdx.y / 6, mr = blockIdx.y % 6;
  const int m = mr / 3, role = mr % 3;                 // 0 = L0, 1 = IP, 2 = L1
  const int Tp = p.T + 1;
  const long long R = (long long)p.NB * Tp;
  const int b0 = chunk * (NR / 2);                     // first utterance of this chunk
  const int NBc = p.NB - b0 < NR / 2 ? p.NB - b0 : NR / 2;
  unsigned short* const hx0 = p.hx0 + chunk * p.hx0_chunk;
  float* const g1x = p.g1x + chunk * p.g1x_chunk;
  unsigned int* const ctrA = p.sync + (chunk * 4 + m * 2) * CL_SYNC_STRIDE;
  unsigned int* const ctrB = ctrA + CL_SYNC_STRIDE;

  if (warp == 1 && lane == 0) {
    mbar_init(wfull, N_EPI_WARPS);
    for (int b = 0; b < 2; ++b) {
      mbar_init(full0 + 8 * b, 1);
      mbar_init(empty0 + 8 * b, 1);
      mbar_init(staged0 + 8 * b, N_EPI_WARPS);
      mbar_init(pubdone0 + 8 * b, 1);
    }
    mbar_init(accfull, 1);
    mbar_init(accempty, N_EPI_WARPS);
    fence_barrier_init();
    // first fill of either B buffer by the cluster (h(0) -> buffer 1, h(1) -> buffer 0); IP arms per load
    if (role != 1) {
      if (T > 1) mbar_expect_tx(full0 + 8, (uint32_t)BB);
      if (T > 2) mbar_expect_tx(full0, (uint32_t)BB);
    }
  }
  if (warp == CL_EPI_WARP0) {
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (*tmem_slot != 0) __trap();                       // (cannot happen: the whole tensor memory was allocated)
  constexpr uint32_t tmem_base = 0;
  const uint32_t tmem_whi = ACC_COLS, tmem_wlo = ACC_COLS + (uint32_t)(H / 2);

  if (warp == 0) {
    // ================================ loader (IP only) ================================
    if (lane == 0 && role == 1) {
      // layer-1 input projection: h0(t), published by the L0 cluster of this module, from global memory
      for (int t = 0; t < T; ++t) {
        const int b = t & 1;
        cl_wait_counter(ctrA, (long long)CS * (t + 1));
        fence_proxy_async_global();
        mbar_wait(empty0 + 8 * b, ((t >> 1) & 1) ^ 1);
        mbar_expect_tx(full0 + 8 * b, (uint32_t)BB);
        bulk_load_g2s(smem_b + b * BB, reinterpret_cast<const uint8_t*>(hx0) + ((long long)t * 2 + m) * BB, (uint32_t)BB,
                      full0 + 8 * b);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // the whole warp runs the loop; the instructions are issued under elect.sync (ptxas then knows that exactly one
    // thread is active: a `lane == 0` branch makes it wrap every tcgen05.mma in an elect / branch loop)
    constexpr uint32_t idesc32 = make_idesc_mn(128, 2 * NR), idesc16 = make_idesc_mn(128, NR);
    constexpr uint32_t LBO = 2 * NRG * 128, SBO = 128;
    const int KS = H / UMMA_K;
    mbar_wait(wfull, 0);
    tc_fence_after();
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      if (role == 1) {
        mbar_wait(full0 + 8 * b, (t >> 1) & 1);
      } else if (t > 0) {
        const int fill = b ? (t - 1) >> 1 : (t >> 1) - 1;
        mbar_wait(full0 + 8 * b, fill & 1);
      }
      mbar_wait(accempty, (t & 1) ^ 1);
      tc_fence_after();
      if (elect_one_sync()) {
        CL_DBG(0);
        if (role != 1 && t > 0 && t + 2 < T) mbar_expect_tx(full0 + 8 * b, (uint32_t)BB);   // arm the next fill of this buffer (h(t+1))
        if (role != 1 && t == 0) {
          mbar_arrive(accfull);            // h(-1) = 0: nothing to multiply, the epilogue takes the accumulator as zero
        } else {
          const uint64_t bd = make_desc_nosw(smem_b + b * BB, LBO, SBO);
#pragma unroll 4
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t b_k = bd + (uint64_t)((ks * 2 * (int)LBO) >> 4);
            umma_bf16_ts(tmem_base, tmem_whi + ks * 8, b_k, idesc32, ks != 0);    // [W_hi h_hi | W_hi h_lo]
            umma_bf16_ts(tmem_base, tmem_wlo + ks * 8, b_k, idesc16, 1);          // + W_lo h_hi
          }
          if (role == 1) umma_commit(empty0 + 8 * b);
          umma_commit(accfull);
        }
        CL_DBG(1);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ================================ publisher (to the next role, through global memory) ================================
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      mbar_wait(staged0 + 8 * b, (t >> 1) & 1);
      if (lane == 0) {
        if (role == 0) {
          // this CTA's block of h0(t) for the input-projection cluster
          bulk_store_s2g(reinterpret_cast<uint8_t*>(hx0) + ((long long)t * 2 + m) * BB + rank * SB, smem_s + b * SB, (uint32_t)SB);
          bulk_commit_group();
          bulk_wait_all();
          fence_proxy_async_global();
          __threadfence();
          atomicAdd(ctrA, 1u);
          CL_DBG(5);
        } else if (role == 1) {
          __threadfence();                 // cumulative over the epilogue warps' G1 stores (synchronised by the mbarrier)
          atomicAdd(ctrB, 1u);
          CL_DBG(5);
        }
        mbar_arrive(pubdone0 + 8 * b);
      }
      __syncwarp();
    }
  } else if (warp >= CL_EPI_WARP0) {
    // ================================ epilogue: thread = (gate column, RPT of the 16 rows) ================================
    const int q = (warp - CL_EPI_WARP0) & 3;           // TMEM lane quarter (= warp % 4)
    const int sub = (warp - CL_EPI_WARP0) >> 2;        // which RPT rows: [RPT * sub, RPT * sub + RPT)
    const int L = q * 32 + lane;                       // TMEM lane = 4 * local unit + gate
    const int g = L & 3, ul = L >> 2;
    const bool lane_ok = L < MROWS;
    const int unit = rank * UPC + ul;
    const int row0 = RPT * sub;
    // ---- this CTA's weight rows -> tensor memory (lane = gate column, two K elements per column), once
    if (q * 32 < MROWS) {
      const unsigned short* wsrc = p.w[role];
      const int nchunk = H / 64;                        // chunks of 32 columns per plane
#pragma unroll 1
      for (int ci = sub; ci < 2 * nchunk; ci += EPW) {
        const int hl = ci / nchunk, c = (ci % nchunk) * 32;
        const uint4* row = reinterpret_cast<const uint4*>(wsrc + ((((long long)hl * 2 + m) * CS + rank) * MROWS + (lane_ok ? L : 0)) * H);
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 x = __ldg(row + (c >> 2) + i);
          v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
        }
        tmem_st32((hl ? tmem_wlo : tmem_whi) + ((uint32_t)(q * 32) << 16) + c, v);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(wfull);
    // DSMEM destinations of this CTA's block: epilogue warp ew pushes to the CTAs ew, ew + N_EPI_WARPS, ... (one elected
    // thread per copy: 12 copies issued by the lanes of ONE warp serialise at ~37 ns each, measured 0.45 us per step)
    const int ew = warp - CL_EPI_WARP0;
    const uint32_t my_blk = smem_b + rank * SB;
    float cst[KEEP];                                   // cell state of (unit, rows row0 + KEEP * g ..)
#pragma unroll
    for (int j = 0; j < KEEP; ++j) cst[j] = 0.f;
    const float bias = (role == 1 && lane_ok) ? __ldg(p.bias1 + ((long long)m * CS + rank) * 128 + L) : 0.f;
    const float a_scale = g == 2 ? 2.f : -1.f, a_num = g == 2 ? 2.f : 1.f;
    const uint32_t tacc = ((uint32_t)(q * 32) << 16) + (uint32_t)row0;
    for (int t = 0; t < T; ++t) {
      const int b = t & 1;
      float gin[RPT];
      if (role == 0) {
        const float* gp = p.g0 + m * p.g_m_off + (long long)g * H + unit;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int n = row0 + i, part = n / (NR / 2), utt = n % (NR / 2);
          gin[i] = (lane_ok && utt < NBc) ? __ldg(gp + part * p.g_p_off + ((long long)(b0 + utt) * Tp + 1 + t) * p.g_ld) : 0.f;
        }
      } else if (role == 2) {
        if (lane == 0) cl_wait_counter(ctrB, (long long)CS * (t + 1));
        __syncwarp();
        const float* gp = g1x + ((((long long)t * 2 + m) * CS + rank) * 128 + L) * NR + row0;
#pragma unroll
        for (int i = 0; i < RPT; i += 4) {
          const float4 v = ldcg4(gp + i);
          gin[i] = v.x; gin[i + 1] = v.y; gin[i + 2] = v.z; gin[i + 3] = v.w;
        }
      }
      mbar_wait(accfull, t & 1);
      tc_fence_after();
      if (warp == CL_EPI_WARP0 && lane == 0) CL_DBG(2);
      uint32_t v[RPT], v2[RPT];                        // columns [row0, row0 + RPT) of W_hi h_hi + W_lo h_hi and of W_hi h_lo
      if constexpr (RPT == 16) {
        tmem_ld16(tacc, v); tmem_ld16(tacc + NR, v2);
      } else if constexpr (RPT == 8) {
        tmem_ld8(tacc, v); tmem_ld8(tacc + NR, v2);
      } else {
        tmem_ld4(tacc, v); tmem_ld4(tacc + NR, v2);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accempty);
      if (warp == CL_EPI_WARP0 && lane == 0) CL_DBG(6);
      float acc[RPT];
      const bool zero_acc = role != 1 && t == 0;
#pragma unroll
      for (int i = 0; i < RPT; ++i) acc[i] = zero_acc ? 0.f : __uint_as_float(v[i]) + __uint_as_float(v2[i]);
      // the staging block / G1 slot of step t-2 has been consumed by the publisher
      if (t >= 2) mbar_wait(pubdone0 + 8 * b, ((t >> 1) - 1) & 1);
      if (role == 1) {
        if (lane_ok) {
          float* gp = g1x + ((((long long)t * 2 + m) * CS + rank) * 128 + L) * NR + row0;
#pragma unroll
          for (int i = 0; i < RPT; i += 4)
            *reinterpret_cast<float4*>(gp + i) = make_float4(acc[i] + bias, acc[i + 1] + bias, acc[i + 2] + bias, acc[i + 3] + bias);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(staged0 + 8 * b);
        continue;
      }
      // own gate's non-linearity on the thread's rows (sigmoid, or tanh for the cell candidate), as in lstm_wave_tc.cu
      float act[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const float e = __expf(a_scale * (acc[i] + gin[i]));
        const float qv = __fdividef(a_num, 1.f + e);
        act[i] = g == 2 ? 1.f - qv : qv;
      }
      // quad exchange: lane `g` of a unit keeps rows [KEEP * g, KEEP * g + KEEP) of the thread's RPT rows and receives
      // the other three gates for them
      float own[KEEP], r1[KEEP], r2[KEEP], r3[KEEP];
#pragma unroll
      for (int j = 0; j < KEEP; ++j) {
        own[j] = sel4(g, act[j], act[KEEP + j], act[2 * KEEP + j], act[3 * KEEP + j]);
        r1[j] = __shfl_xor_sync(0xffffffffu, sel4(g ^ 1, act[j], act[KEEP + j], act[2 * KEEP + j], act[3 * KEEP + j]), 1);
        r2[j] = __shfl_xor_sync(0xffffffffu, sel4(g ^ 2, act[j], act[KEEP + j], act[2 * KEEP + j], act[3 * KEEP + j]), 2);
        r3[j] = __shfl_xor_sync(0xffffffffu, sel4(g ^ 3, act[j], act[KEEP + j], act[2 * KEEP + j], act[3 * KEEP + j]), 3);
      }
      float hn[KEEP];
#pragma unroll
      for (int j = 0; j < KEEP; ++j) {
        const float ig = sel4(g, own[j], r1[j], r2[j], r3[j]);          // gate k lives in lane g ^ (g ^ k)
        const float fg = sel4(g ^ 1, own[j], r1[j], r2[j], r3[j]);
        const float gg = sel4(g ^ 2, own[j], r1[j], r2[j], r3[j]);
        const float og = sel4(g ^ 3, own[j], r1[j], r2[j], r3[j]);
        cst[j] = fg * cst[j] + ig * gg;
        hn[j] = og * (1.f - __fdividef(2.f, __expf(2.f * cst[j]) + 1.f));
      }
      if (warp == CL_EPI_WARP0 && lane == 0) CL_DBG(7);
      if (lane_ok) {
        // staging block in the B-operand layout: [k-core (ul / 8)][hi | lo][row group][row % 8][k % 8]
        uint8_t* sp = stg + b * SB + (ul >> 3) * (2 * NRG * 128) + (ul & 7) * 2;
#pragma unroll
        for (int j = 0; j < KEEP; ++j) {
          const int n = row0 + KEEP * g + j;
          unsigned short hi, lo;
          split_bf16(hn[j], hi, lo);
          uint8_t* rp = sp + (n >> 3) * 128 + (n & 7) * 16;
          *reinterpret_cast<unsigned short*>(rp) = hi;
          *reinterpret_cast<unsigned short*>(rp + NRG * 128) = lo;
        }
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * N_EPI_WARPS) : "memory");
      if (warp == CL_EPI_WARP0 && lane == 0) CL_DBG(3);
      // this CTA's block of h(t) -> buffer (t+1)&1 of every CTA of the cluster (h(T-1) has no consumer here)
      if (t + 1 < T && elect_one_sync()) {
        for (int d = ew; d < CS; d += N_EPI_WARPS)
          dsmem_bulk_copy(mapa_u32(my_blk + (b ? 0 : BB), (uint32_t)d), smem_s + b * SB, (uint32_t)SB,
                          mapa_u32(full0 + (b ? 0 : 8), (uint32_t)d));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(staged0 + 8 * b);
      if (warp == CL_EPI_WARP0 && lane == 0) CL_DBG(4);
      if (role == 2 && lane_ok) {
#pragma unroll
        for (int j = 0; j < KEEP; ++j) {
          const int n = row0 + KEEP * g + j, part = n / (NR / 2), utt = n % (NR / 2);
          if (utt < NBc) p.hseq1[((long long)(m * 2 + part) * R + (long long)(b0 + utt) * Tp + 1 + t) * H + unit] = hn[j];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == CL_EPI_WARP0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

static size_t cluster_smem(int H, int upc, int nr) { return (size_t)2 * H * 4 * nr + (size_t)2 * upc * 4 * nr + 256 + 1024; }
// rows of a step: 16 (2 input parts x 8 utterance slots) for <= 8 utterances, else chunks of 16 utterances = 32 rows
// (64 rows per chunk measured slower than the wavefront kernel: the exchange is bound by the DSMEM ingest of an SM)
static int cluster_rows(int NB) { return NB <= 8 ? 16 : 32; }
static int cluster_chunks(int NB) { const int ch = cluster_rows(NB) / 2; return (NB + ch - 1) / ch; }

// hidden units per CTA: a multiple of 8 (whole k-cores), <= 32 (M = 128 lanes); cluster of <= 16 CTAs; W hi + lo + the
// accumulator within the 512 TMEM columns.  `alt` selects the second choice (smaller CTAs, larger cluster).
static int cluster_upc(int H, int nr, int alt, int* cs_out) {
  if (H % 64 != 0 || H + 2 * nr > 512 || cluster_smem(H, 32, nr) > 232448) return 0;     // 227 KB: opt-in shared memory of an sm_100 SM
  const int cand[2] = {32, 24};
  int found = 0;
  for (int i = 0; i < 2; ++i) {
    const int upc = cand[i];
    if (H % upc != 0) continue;
    const int cs = H / upc;
    if (cs > 16 || cs < 1) continue;
    if (found++ < alt) continue;
    *cs_out = cs;
    return upc;
  }
  return 0;
}

template <int UPC, int EPW, int NR>
static int cluster_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int cs, size_t smem, int n_chunks, cudaStream_t st) {
  IDV_CUDA(cudaFuncSetAttribute(lstm_cluster_tc_kernel<UPC, EPW, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cs > 8) IDV_CUDA(cudaFuncSetAttribute(lstm_cluster_tc_kernel<UPC, EPW, NR>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(cs, 6 * n_chunks, 1); cfg.blockDim = dim3(32 * (CL_EPI_WARP0 + 4 * EPW)); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return IDV_OK;
}

// clusters of this shape that can be resident at once (0 + IDV_E_RESOURCE when the cluster cannot be scheduled at all)
template <int UPC, int EPW, int NR>
static int cluster_occupancy(int cs, size_t smem, int* max_clusters) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  int rc = cluster_cfg<UPC, EPW, NR>(cfg, attr, cs, smem, 1, nullptr);
  if (rc) return rc;
  *max_clusters = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(max_clusters, lstm_cluster_tc_kernel<UPC, EPW, NR>, &cfg);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("idv_lstm2_cluster_tc: cluster of %d CTAs cannot be scheduled: %s", cs, cudaGetErrorString(e));
    *max_clusters = 0;
    return IDV_E_RESOURCE;
  }
  return IDV_OK;
}

template <int UPC, int EPW, int NR>
static int launch_cluster(const ClusterParams& p, size_t smem, int n_chunks, cudaStream_t st) {
  // the six clusters of a chunk wait on one another (L0 -> IP -> L1): they should be resident at once (for speed; the
  // waits are acyclic, see ClusterParams).  No cooperative attribute (Nsight Compute cannot replay cooperative + cluster
  // launches): same precondition as the CTA-pair kernels (the process has the GPU to itself; every wait has a timeout that
  // traps)
  int max_clusters = 0;
  int rc = cluster_occupancy<UPC, EPW, NR>(p.CS, smem, &max_clusters);
  if (rc) return rc;
  if (max_clusters < 6) {
    set_error("idv_lstm2_cluster_tc: only %d of 6 clusters of %d CTAs are co-resident", max_clusters, p.CS);
    return IDV_E_RESOURCE;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  rc = cluster_cfg<UPC, EPW, NR>(cfg, attr, p.CS, smem, n_chunks, st);
  if (rc) return rc;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_cluster_tc_kernel<UPC, EPW, NR>, p);
  if (e != cudaSuccess) {
    set_error("idv_lstm2_cluster_tc: launch failed: %s", cudaGetErrorString(e));
    return IDV_E_CUDA;
  }
  return IDV_OK;
}

}  // namespace tc
}  // namespace idv

extern "C" int idv_lstm2_cluster_config(int H, int NB, int T, int* upc, int* cs, int64_t* work_bytes) {
  using namespace idv;
  IDV_CHECK_ARG(upc && cs && work_bytes, "idv_lstm2_cluster_config: null pointer");
  IDV_CHECK_ARG(NB >= 1, "idv_lstm2_cluster_config: empty batch");
  const int nr = tc::cluster_rows(NB), nch = tc::cluster_chunks(NB);
  IDV_CHECK_ARG(T >= 1, "idv_lstm2_cluster_config: T must be positive");
  int c = 0;
  const int u = tc::cluster_upc(H, nr, option_lstm_cluster_alt(), &c);
  IDV_CHECK_ARG(u > 0, "idv_lstm2_cluster_config: hidden size %d is not supported by the cluster recurrence", H);
  *upc = u;
  *cs = c;
  // per chunk: hx0 bf16 [T][2][H * 2 * nr] + g1x fp32 [T][2][cs][128][nr]
  *work_bytes = (int64_t)nch * ((int64_t)T * 2 * H * 4 * nr + (int64_t)T * 2 * c * 128 * nr * 4);
  return IDV_OK;
}

extern "C" int idv_lstm2_cluster_concurrency(int H, int NB, int* max_clusters) {
  using namespace idv;
  using namespace idv::tc;
  IDV_CHECK_ARG(max_clusters, "idv_lstm2_cluster_concurrency: null pointer");
  int upc = 0, cs = 0;
  int64_t work_bytes = 0;
  int rc = idv_lstm2_cluster_config(H, NB, 1, &upc, &cs, &work_bytes);
  if (rc) return rc;
  const int nr = cluster_rows(NB);
  const size_t smem = cluster_smem(H, upc, nr);
  if (nr == 16) return upc == 32 ? cluster_occupancy<32, 4, 16>(cs, smem, max_clusters) : cluster_occupancy<24, 4, 16>(cs, smem, max_clusters);
  return upc == 32 ? cluster_occupancy<32, 4, 32>(cs, smem, max_clusters) : cluster_occupancy<24, 4, 32>(cs, smem, max_clusters);
}

extern "C" int idv_lstm2_cluster_tc(const float* g0, int64_t g_m_off, int64_t g_p_off, int g_ld, const void* w_hh0,
                                    const void* w_ih1, const void* w_hh1, const float* bias1, int NB, int T, int H,
                                    float* hseq1, void* work, unsigned int* sync, int t_valid, void* stream) {
  using namespace idv;
  using namespace idv::tc;
  IDV_CHECK_ARG(g0 && w_hh0 && w_ih1 && w_hh1 && bias1 && hseq1 && work && sync, "idv_lstm2_cluster_tc: null pointer");
  // cluster launch without the cooperative guarantee: not when kernels of several streams share the GPU
  // ("gemm_dynamic_tiles") or the device is shared with other processes ("lstm_wave_cta_pairs" = 0, lib.check_exclusive_device)
  if (option_dynamic_tiles() || !option_lstm_wave_pairs()) {
    set_error("idv_lstm2_cluster_tc: the GPU is shared (gemm_dynamic_tiles / lstm_wave_cta_pairs): cluster recurrence off");
    return IDV_E_RESOURCE;
  }
  int upc = 0, cs = 0;
  int64_t work_bytes = 0;
  int rc = idv_lstm2_cluster_config(H, NB, T, &upc, &cs, &work_bytes);
  if (rc) return rc;
  int dev = 0, smem_optin = 0, sms = 0;
  IDV_CUDA(cudaGetDevice(&dev));
  IDV_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  IDV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int nr = cluster_rows(NB);
  const size_t smem = cluster_smem(H, upc, nr);
  if (smem > (size_t)smem_optin || 6 * cs > sms) {
    set_error("idv_lstm2_cluster_tc: the device cannot hold 6 clusters of %d CTAs with %zu bytes of shared memory", cs, smem);
    return IDV_E_RESOURCE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ClusterParams p;
  p.g0 = g0; p.g_m_off = g_m_off; p.g_p_off = g_p_off; p.g_ld = g_ld; p.bias1 = bias1;
  p.w[0] = reinterpret_cast<const unsigned short*>(w_hh0);
  p.w[1] = reinterpret_cast<const unsigned short*>(w_ih1);
  p.w[2] = reinterpret_cast<const unsigned short*>(w_hh1);
  p.NB = NB; p.T = T; p.Tsteps = (t_valid > 0 && t_valid < T) ? t_valid : T; p.H = H; p.CS = cs;
  p.hseq1 = hseq1;
  const int nch = cluster_chunks(NB);
  p.hx0 = reinterpret_cast<unsigned short*>(work);
  p.hx0_chunk = (long long)T * 2 * H * 2 * nr;
  p.g1x = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(work) + (size_t)nch * (size_t)T * 2 * H * 4 * nr);
  p.g1x_chunk = (long long)T * 2 * cs * 128 * nr;
  p.sync = sync;
  p.dbg = nullptr;
  const bool dbg = getenv("IDV_LSTM_DBG") != nullptr && p.Tsteps > 304;
  if (dbg) {
    IDV_CUDA(cudaMalloc(&p.dbg, 96 * sizeof(unsigned long long)));
    IDV_CUDA(cudaMemsetAsync(p.dbg, 0, 96 * sizeof(unsigned long long), st));
  }
  IDV_CUDA(cudaMemsetAsync(sync, 0, (size_t)nch * 4 * CL_SYNC_STRIDE * sizeof(unsigned int), st));
  if (nr == 16) rc = upc == 32 ? launch_cluster<32, 4, 16>(p, smem, nch, st) : launch_cluster<24, 4, 16>(p, smem, nch, st);
  else rc = upc == 32 ? launch_cluster<32, 4, 32>(p, smem, nch, st) : launch_cluster<24, 4, 32>(p, smem, nch, st);
  if (dbg && rc == IDV_OK) {
    unsigned long long h[96];
    IDV_CUDA(cudaStreamSynchronize(st));
    IDV_CUDA(cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    const char* names[3] = {"L0", "IP", "L1"};
    // slots: 0 operand landed, 1 MMAs issued, 2 accumulator seen, 6 TMEM drained, 7 gates done, 3 block staged, 4 pushed, 5 published
    const int order[8] = {0, 1, 2, 6, 7, 3, 4, 5};
    for (int ro = 0; ro < 3; ++ro)
      for (int i = 0; i < 4; ++i) {
        fprintf(stderr, "[cluster dbg] %s t=%d (cs %d):", names[ro], 300 + i, cs);
        for (int k = 0; k < 8; ++k) {
          const unsigned long long v = h[(ro * 4 + i) * 8 + order[k]];
          fprintf(stderr, " %lld", v ? (long long)(v - h[0]) : -1LL);
        }
        fprintf(stderr, "\n");
      }
  }
  if (p.dbg) cudaFree(p.dbg);
  return rc;
}
