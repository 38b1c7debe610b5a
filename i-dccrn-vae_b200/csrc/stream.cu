// Frame-streaming enhancement (causal network, hop-synchronous streams with carried state): the small state kernels
// around the tap-GEMMs.  One step consumes hop*k new samples per stream and produces k frames:
//   stream_frames_split : [history | new samples] -> split-bf16 STFT frames of the k new frames (reflect start)
//   stream_hist_shift   : history <- last (win - hop) samples of the window
//   lstm_cell_step      : one LSTM time step (gate math) on pre-computed input + recurrent projections
//   carry_rows          : last frame row -> causal pad row of every state plane set (conv / transposed conv x[t-1])
//   stream_ola          : overlap-add of the k new frames into the carried tail, emit hop*k finished samples
// The whole-utterance reference semantics are model/pvae_module.py:L21-27 (STFT), L38-42 (iSTFT),
// model/complex_progress.py:L16-22 / L244-250 (causal time padding) and nn.LSTM's carried (h, c).
#include "idv_common.cuh"

namespace idv {

__device__ __forceinline__ float sg_f(float x) { return 1.f / (1.f + expf(-x)); }

// frames[(b*k + f)][j] = window[b][hop*f + j] for j < win (window = [hist (win - hop) | x_new (hop*k)] holds the
// samples base .. base + win - hop + hop*k - 1 of the stream); a negative global index g reads sample -g (the
// reference's reflect padding at the start of the signal), samples not received yet read as 0 (their window tap is 0).
__global__ void __launch_bounds__(256) stream_frames_split_kernel(const float* __restrict__ hist,
                                                                  const float* __restrict__ xnew, int NB, int k,
                                                                  long long base, int hop, int win, int kpad,
                                                                  unsigned short* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int hl_len = win - hop, wlen = hl_len + hop * k;
  const long long n = (long long)NB * k * (kpad / 4);
  const long long hl = (long long)NB * k * kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j4 = (int)(i % (kpad / 4)) * 4;
    const long long bf = i / (kpad / 4);
    const int f = (int)(bf % k), b = (int)(bf / k);
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j4 + e;
      float x = 0.f;
      if (j < win) {
        long long w = (long long)hop * f + j;               // index inside the window
        const long long g = base + w;                        // global sample index
        if (g < 0) w = -g - base;                            // reflect: sample -g
        if (w >= 0 && w < wlen)
          x = w < hl_len ? __ldg(hist + (long long)b * hl_len + w) : __ldg(xnew + (long long)b * hop * k + (w - hl_len));
      }
      v[e] = x;
    }
    st_split4(out, hl, bf * kpad + j4, make_float4(v[0], v[1], v[2], v[3]));
  }
}

// one block per stream: hist <- window[hop*k .. hop*k + win - hop)
__device__ __forceinline__ void hist_shift_stream(float* __restrict__ hist, const float* __restrict__ xnew, int b, int k,
                                                  int hop, int win, float* tmp) {
  const int hl_len = win - hop;
  for (int i = threadIdx.x; i < hl_len; i += blockDim.x) {
    const int w = hop * k + i;
    tmp[i] = w < hl_len ? hist[(long long)b * hl_len + w] : __ldg(xnew + (long long)b * hop * k + (w - hl_len));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hl_len; i += blockDim.x) hist[(long long)b * hl_len + i] = tmp[i];
}
__global__ void stream_hist_shift_kernel(float* __restrict__ hist, const float* __restrict__ xnew, int k, int hop,
                                         int win) {
  extern __shared__ float tmp[];
  hist_shift_stream(hist, xnew, blockIdx.x, k, hop, win, tmp);
}

// gates = g_in[stream row (b, frame)] + g_rec[stream][b]; c, h update in place; h also to the frame's hseq row
__global__ void __launch_bounds__(256) lstm_cell_step_kernel(const float* __restrict__ g_in, long long g_m_off,
                                                             long long g_p_off, int g_ld,
                                                             const float* __restrict__ g_rec, int NB, int H, int Tp,
                                                             int frame, float* __restrict__ c,
                                                             unsigned short* __restrict__ h_split,
                                                             float* __restrict__ hseq) {
  pdl_trigger();
  pdl_wait();
  const long long n = 4LL * NB * H;
  const long long hl = n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % H);
    const int b = (int)((i / H) % NB);
    const int s = (int)(i / ((long long)H * NB));            // stream = m*2 + p
    const float* gr = g_rec + ((long long)s * NB + b) * 4 * H;
    float a[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = __ldg(gr + q * H + j);
    if (g_in) {
      const float* gi = g_in + (s >> 1) * g_m_off + (s & 1) * g_p_off + ((long long)b * Tp + 1 + frame) * g_ld;
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] += __ldg(gi + q * H + j);
    }
    const float cn = sg_f(a[1]) * c[i] + sg_f(a[0]) * tanhf(a[2]);
    const float hn = sg_f(a[3]) * tanhf(cn);
    c[i] = cn;
    st_split1(h_split, hl, i, hn);
    if (hseq) hseq[((long long)s * NB * Tp + (long long)b * Tp + 1 + frame) * H + j] = hn;
  }
}

struct CarryEntry {
  unsigned long long base;      // device address of the first plane
  long long n_planes;           // planes (both bf16 halves count separately)
  long long plane_bytes;        // R * row_bytes
  int row_bytes;                // multiple of 16
  int NB, Tp, src_row;          // row b*Tp + src_row -> row b*Tp
};

// grid (chunks, entries): a chunk walks over (plane, stream) rows, the threads of a block over the 16-byte vectors of
// 256/vec rows at a time; four independent row copies in flight per thread
__device__ __forceinline__ void carry_rows_entry(const CarryEntry& e) {
  const int vec = e.row_bytes / 16;                        // vectors per row
  const int rows_per_pass = 256 / vec > 0 ? 256 / vec : 1;
  const int v = threadIdx.x % vec, rsub = threadIdx.x / vec;
  const long long n_rows = e.n_planes * e.NB;
  const long long stride = (long long)gridDim.x * rows_per_pass;
  const long long src_off = (long long)e.src_row * e.row_bytes + v * 16;
  char* base = reinterpret_cast<char*>(e.base);
  if (vec <= 256 && rsub < rows_per_pass) {
    long long r = (long long)blockIdx.x * rows_per_pass + rsub;
    for (; r + 3 * stride < n_rows; r += 4 * stride) {
      char* p[4];
      uint4 val[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + u * stride;
        p[u] = base + (rr / e.NB) * e.plane_bytes + (rr % e.NB) * (long long)e.Tp * e.row_bytes;
        val[u] = *reinterpret_cast<const uint4*>(p[u] + src_off);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(p[u] + v * 16) = val[u];
    }
    for (; r < n_rows; r += stride) {
      char* p = base + (r / e.NB) * e.plane_bytes + (r % e.NB) * (long long)e.Tp * e.row_bytes;
      *reinterpret_cast<uint4*>(p + v * 16) = *reinterpret_cast<const uint4*>(p + src_off);
    }
  }
}
__global__ void __launch_bounds__(256) carry_rows_kernel(const CarryEntry* __restrict__ table,
                                                         unsigned long long* __restrict__ counter) {
  carry_rows_entry(table[blockIdx.y]);
  if (counter && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *counter += 1ULL;
}

// acc: carried tail (NB, win - hop) of partial sums for the outputs o0 + hop*k .. ; frames (NB*k, frame_ld).
// Position i of this step is output sample o = o0 + i, o0 = hop*t0 - (win/2) (t0 = global index of the first new
// frame): the frames t0 + f cover i - hop*f in [0, win).  Emits i < hop*k, keeps the rest as the new tail.
__device__ __forceinline__ void ola_stream(const float* __restrict__ frames, int frame_ld, const float* __restrict__ wsq,
                                           float* __restrict__ acc, int b, int k, long long t0, int hop, int win,
                                           float* __restrict__ out, float* tail) {
  const int tl = win - hop, total = hop * k + tl;
  const float* fb = frames + (long long)b * k * frame_ld;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    float a = i < tl ? acc[(long long)b * tl + i] : 0.f;
    int f_hi = i / hop;
    if (f_hi > k - 1) f_hi = k - 1;
    int f_lo = (i - win + hop) / hop;
    if (i - win + 1 <= 0) f_lo = 0;
    for (int f = f_lo; f <= f_hi; ++f) {
      const int j = i - hop * f;
      if (j >= 0 && j < win) a += __ldg(fb + (long long)f * frame_ld + j);
    }
    if (i < hop * k) {
      // window envelope of the finished sample: every frame t >= 0 covering it (a stream has no last frame)
      const long long q = hop * t0 + i;                      // o + win/2: position relative to frame 0's first tap
      float env = 0.f;
      long long t_hi = q / hop, t_lo = (q - win + hop) / hop;
      if (q - win + 1 <= 0) t_lo = 0;
      for (long long t = t_lo; t <= t_hi; ++t) {
        const long long j = q - hop * t;
        if (j >= 0 && j < win) env += __ldg(wsq + j);
      }
      out[(long long)b * hop * k + i] = env > 0.f ? a / env : 0.f;   // env == 0 only at o = -win/2 (pre-roll)
    } else {
      tail[i - hop * k] = a;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tl; i += blockDim.x) acc[(long long)b * tl + i] = tail[i];
}
__global__ void __launch_bounds__(256) stream_ola_kernel(const float* __restrict__ frames, int frame_ld,
                                                         const float* __restrict__ wsq, float* __restrict__ acc,
                                                         int k, long long t0, int hop, int win,
                                                         float* __restrict__ out) {
  extern __shared__ float tail[];
  ola_stream(frames, frame_ld, wsq, acc, blockIdx.x, k, t0, hop, win, out, tail);
}

// Everything of a step that only updates carried state or emits the output, in ONE launch (the step is a chain of
// dependent small kernels: every launch costs its latency): blockIdx.y < n_entries: carry_rows of that entry;
// n_entries: history shift; n_entries + 1: last STFT frame -> prev; n_entries + 2: overlap-add + output.
struct TailParams {
  const CarryEntry* table;
  int n_entries;
  unsigned long long* counter;
  float* hist; const float* xnew;
  const float* stft; int F; float* prev;
  const float* frames; int frame_ld; const float* wsq; float* acc; long long t0; float* out;
  int NB, k, hop, win;
};
__global__ void __launch_bounds__(256) stream_tail_kernel(const TailParams p) {
  extern __shared__ float tmp[];
  pdl_trigger();
  pdl_wait();
  const int job = (int)blockIdx.y - p.n_entries;
  if (job < 0) {
    carry_rows_entry(p.table[blockIdx.y]);
    if (p.counter && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.counter += 1ULL;
  } else if (job == 0) {
    if (p.hist)
      for (int b = blockIdx.x; b < p.NB; b += gridDim.x) {
        hist_shift_stream(p.hist, p.xnew, b, p.k, p.hop, p.win, tmp);
        __syncthreads();
      }
  } else if (job == 1) {
    if (p.prev) {
      const int n = p.NB * p.F;
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        *reinterpret_cast<float2*>(p.prev + (long long)i * 2) =
            __ldg(reinterpret_cast<const float2*>(p.stft + ((long long)i * p.k + (p.k - 1)) * 2));
    }
  } else {
    for (int b = blockIdx.x; b < p.NB; b += gridDim.x) {
      ola_stream(p.frames, p.frame_ld, p.wsq, p.acc, b, p.k, p.t0, p.hop, p.win, p.out, tmp);
      __syncthreads();
    }
  }
}

}  // namespace idv

extern "C" int idv_stream_tail(const idv_carry_t* table, int n_entries, uint64_t* counter, float* hist, const float* x_new,
                               const float* stft, int F, float* prev, const float* frames, int frame_ld, const float* wsq,
                               float* acc, int64_t t0, float* out, int NB, int k, int hop, int win, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(table && n_entries > 0 && n_entries <= 65000, "idv_stream_tail: bad carry table");
  IDV_CHECK_ARG(frames && wsq && acc && out && NB > 0 && k > 0 && t0 >= 0 && hop > 0 && win > hop && frame_ld >= win &&
                    win - hop <= 8192,
                "idv_stream_tail: bad argument");
  IDV_CHECK_ARG((hist == nullptr) == (x_new == nullptr) && (prev == nullptr || (stft && F > 0)),
                "idv_stream_tail: hist needs x_new, prev needs stft");
  TailParams p;
  p.table = reinterpret_cast<const CarryEntry*>(table); p.n_entries = n_entries;
  p.counter = reinterpret_cast<unsigned long long*>(counter);
  p.hist = hist; p.xnew = x_new; p.stft = stft; p.F = F; p.prev = prev;
  p.frames = frames; p.frame_ld = frame_ld; p.wsq = wsq; p.acc = acc; p.t0 = (long long)t0; p.out = out;
  p.NB = NB; p.k = k; p.hop = hop; p.win = win;
  dim3 grid(48, n_entries + 3);
  IDV_CUDA(launch_pdl(stream_tail_kernel, grid, dim3(256), (size_t)(win - hop) * sizeof(float), (cudaStream_t)stream, p));
  return IDV_OK;
}

extern "C" int idv_stream_frames_split(const float* hist, const float* x_new, int NB, int k, int64_t base, int hop,
                                       int win, int kpad, void* frames, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(hist && x_new && frames && NB > 0 && k > 0 && hop > 0 && win > hop && kpad >= win && kpad % 64 == 0,
                "idv_stream_frames_split: bad argument");
  const long long n = (long long)NB * k * (kpad / 4);
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  IDV_CUDA(launch_pdl(stream_frames_split_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, hist, x_new, NB, k,
                      (long long)base, hop, win, kpad, reinterpret_cast<unsigned short*>(frames)));
  return IDV_OK;
}

extern "C" int idv_stream_hist_shift(float* hist, const float* x_new, int NB, int k, int hop, int win, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(hist && x_new && NB > 0 && k > 0 && hop > 0 && win > hop && win - hop <= 8192,
                "idv_stream_hist_shift: bad argument");
  stream_hist_shift_kernel<<<NB, 128, (size_t)(win - hop) * sizeof(float), (cudaStream_t)stream>>>(hist, x_new, k, hop,
                                                                                                    win);
  IDV_LAUNCH_CHECK("stream_hist_shift_kernel");
  return IDV_OK;
}

extern "C" int idv_lstm_cell_step(const float* g_in, int64_t g_m_off, int64_t g_p_off, int g_ld, const float* g_rec,
                                  int NB, int H, int T, int frame, float* c, void* h_split, float* hseq,
                                  void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(g_rec && c && h_split && NB > 0 && H > 0 && T > 0 && frame >= 0 && frame < T,
                "idv_lstm_cell_step: bad argument");
  const long long n = 4LL * NB * H;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  IDV_CUDA(launch_pdl(lstm_cell_step_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, g_in, (long long)g_m_off,
                      (long long)g_p_off, g_ld, g_rec, NB, H, T + 1, frame, c, reinterpret_cast<unsigned short*>(h_split),
                      hseq));
  return IDV_OK;
}

extern "C" int idv_carry_rows(const idv_carry_t* table, int n_entries, uint64_t* counter, void* stream) {
  using namespace idv;
  static_assert(sizeof(idv_carry_t) == sizeof(CarryEntry), "idv_carry_t layout");
  IDV_CHECK_ARG(table && n_entries > 0 && n_entries <= 65535, "idv_carry_rows: bad argument");
  dim3 grid(48, n_entries);        // row_bytes <= 4096 (vec <= 256) is checked by the caller building the table
  carry_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const CarryEntry*>(table),
                                                            reinterpret_cast<unsigned long long*>(counter));
  IDV_LAUNCH_CHECK("carry_rows_kernel");
  return IDV_OK;
}

extern "C" int idv_stream_ola(const float* frames, int frame_ld, const float* wsq, float* acc, int NB, int k,
                              int64_t t0, int hop, int win, float* out, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(frames && wsq && acc && out && NB > 0 && k > 0 && t0 >= 0 && hop > 0 && win > hop && frame_ld >= win &&
                    win - hop <= 8192,
                "idv_stream_ola: bad argument");
  stream_ola_kernel<<<NB, 256, (size_t)(win - hop) * sizeof(float), (cudaStream_t)stream>>>(
      frames, frame_ld, wsq, acc, k, (long long)t0, hop, win, out);
  IDV_LAUNCH_CHECK("stream_ola_kernel");
  return IDV_OK;
}

namespace idv {
// prev[b][f] = stft[b][f][k-1]  (user-layout STFT chunk (NB, F, k, 2) -> (NB, F, 2))
__global__ void __launch_bounds__(256) stream_last_frame_kernel(const float* __restrict__ stft, int n, int k,
                                                                float* __restrict__ prev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    *reinterpret_cast<float2*>(prev + (long long)i * 2) =
        __ldg(reinterpret_cast<const float2*>(stft + ((long long)i * k + (k - 1)) * 2));
}
}  // namespace idv

extern "C" int idv_stream_last_frame(const float* stft, int NB, int F, int k, float* prev, void* stream) {
  using namespace idv;
  IDV_CHECK_ARG(stft && prev && NB > 0 && F > 0 && k > 0, "idv_stream_last_frame: bad argument");
  const int n = NB * F;
  stream_last_frame_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(stft, n, k, prev);
  IDV_LAUNCH_CHECK("stream_last_frame_kernel");
  return IDV_OK;
}
