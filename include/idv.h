/*
 * idv.h — C ABI of libidv_b200.so: the B200 (sm_100a) kernels behind the I-DCCRN-VAE enhancement
 * forward path.  The reference (iris1997jiatong/I-DCCRN-VAE) is pure PyTorch and has no FFI; each
 * entry point below names the reference call site (file:line under /root/reference) whose library
 * kernels it replaces.  See INTEGRATION.md for the ctypes binding the Python wrappers use.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named h_*; nothing is allocated internally;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 = ok, otherwise an IDV_E_* code; idv_last_error() gives the message of the
 *     last failure on the calling thread.  Nothing throws across the boundary;
 *   - all floating point is IEEE fp32.
 *
 * Internal activation layout ("planes"): fp32 [F][R][Cp], or split bf16 [2][F][R][Cp] (see
 * idv_tapgemm_tc), with
 *     R  = NB * Tp,  Tp = T + 1,  row(b, t) = b*Tp + 1 + t,  row(b, -1) = b*Tp is an all-zero
 *          causal pad row (the reference's left time padding, model/complex_progress.py:L16-22);
 *     Cp = 2*Ch, Ch = C rounded up to 8: real part of complex channel c at c, imaginary at Ch + c.
 * "user" layout is the reference's (B, C, F, T, 2) fp32 with re/im interleaved innermost.
 *
 * Valid frames (`t_valid`): the non-causal network (model/net_config.py, ComplexConv2d /
 * ComplexConvTranspose2d at model/complex_progress.py:L24-36, L253-279) changes the number of frames per
 * layer (T-1 per encoder layer, T+1 per decoder layer).  The row layout stays that of the allocation (`T`);
 * every entry point that walks rows takes `t_valid` = number of valid frames per utterance of the tensor it
 * WRITES (0 or >= T: all T).  Rows of frames t >= t_valid are written as zero exactly like the pad row, so
 * a later layer reading x[t+1] / x[t] beyond the valid range sees the zero padding of the reference.
 * User-layout tensors have exactly t_valid frames.
 */
#ifndef IDV_H_
#define IDV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDV_OK 0
#define IDV_E_ARG 1      /* bad argument / unsupported shape */
#define IDV_E_CUDA 2     /* CUDA runtime error (message has the cudaError string) */
#define IDV_E_RESOURCE 3 /* kernel cannot be made resident (cooperative launch too large) */

#define IDV_ABI_VERSION 8

int idv_abi_version(void);
const char* idv_last_error(void);
/* Process-wide tuning options.  "lstm_ncols": gate columns per CTA of idv_lstm_recurrent_tc (0 = auto; 64 = half as
 * many CTAs, leaves SMs free for kernels running concurrently on other streams).  "gemm_dynamic_tiles": 1 = the
 * tensor-core tap-GEMM claims its tiles from a global counter (for kernels sharing the GPU across streams),
 * 0 (default) = static round-robin tiles.  "gemm_cta_pairs": 1 (default) = tiles of width 256 run as CTA pairs
 * (thread-block clusters of 2, tcgen05 cta_group::2: M = 256 MMAs, every CTA stages half of the weight tile) when the
 * tiles are static; 0 = one CTA per tile everywhere.  "lstm_wave_cta_pairs": 1 (default) = neighbouring CTAs of
 * idv_lstm2_wave_tc run as pairs (each streams 64 of the 128 rows of h), 0 = every CTA streams all 128 rows.
 * "lstm_chunk_sync": 1 (default) = idv_lstm_layer_pair_tc publishes / polls one step counter per 64-wide K chunk of h (a
 * consumer loads chunk k as soon as its producers have published) instead of one per module, 0 = one counter per module.
 * "lstm_tma_publish": how the wavefront / CTA-pair LSTM kernels write h(t): -1 (default) = auto (staged 16-byte stores for the
 * 48-column one-layer kernel of H = 768, direct stores otherwise), 0 = direct stores, 2 = staged 16-byte stores, 1 = one
 * tensor store per CTA and step (measured slower, kept as a switch).
 * "lstm_cluster_alt": 0 (default) = idv_lstm2_cluster_tc uses the largest CTAs (32 hidden units: H = 384 as clusters of
 * 12), 1 = the second choice (24 units: clusters of 16), for devices on which the first cannot be scheduled.          */
int idv_set_option(const char* name, int value);
/* SM count of the current device (grids are sized against it). */
int idv_device_sm_count(int* out);

/* ---- tap-GEMM descriptors -------------------------------------------------------------------
 * One launch computes, for every unit u, output rows r in [0,R) and columns n in [0,N):
 *   out[u.out_f][r][u.out_ch_off + n] =
 *       act( bias[u.bias_off + n] + sum_{tap in u} sum_{c < tap.kc}
 *            A_{tap.src}[tap.f_in][r - tap.dt][tap.ch_off + c] * W[tap.w_off + c*N + n] )
 * rows outside [0, R) read as zero; if Tp > 0, output rows with r % Tp == 0 are written as 0; if Tp < 0 the
 * row period is -Tp and the pad rows are NOT written (streaming: they carry x[t-1] of the previous step, see
 * idv_carry_rows).
 * act = PReLU(slope) if apply_prelu else identity.  A complex (transposed) convolution, the LSTM
 * input projection and the complex dense layer are all instances (see pack.py).            */
typedef struct {
  int32_t src, f_in, dt, ch_off, kc, w_off;
} idv_tap_t;
typedef struct {
  int32_t tap_begin, n_taps, out_f, out_ch_off, bias_off, reserved;
} idv_unit_t;

/* replaces nn.Conv2d x4 + sub/add/slice/stack + ComplexBatchNormal(eval) + nn.PReLU
 * (model/complex_progress.py:L16-22, L161-209; model/pvae_module.py:L64-68), nn.ConvTranspose2d x4
 * + torch.cat skip (complex_progress.py:L244-250; pvae_module.py:L2092-2099, L2556-2568), the
 * nn.LSTM input projections (complex_progress.py:L58-61) and ComplexDense (L83-89).              */
int idv_tapgemm_f32(const float* a0, int a0_ld, int64_t a0_plane,
                    const float* a1, int a1_ld, int64_t a1_plane,
                    int R, int Tp,
                    const float* w, const float* bias, int N,
                    const idv_unit_t* units, const idv_tap_t* taps, int n_units,
                    float* out, int out_ld, int64_t out_plane,
                    int apply_prelu, float prelu_slope, int t_valid, void* stream);

/* Tensor-core version (tcgen05.mma kind::f16, TMEM accumulators, TMA-fed): same contract on the
 * "split" activation format: every fp32 value x is stored as two bf16 planes hi = bf16(x),
 * lo = bf16(x - hi), tensor = bf16 [2 (hi,lo)][planes][R][cp]; products are evaluated as
 * a_hi*w_hi + a_hi*w_lo + a_lo*w_hi with fp32 accumulation.
 * wt: bf16 [2][n_slots][N][kc_max] (K-major), tap.w_off = slot index, tap.kc % 64 == 0,
 * unit.reserved = number of 64-wide K steps of the unit.  out: split bf16 (out_hl = elements between
 * the hi and lo sets) when out_split, else fp32.  N = 32 or a multiple of 64.  N > out_ld (out_ld % 32 == 0,
 * N % out_ld == 0, unit.out_ch_off = 0): the columns of a unit wrap into CONSECUTIVE output planes, column n goes to
 * plane unit.out_f + n / out_ld, column n % out_ld - two output planes of a narrow transposed conv as one tile over
 * the input planes they share; planes at or past out_hl / out_plane (out_hl > 0) are not written.            */
int idv_tapgemm_tc(const void* a0, int a0_cp, int a0_planes, const void* a1, int a1_cp, int a1_planes,
                   int R, int Tp, const void* wt, int kc_max, int n_slots, const float* bias, int N,
                   const idv_unit_t* units, const idv_tap_t* taps, int n_units, void* out, int out_ld,
                   int64_t out_plane, int64_t out_hl, int out_split, int apply_prelu, float prelu_slope,
                   int t_valid, void* stream);

/* Same kernel with the reconstruction head fused into the epilogue (last decoder layer, Cout = 1):
 * N == 32, a unit produces unit.out_ch_off (1..16) consecutive output bins starting at bin unit.out_f, bin e in the
 * accumulator columns (2e, 2e+1) = (re, im) - 8 bins per unit read 6 input planes where 2 bins read 3: the layer is
 * bound by its activation reads; bias[unit.bias_off + {0,1}] = (re, im); PReLU(slope) is always applied; head = 1 writes the complex
 * spectrum, head = 2 the mask head of model/pvae_module.py:L2594-2609 (needs stft_x (NB, head_fout, T, 2)).
 * predict: (NBtot, head_fout, T, 2), utterance index bo = b*head_bmul + head_boff.  `out` (optional, may be NULL): the
 * same spectrum as split-bf16 K-major rows bf16 [2][NBtot*T][out_ld] (row bo*T + t, columns 2*bin + part; hi / lo sets
 * out_hl elements apart) = the operand of the iSTFT's DFT GEMM (idv_spec_rows_split's layout; columns >= 2*head_fout
 * are not written: the caller keeps them zero).
 * Replaces idv_dec5_head_fwd on split-bf16 planes.                                                        */
int idv_tapgemm_tc_head(const void* a0, int a0_cp, int a0_planes, const void* a1, int a1_cp, int a1_planes,
                        int R, int Tp, const void* wt, int kc_max, int n_slots, const float* bias, int N,
                        const idv_unit_t* units, const idv_tap_t* taps, int n_units, void* out, int out_ld,
                        int64_t out_plane, int64_t out_hl, int out_split, int apply_prelu, float prelu_slope,
                        int head, int head_fout, int head_bmul, int head_boff, const float* stft_x,
                        float* predict, int t_valid, void* stream);
/* idv_tapgemm_tc with a second bias vector for the FIRST frame of every utterance (rows r with r % Tp == 1, same
 * per-unit offsets as `bias`): a layer composed at pack time with the affine map in front of it - ComplexDense followed
 * by the first causal transposed conv (model/pvae_module.py:L2085-2099: dense -> reshape -> decoders[0]) as ONE tap-GEMM
 * on the z planes with K = 2*zdim per time tap instead of 2*C per (plane, time tap) - has a different constant there:
 * the time tap that reads the zero pad row x[-1] does not see the dense layer's bias.                            */
int idv_tapgemm_tc_b2(const void* a0, int a0_cp, int a0_planes, const void* a1, int a1_cp, int a1_planes, int R, int Tp,
                      const void* wt, int kc_max, int n_slots, const float* bias, const float* bias_first, int N,
                      const idv_unit_t* units, const idv_tap_t* taps, int n_units, void* out, int out_ld,
                      int64_t out_plane, int64_t out_hl, int out_split, int apply_prelu, float prelu_slope, int t_valid,
                      void* stream);
/* idv_tapgemm_tc_b2 (bias_first may be NULL) for SMALL problems, with split-K: min_ksteps = the smallest unit.reserved
 * of the unit table (the caller packed it; 0 = never split).  When the output tiles would leave more than half of the
 * SMs idle (a frame-streaming step: 10-40 tiles, each streaming megabytes of weights and activation rows through ONE
 * SM's L2 port), the 64-wide K steps of every tile are divided over up to min_ksteps CTAs; partial accumulators are added
 * in an fp32 workspace owned by the library (red.global.add), the CTA that arrives last applies bias / PReLU and stores.
 * Same results up to fp32 summation order.  Not with gemm_dynamic_tiles (several streams).                       */
int idv_tapgemm_tc_splitk(const void* a0, int a0_cp, int a0_planes, const void* a1, int a1_cp, int a1_planes, int R,
                          int Tp, const void* wt, int kc_max, int n_slots, const float* bias, const float* bias_first,
                          int N, const idv_unit_t* units, const idv_tap_t* taps, int n_units, void* out, int out_ld,
                          int64_t out_plane, int64_t out_hl, int out_split, int apply_prelu, float prelu_slope,
                          int t_valid, int min_ksteps, void* stream);

/* ---- STFT / iSTFT -----------------------------------------------------------------------------
 * replaces torch.stft at model/pvae_module.py:L22 (n_fft 512, hop, win, periodic Hann, center,
 * reflect pad, onesided).  basis: [win][2*(n_fft/2+1)] fp32, column 2k = cos*w, 2k+1 = -sin*w
 * (built by pack.py).  x: (B, L) -> out: (B, n_fft/2+1, T, 2), T = L/hop + 1.                    */
int idv_stft_fwd(const float* x, int B, int L, const float* basis, int n_fft, int hop, int win,
                 float* out, void* stream);
/* replaces torch.istft at model/pvae_module.py:L41.  spec: (B, n_fft/2+1, T, 2) fp32;
 * basis: [2*(n_fft/2+1)][win] (synthesis window, 1/n_fft and c_k folded); wsq: [win] = w^2;
 * frames: workspace (B*T, win) fp32; out: (B, hop*(T-1)).                                        */
int idv_istft_fwd(const float* spec, int B, int T, const float* basis, const float* wsq,
                  int n_fft, int hop, int win, float* frames, float* out, void* stream);

/* Tensor-core STFT / iSTFT = operand preparation + idv_tapgemm_tc_head (epilogue mode 3 for the STFT) + overlap-add:
 *   idv_stft_frames_split: x (B, L) -> split-bf16 frames [2][B*T][kpad], frames[(b,t)][j] = reflect-padded signal
 *                          at hop*t + (n_fft-win)/2 + j for j < win, 0 up to kpad (kpad % 64 == 0);
 *   idv_tapgemm_tc_head(head = 3): D = frames . basis^T, column pair (2k, 2k+1) of row (b,t) is written to
 *                          predict[(b*head_fout + k)*Tp + t] (Tp = frames per utterance; no bias, no pad rows); with
 *                          `out` != NULL the same values also go to split-bf16 activation rows bf16 [2][B*(Tp+1)][out_ld]
 *                          (row b*(Tp+1)+1+t, column head_boff + 2k + part, hi / lo sets out_hl apart; everything the
 *                          kernel does not write - pad rows, padding columns - is kept zero by the caller): the input
 *                          of the first encoder layer run as a tap-GEMM (pack.pack_enc0_tc);
 *   idv_spec_rows_split:   spec (B, nbins, T, 2) -> split-bf16 rows [2][B*T][kpad], rows[(b,t)][2k+part];
 *   idv_ola_fwd:           frames (B*T, frame_ld) -> overlap-add / window envelope / centre trim -> (B, hop*(T-1)).
 * Ragged batches: lengths (device int32 [B], NULL = all L) gives the true sample count of every utterance of the
 * zero-padded batch: the framing reflects at each utterance's own end and zeroes the frames beyond its last one, the
 * overlap-add uses each utterance's own frame count for the envelope and zeroes the samples beyond hop*(T_b-1) - so a
 * padded batch of the causal network reproduces the per-utterance results exactly.                              */
int idv_stft_frames_split(const float* x, int B, int L, int n_fft, int hop, int win, int kpad, const int* lengths,
                          void* out, void* stream);
int idv_spec_rows_split(const float* spec, int B, int nbins, int T, int kpad, void* out, void* stream);
int idv_ola_fwd(const float* frames, int frame_ld, const float* wsq, int B, int T, int n_fft, int hop, int win,
                const int* lengths, float* out, void* stream);

/* ---- first encoder layer (Cin = 1) -------------------------------------------------------------
 * Encoder 0: ComplexConv2d(1 -> Cout, (5,2), stride (2,1), freq pad 2) + CBN(eval) + PReLU,
 * reading the user-layout STFT (B,257,T,2) and writing planes [Fout][R][2*Cout].
 * causal != 0: time pad 1 / last column dropped (time tap kt reads x[t-1+kt], T valid frames);
 * causal == 0: no time pad (tap kt reads x[t+kt]; pass t_valid = T-1).
 * w: [10 taps (kf*2+kt)][2 (re,im in)][2*Cout] with CBN folded, bias: [2*Cout].
 * Streaming: prev (B, Fin, 2) or NULL = the STFT frame before frame 0 (x[-1]); keep_pad != 0 leaves the pad rows
 * of `out` untouched (they carry this layer's previous output frame).                                 */
int idv_enc0_fwd(const float* stft, int B, int Fin, int T, const float* w, const float* bias,
                 int Cout, float prelu_slope, void* out, int out_split, int causal, int t_valid,
                 const float* prev, int keep_pad, void* stream);

/* ---- last decoder layer (Cout = 1) + reconstruction head ---------------------------------------
 * Decoder 5: causal ComplexConvTranspose2d(Cin -> 1) + CBN(eval) + PReLU (+ mask head,
 * model/pvae_module.py:L2594-2609 / L224-234) writing `predict` in user layout (NBdec, Fout, T, 2).
 * p: planes [Fin][R][p_cp]; skip: planes [Fin][R][s_cp] or NULL (zero skip).
 * w: [10 taps (kf*2+kt)][p_cp + s_cp][2], bias[2].  stft_x: (NBdec/?,Fout,T,2) noisy STFT rows
 * for this pass (mask head only, may be NULL when mask == 0); out_bstride lets a pass write
 * every S-th utterance of the (B*S) batch: out utterance index = b*out_bmul + out_boff.          */
int idv_dec5_head_fwd(const void* p, int p_cp, const void* skip, int s_cp, int in_split, int NB, int Fin,
                      int T, const float* w, const float* bias, float prelu_slope, int mask,
                      const float* stft_x, float* predict, int out_bmul, int out_boff, void* stream);

/* ---- complex LSTM ------------------------------------------------------------------------------
 * Recurrent part of one nn.LSTM layer for the four (module, input-part) streams of ComplexLSTM
 * (model/complex_progress.py:L58-74), gates i,f,g,o.  g: pre-computed input projections incl. both
 * biases; stream (m,p) starts at g + m*g_m_off + p*g_p_off, row stride g_ld.  whh: [2][4H][H]
 * (lstm_re, lstm_im).  hseq: [4 streams (m*2+p)][R][H]; the kernel zeroes the pad rows itself.
 * hsplit: optional (NULL) split-bf16 copy [2][4][R][H] of hseq for a tensor-core in-proj of the next
 * layer.  sync: 2 x uint32 workspace, zeroed by the call.  Cooperative launch.                     */
int idv_lstm_recurrent_fwd(const float* g, int64_t g_m_off, int64_t g_p_off, int g_ld,
                           const float* whh, int NB, int T, int H, float* hseq, void* hsplit,
                           unsigned int* sync, int t_valid, void* stream);
/* Tensor-core recurrence (tcgen05): same contract as idv_lstm_recurrent_fwd with the recurrent product
 * evaluated as h_hi*W_hi + h_hi*W_lo + h_lo*W_hi on the split-bf16 value of h(t-1) (fp32 accumulate, fp32
 * cell state).  idv_lstm_tc_config gives the gate columns per CTA (n_cols = 4*Hs) and CTAs per module for a
 * hidden size; wpack: bf16 [2 (hi,lo)][2 (module)][n_ctas][n_cols (gate*Hs + j)][H] holds row
 * W_hh^m[gate*H + c*Hs + j] at (c, gate*Hs + j).  hseq and/or hsplit may be NULL (not both).
 * hx: workspace bf16 [ceil(NB/64)][2][2][2][128][H]; sync: ceil(NB/64)*2 x uint32; both zeroed by the call.
 * H % 64 == 0; ceil(NB/64) * 2 * n_ctas must not exceed the SM count (cooperative launch).                  */
int idv_lstm_tc_config(int H, int* n_cols, int* n_ctas);
int idv_lstm_recurrent_tc(const float* g, int64_t g_m_off, int64_t g_p_off, int g_ld, const void* wpack,
                          int NB, int T, int H, float* hseq, void* hsplit, void* hx, unsigned int* sync,
                          int t_valid, void* stream);
/* Both layers of a 2-layer ComplexLSTM as one wavefront kernel (layer 0, the layer-1 input projection and layer 1
 * run concurrently, one time step apart): same arithmetic as two idv_lstm_recurrent_tc calls around a tensor-core
 * input projection, without materialising the layer-1 gate pre-activations.  g0: layer-0 input projection
 * (as for idv_lstm_recurrent_tc).  w_hh0 / w_ih1 / w_hh1: packs in the layout of idv_lstm_recurrent_tc's wpack
 * with (n_cols, n_ctas) from idv_lstm2_wave_config; bias1: fp32 [2][n_ctas][n_cols] = b_ih_l1 + b_hh_l1 in the
 * same CTA-major order.  hseq1: fp32 [4][R][H] layer-1 output.  work: work_bytes workspace, sync: 384 x uint32 (2 chunks x
 * 6 step counters, one 128-byte line each)
 * (both zeroed by the call).  Any NB: a launch interleaves up to two chunks of 64 utterances (two independent
 * recurrences in the same CTAs; option "lstm_interleave"), larger batches run as consecutive launches; 6 * n_ctas CTAs
 * must be co-resident.                                                                                                  */
int idv_lstm2_wave_config(int H, int* n_cols, int* n_ctas, int64_t* work_bytes);
int idv_lstm2_wave_tc(const float* g0, int64_t g_m_off, int64_t g_p_off, int g_ld, const void* w_hh0,
                      const void* w_ih1, const void* w_hh1, const float* bias1, int NB, int T, int H,
                      float* hseq1, void* work, unsigned int* sync, int t_valid, void* stream);
/* Cluster form of idv_lstm2_wave_tc: same arithmetic and contract, but every (module, role) is ONE thread-block cluster
 * (cs CTAs of upc hidden units each) with its weights resident in tensor memory that exchanges h(t) through distributed
 * shared memory instead of the L2 (csrc/lstm_cluster_tc.cu): 2.4 us instead of 7-8 us per dependent step for <= 8
 * utterances (a step works on 16 rows = 2 input parts x 8 utterance slots), 4.1 us for <= 16 (32 rows).  Larger batches run
 * as independent chunks of 16 utterances inside ONE launch, each chunk on its own six clusters (the chunks whose clusters
 * fit the device together run concurrently: 6 at H = 128, 2 at H = 384).  idv_lstm2_cluster_config gives (upc, cs,
 * work_bytes) for (H, NB, T) or fails when the hidden size is not supported (H = 384: 32 x 12 or, with option
 * "lstm_cluster_alt", 24 x 16; H = 128: 32 x 4; H + 64 <= 512 TMEM columns).
 * w_hh0 / w_ih1 / w_hh1: bf16 [2 (hi,lo)][2 (module)][cs][4*upc][H], CTA c holds row W[gate*H + c*upc + j] at 4*j + gate;
 * bias1: fp32 [2][cs][128] = b_ih_l1 + b_hh_l1 in the same order (entries >= 4*upc unused).  work: work_bytes, sync:
 * 128 * ceil(NB / 8) x uint32 (zeroed by the call).  Returns IDV_E_RESOURCE when the 6 clusters of a chunk cannot be
 * co-resident or the GPU is shared (options "gemm_dynamic_tiles", "lstm_wave_cta_pairs" = 0): the caller then uses
 * idv_lstm2_wave_tc.                                                                                                 */
int idv_lstm2_cluster_config(int H, int NB, int T, int* upc, int* cs, int64_t* work_bytes);
/* Clusters of the (H, NB) shape that the current device can hold at once (cudaOccupancyMaxActiveClusters): / 6 = chunks of
 * idv_lstm2_cluster_tc that run concurrently (clusters do not span GPCs: 5 chunks at H = 128, 1 at H = 384 on a B200). */
int idv_lstm2_cluster_concurrency(int H, int NB, int* max_clusters);
int idv_lstm2_cluster_tc(const float* g0, int64_t g_m_off, int64_t g_p_off, int g_ld, const void* w_hh0,
                         const void* w_ih1, const void* w_hh1, const float* bias1, int NB, int T, int H,
                         float* hseq1, void* work, unsigned int* sync, int t_valid, void* stream);
/* ONE nn.LSTM layer of both modules per launch as CTA pairs (the kernel of idv_lstm2_wave_tc restricted to its first
 * role): same contract as idv_lstm_recurrent_tc, wpack packed with (n_cols, n_ctas) from idv_lstm_layer_pair_config
 * (H = 768: 48 gate columns x 64 CTAs per module).  work: work_bytes workspace, sync: 384 x uint32 (both zeroed by the
 * call).  Any NB (chunks of 64 utterances run as consecutive launches).                                          */
int idv_lstm_layer_pair_config(int H, int* n_cols, int* n_ctas, int64_t* work_bytes);
int idv_lstm_layer_pair_tc(const float* g, int64_t g_m_off, int64_t g_p_off, int g_ld, const void* wpack, int NB, int T,
                           int H, float* hseq, void* hsplit, void* work, unsigned int* sync, int t_valid, void* stream);
/* Real nn.LSTM with ONE hidden unit and num_layers <= 4 layers (the GAN distinguisher, model/pvae_module.py:L2319-2345).
 * g: layer-0 gate pre-activations fp32 [R][g_ld] (columns 0-3 = i, f, g, o with both layer-0 biases), R = NB*(T+1),
 * row b*(T+1)+1+t.  wrec: fp32 [num_layers][12] = (w_ih[4], w_hh[4], bias[4]) (w_ih / bias unused for layer 0).
 * out: (NB, Tv, 1) = the top layer's h.  One thread per utterance (the recurrence is scalar).                  */
int idv_lstm_h1_fwd(const float* g, int g_ld, const float* wrec, int num_layers, int NB, int T, int t_valid, float* out,
                    void* stream);
/* Combine the four streams (real = rr - ii, imag = ir + ri), emit the user-layout latent
 * (NB, T, H, 2).  Replaces the stack/permute at complex_progress.py:L62-73, pvae_module.py:L2247. */
int idv_lstm_combine_fwd(const float* hseq, int NB, int T, int H, float* latent, int t_valid, void* stream);
/* reparameterization (model/pvae_module.py:L2177-2231).  latent: (NB,T,Htot,2); the (mu, log
 * sigma, delta) triplet starts at channel ch0 (zdim each).  eps_r/eps_i: (NB,S,T,zdim) or NULL ->
 * Philox4x32-10 N(0,1) from (seed, offset + *offset_dev); offset_dev (device, may be NULL) lets a captured CUDA
 * graph draw fresh noise on every replay.  z: (NB*S, T, zdim, 2).  variant 1 = the clamped formula of the
 * *_fc_latent encoders (model/pvae_module.py:L2403-2450).                                          */
int idv_reparam_fwd(const float* latent, int NB, int T, int Htot, int ch0, int zdim, int S,
                    const float* eps_r, const float* eps_i, uint64_t seed, uint64_t offset,
                    const uint64_t* offset_dev, int variant, float* z, void* stream);
/* Fused latent stage of the VAE encoders: the four LSTM streams hseq [4][R][H] (R = NB*(T+1)) ->
 *   latent  (NB, Tv, H, 2)            combine of complex_progress.py:L62-73 (re = rr - ii, im = ir + ri),
 *   z0 / z1 (NB*S, Tv, zdim, 2)       reparameterisation of latent k = 0 / 1 (model/pvae_module.py:L2177-2231;
 *                                     triplet k at channels [3k zdim, 3(k+1) zdim); z1 and its eps NULL when latent_num = 1),
 *   zplanes [S][1 plane][R][2*round8(zdim)]   z0 of sample s in the activation-plane layout the decoder's ComplexDense
 *                                     tap-GEMM reads (fp32, or bf16 hi/lo [S][2][R][Cp] when out_split; pad rows and
 *                                     rows of frames >= Tv written as zero) - replaces idv_lstm_combine_fwd +
 *                                     idv_reparam_fwd (x latent_num) + idv_z_to_planes (x S) by ONE launch.
 * eps_*: (NB, S, Tv, zdim) or all NULL -> Philox4x32-10 from (seed, offset + *offset_dev), one Philox block per
 * element and draw.  H == 3 * zdim * latent_num.  keep_pad (frame streaming): the pad rows of zplanes are left
 * untouched (they carry z of the previous step's last frame, refreshed by idv_carry_rows).                      */
int idv_latent_fwd(const float* hseq, int NB, int T, int H, int t_valid, int zdim, int latent_num, int S,
                   const float* eps_r0, const float* eps_i0, const float* eps_r1, const float* eps_i1,
                   uint64_t seed, uint64_t offset, const uint64_t* offset_dev, float* latent, float* z0, float* z1,
                   void* zplanes, int out_split, int keep_pad, void* stream);
/* idv_lstm_combine_fwd that ALSO writes the combined latent as one activation plane [1][R][2*round8(H)] (the input
 * of the supervised DCCRN's ComplexDense, model/pvae_module.py:L189-192): replaces idv_lstm_combine_fwd +
 * idv_z_to_planes.  planes: fp32, or bf16 hi/lo [2][R][Cp] when out_split; pad rows / rows >= Tv are zeroed. */
int idv_lstm_combine_planes(const float* hseq, int NB, int T, int H, int t_valid, float* latent, void* planes,
                            int out_split, void* stream);
/* Per-bin affine on a spectrum (B, F, T, 2): out = x * scale[f][part] + shift[f][part] (in place allowed) - the
 * data_mean / data_std normalisation of the CVAE encoders (model/pvae_module.py:L367-371: zero_edge_imag = 1 clears
 * the imaginary part of the first and last bin afterwards) and its inverse in the decoders (L483-484). */
int idv_bin_affine(const float* x, int B, int F, int T, const float* scale, const float* shift, int zero_edge_imag,
                   float* out, void* stream);

/* ---- layout conversion at the module boundary --------------------------------------------------*/
/* planes [F][R][Cp] (fp32, or split bf16 when in_split) -> user (NB, C, F, T, 2) */
int idv_planes_to_user(const void* planes, int in_split, int NB, int C, int F, int T, float* user,
                       int t_valid, void* stream);
/* user (NB, C, F, T, 2) -> planes (pad rows and pad channels written as zero) */
int idv_user_to_planes(const float* user, int NB, int C, int F, int T, void* planes, int out_split,
                       int t_valid, void* stream);
/* z (NB*S, t_valid, zdim, 2) sample s -> planes [1][NB*Tp][2*zdim] */
int idv_z_to_planes(const float* z, int NB, int S, int s, int T, int zdim, void* planes, int out_split,
                    int t_valid, void* stream);

/* stand-alone ComplexBatchNormal.forward(x, train=False) (model/complex_progress.py:L161-209) on the
 * reference layout x: (outer, C, inner, 2).  zb: [C][6] = Zrr, Zri, Zir, Zii, b'_r, b'_i with
 * b' = beta - Z mu.                                                                */
int idv_cbn_eval_user(const float* x, int64_t outer, int C, int64_t inner, const float* zb, float* out,
                      void* stream);

/* ---- ComplexBatchNormal(train=True), forward only (model/complex_progress.py:L131-160) -------------------------
 * idv_cbn_stats_planes: acc[c][5] (double) = sum r, sum i, sum r^2, sum i^2, sum r*i over every plane and every
 *   non-pad row of the planes tensor (fp32 or split bf16);
 * idv_cbn_train_finalize: batch mean / biased (co)variances (eps added to Vrr, Vii as the reference does), update
 *   of the running buffers (first != 0: copy, else EMA with `momentum`), zb[c][6] = Z, b' from the batch statistics;
 * idv_cbn_apply_planes: y <- act(Z y + b') in place (pad rows untouched), or into `out` (fp32 or split bf16 per
 *   out_split; pad rows, invalid frames and padding channels written as 0) when out != NULL - the training forward
 *   keeps the raw fp32 values for the backward pass and feeds the next layer's tensor-core GEMM the split copy.
 * idv_cbn_train_finalize also writes stats[c][5] = batch mean (re, im), Vrr, Vri, Vii (NULL = not wanted).           */
int idv_cbn_stats_planes(const void* planes, int split, int NB, int C, int F, int T, double* acc, int t_valid,
                         void* stream);
int idv_cbn_train_finalize(const double* acc, double count, int C, const float* gamma_rr, const float* gamma_ri,
                           const float* gamma_ii, const float* beta_r, const float* beta_i, float* run_mean_r,
                           float* run_mean_i, float* run_vrr, float* run_vri, float* run_vii, float momentum,
                           int first, float* zb, float* stats, void* stream);
int idv_cbn_apply_planes(void* planes, int split, int NB, int C, int F, int T, const float* zb, int apply_prelu,
                         float prelu_slope, int t_valid, void* out, int out_split, void* stream);
/* statistics on the reference layout x (outer, C, inner, 2) (stand-alone ComplexBatchNormal(train=True); the last
 * decoder layer whose raw output is written in the reference layout) and the in-place recon head on
 * y (n_utt, n_per_utt, 2): PReLU(slope) then, if mask, the mask head with stft_x[b / s_rep].                       */
int idv_cbn_stats_user(const float* x, int64_t outer, int C, int64_t inner, double* acc, void* stream);
int idv_head_user(float* y, int64_t n_per_utt, int64_t n_utt, float prelu_slope, int mask, const float* stft_x,
                  int s_rep, void* stream);

/* ---- backward of the NSVAE encoder (phase-1 training step, i_dccrn_vae/nsvae_dccrn/train_nsvae.py:L505-566) --------
 * GEMM-shaped gradients run on idv_tapgemm_tc: data gradients with transposed weights and mirrored taps, weight
 * gradients as GEMMs over the ROW dimension on the transposed copies made by idv_planes_transpose_split.
 *   idv_planes_transpose_split: planes [F][R][Cp] (fp32 / split) -> split bf16 [2][F][Cp][Rpad],
 *       out[f][c][k] = in[f][k + shift][c] (0 outside [0, R)), Rpad % 64 == 0;
 *   idv_f32_to_split: fp32 [n] -> split bf16 [2][n];
 *   idv_cbn_bwd_reduce / _finalize / _apply: ComplexBatchNormal(train=True) + PReLU backward
 *       (model/complex_progress.py:L131-209 differentiated through the batch mean and covariance; model/pvae_module.py:L64-68).
 *       y = raw conv output, g = gradient w.r.t. the layer output, stats / zb from idv_cbn_train_finalize;
 *       reduce -> acc[C][8] (double); finalize -> parameter gradients (+=, d_slope double accumulator) and
 *       coef[C][10]; apply -> dy = gradient w.r.t. y (planes, pad rows 0);
 *   idv_lstm_combine_bwd: gradient of the latent (NB, T, H, 2) -> dH [4][R][H] (complex_progress.py:L62-73);
 *   idv_lstm_scan_c: gate pre-activations P [4][R][4H] -> cell states [4][R][H];
 *   idv_lstm_cell_bwd_step: one BPTT step t (t = T-1 first with last = 1): dP[t] (fp32 planes and the split-bf16
 *       copy [2][4][NB][4H] for the dh = dP W_hh tap-GEMM of the next step), dc carried; dh_rec holds dh_parts
 *       partial planes [dh_parts][4][NB][H] (the recurrent tap-GEMM is split along K to fill the GPU) that are summed;
 *   idv_colsum_add: out[col] += sum_rows x (bias gradients);
 *   idv_enc0_wgrad: weight gradient of idv_enc0_fwd, dW [10][2][2*Cout];
 *   idv_adam_step: torch.optim.Adam(lr, betas, eps, weight_decay) on a flat buffer (train_nsvae.py:L200).           */
int idv_planes_transpose_split(const void* planes, int in_split, int F, int R, int Cp, int Rpad, int shift, void* out,
                               void* stream);
int idv_f32_to_split(const float* x, int64_t n, void* out, void* stream);
int idv_cbn_bwd_reduce(const void* y, int y_split, const void* g, int g_split, int NB, int C, int F, int T,
                       const float* stats, const float* zb, float slope, double* acc, int t_valid, void* stream);
int idv_cbn_bwd_finalize(const double* acc, double count, int C, const float* stats, const float* gamma_rr,
                         const float* gamma_ri, const float* gamma_ii, float* coef, float* d_gamma_rr,
                         float* d_gamma_ri, float* d_gamma_ii, float* d_beta_r, float* d_beta_i, double* d_slope,
                         void* stream);
int idv_cbn_bwd_apply(const void* y, int y_split, const void* g, int g_split, int NB, int C, int F, int T,
                      const float* stats, const float* zb, const float* coef, float slope, void* dy, int dy_split,
                      int t_valid, void* stream);
int idv_lstm_combine_bwd(const float* dlatent, int NB, int T, int H, float* dH, int t_valid, void* stream);
int idv_lstm_scan_c(const float* P, int NB, int T, int H, float* cst, int t_valid, void* stream);
int idv_lstm_cell_bwd_step(const float* P, const float* cst, const float* dH, const float* dh_rec, float* dc, int NB,
                           int T, int H, int t, int last, int dh_parts, float* dP, void* dP_step, void* stream);
int idv_colsum_add(const float* x, int64_t rows, int cols, int ld, float* out, void* stream);
int idv_enc0_wgrad(const float* stft, const float* dY, int B, int Fin, int T, int Cout, int causal, float* dW,
                   void* stream);
int idv_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, void* stream);
/* Closed-form KL(q1 || q2) between complex Gaussians, value and gradient in one pass (standard_nsvae_loss_true_kl.cal_kl,
 * model/nsvae_loss.py:L275-328).  lat1 (n_bt, H1, 2) holds (mu, log sigma, delta) of q1 from channel ch1 (zdim each),
 * lat2 the same for the fixed q2.  acc[0] += scale * sum_bt kl, acc[1] += mean_scale * sum_bt kl;
 * dlat1 (may be NULL) += scale * d(sum_bt kl)/dlat1.                                                            */
int idv_kl_fwd_bwd(const float* lat1, int H1, int ch1, const float* lat2, int H2, int ch2, int64_t n_bt, int zdim,
                   float scale, float mean_scale, float* dlat1, double* acc, void* stream);

/* ---- decoder side of the training step (train_second_phase_decoder.py:L376-433: SI-SNR through the decoder) -------
 *   idv_sisnr_fwd_bwd: si_snr (model/nsvae_loss.py:L877-889) of est (B, L) against src (B, L): loss[0] += -mean_b snr_b,
 *       d_est (may be NULL) += scale * d loss / d est; sums (B*3 doubles, overwritten) <- per utterance <est,src>, |src|^2,
 *       |est|^2 (what a per-utterance SI-SDR score needs, utils/eval_metrics.py:L49-64);
 *   idv_spec_loss_fwd_bwd: the two spectral terms of two_phase_loss.multi_recon_loss (model/nsvae_loss.py:L891-906):
 *       pred / ori = n_bins (re, im) pairs ((B, F, T, 2) fp32); acc[0] += inv_bt * sum |pred - ori|^2,
 *       acc[1] += inv_bt * sum (|pred|_eps - |ori|_quirk)^2 with |ori|_quirk = sqrt(or^2 + or^2 + 1e-6) (the reference
 *       squares the real part twice, L899); d_pred (may be NULL) += inv_bt * (w_cpx d/dpred cpx + w_mag d/dpred mag);
 *       inv_bt = 1 / (B * T) (the reference averages over batch and frames, sums over bins);
 *   idv_ola_bwd: adjoint of idv_ola_fwd: dframes (B*T, frame_ld) <- dsig (B, hop*(T-1)) / window envelope;
 *   idv_head_bwd: backward of the reconstruction head (real_imag: identity; mask: model/pvae_module.py:L2594-2609) for
 *       the last decoder layer.  raw (NB, F, T, 2) = transposed-conv output before ComplexBatchNormal, zb[6] its batch
 *       Z / b', drows[(b*T+t)][2k+part] (+ dpred (NB, F, T, 2) when not NULL: the gradient that reached `predict`
 *       directly) = gradient of the spectrum; writes fp32 planes (C = 1: [F][NB*(T+1)][16])
 *       y_planes <- raw and g_planes <- gradient w.r.t. the PReLU output, ready for idv_cbn_bwd_*.                  */
int idv_sisnr_fwd_bwd(const float* src, const float* est, int B, int L, float scale, float* d_est, double* sums,
                      double* loss, void* stream);
int idv_spec_loss_fwd_bwd(const float* pred, const float* ori, int64_t n_bins, float w_cpx, float w_mag, float inv_bt,
                          float* d_pred, double* acc, void* stream);
int idv_ola_bwd(const float* dsig, const float* wsq, int B, int T, int n_fft, int hop, int win, int frame_ld,
                float* dframes, void* stream);
int idv_head_bwd(const float* raw, const float* zb, float slope, int mask, const float* stft_x, const float* drows,
                 int drows_ld, const float* dpred, int NB, int F, int T, float* y_planes, float* g_planes, void* stream);
/* Last decoder layer (Cout = 1; the tap-GEMM needs K % 64 == 0, its output gradient has K = 2), SIMT:
 *   dy: fp32 planes [2 Fin - 1][NB*(T+1)][16] from idv_cbn_bwd_apply (re at channel 0, im at channel 8);
 *   w10: [10 (kf*2+kt)][Ktot][2] raw block weights (pack_dec5 without the CBN fold), this source's channels at k_off;
 *   idv_dec5_dgrad: dx fp32 planes [Fin][R][Cp] = gradient of the layer input (pad rows 0);
 *   idv_dec5_wgrad: dW [10][Ktot][2] += x^T dy  (x: fp32 or split-bf16 planes [Fin][R][Cp], Cp divides 256).
 * idv_reparam_bwd: gradient of the reparameterisation (model/pvae_module.py:L2177-2231, num_samples = 1) with the eps
 *   of the forward: dlatent (NB, T, Htot, 2) += dz (NB, T, zdim, 2) . dz/d(mu, log sigma, delta) at channel ch0.     */
int idv_dec5_dgrad(const float* dy, const float* w10, int Ktot, int k_off, int Cp, int Fin, int NB, int T, float* dx,
                   void* stream);
int idv_dec5_wgrad(const void* x, int x_split, const float* dy, int Ktot, int k_off, int Cp, int Fin, int NB, int T,
                   float* dW, void* stream);
/* y += a x (fp32, n % 4 == 0): adds the skip-connection gradient to an encoder layer's output gradient. */
int idv_axpy(float* y, const float* x, float a, int64_t n, void* stream);
int idv_reparam_bwd(const float* latent, int NB, int T, int Htot, int ch0, int zdim, const float* eps_r,
                    const float* eps_i, const float* dz, float* dlatent, void* stream);

/* ---- frame streaming (causal network; carried state instead of whole utterances) ------------------------------
 * Hop-synchronous streams: a step consumes hop*k new samples per stream and runs the same tap-GEMMs on k-frame
 * planes (Tp = -(k+1): pad rows kept) whose pad rows hold the last frame of the previous step.  The state kernels:
 *   idv_stream_frames_split: window = [hist (NB, win-hop) | x_new (NB, hop*k)] = samples base .. of every stream;
 *     frames [2][NB*k][kpad] split bf16, frame f = window[hop*f .. hop*f + win); a negative global sample index g
 *     reads sample -g (torch.stft's reflect padding at the start, model/pvae_module.py:L22);
 *   idv_stream_hist_shift:   hist <- the last win-hop samples of the window;
 *   idv_lstm_cell_step:      one nn.LSTM time step for the 4 (module, part) streams: gates (i,f,g,o) =
 *     g_in[stream row (b, frame)] (layout of idv_lstm_recurrent_fwd's g; NULL = none) + g_rec [4][NB][4H];
 *     c fp32 [4][NB][H] and h_split bf16 [2][4][NB][H] are updated in place, hseq (may be NULL) [4][NB*(T+1)][H]
 *     receives h at row b*(T+1)+1+frame;
 *   idv_carry_rows:          for every table entry copy row b*Tp + src_row to row b*Tp of every plane (the causal
 *     x[t-1] of the next step); *counter (may be NULL) += 1;
 *   idv_stream_ola:          acc (NB, win-hop) carried partial sums; frames (NB*k, frame_ld) synthesis frames of the
 *     global frames t0 .. t0+k-1; out (NB, hop*k) = the finished samples o = hop*t0 - win/2 + i divided by the
 *     window envelope over all frames t >= 0 (o < 0 is pre-roll and is discarded by the caller).              */
typedef struct {
  uint64_t base;        /* device address of the first plane */
  int64_t n_planes;     /* number of planes (the two bf16 halves of a split tensor count separately) */
  int64_t plane_bytes;  /* bytes between planes */
  int32_t row_bytes;    /* bytes per row, multiple of 16 */
  int32_t NB, Tp, src_row;
} idv_carry_t;
int idv_stream_frames_split(const float* hist, const float* x_new, int NB, int k, int64_t base, int hop, int win,
                            int kpad, void* frames, void* stream);
int idv_stream_hist_shift(float* hist, const float* x_new, int NB, int k, int hop, int win, void* stream);
int idv_lstm_cell_step(const float* g_in, int64_t g_m_off, int64_t g_p_off, int g_ld, const float* g_rec, int NB,
                       int H, int T, int frame, float* c, void* h_split, float* hseq, void* stream);
int idv_carry_rows(const idv_carry_t* table, int n_entries, uint64_t* counter, void* stream);
/* prev (NB, F, 2) <- last frame of the user-layout STFT chunk stft (NB, F, k, 2) (x[-1] of idv_enc0_fwd's next step) */
int idv_stream_last_frame(const float* stft, int NB, int F, int k, float* prev, void* stream);
int idv_stream_ola(const float* frames, int frame_ld, const float* wsq, float* acc, int NB, int k, int64_t t0,
                   int hop, int win, float* out, void* stream);
/* Everything of a step that only updates carried state or emits its output, as ONE launch (a step is a chain of small
 * dependent kernels: every launch costs its latency): idv_carry_rows(table, n_entries, counter) + idv_stream_hist_shift
 * (hist, x_new; both NULL = skip) + idv_stream_last_frame(stft (NB, F, k, 2), prev; prev NULL = skip) +
 * idv_stream_ola(frames, frame_ld, wsq, acc, t0, out).  Same contracts as the four entry points above.           */
int idv_stream_tail(const idv_carry_t* table, int n_entries, uint64_t* counter, float* hist, const float* x_new,
                    const float* stft, int F, float* prev, const float* frames, int frame_ld, const float* wsq,
                    float* acc, int64_t t0, float* out, int NB, int k, int hop, int win, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IDV_H_ */
