"""Thin Python front of the C ABI: shape bookkeeping + output allocation (torch is used only for device
memory and streams).  Every function ends in a ``lib.call`` – there is no other compute path."""
import os

import torch

from . import lib
from .pack import round8

# "tc": tcgen05 tensor-core tap-GEMMs on split-bf16 activations (default); "simt": fp32 SIMT kernels on fp32
# planes (bring-up / cross-check path).  Both are CUDA; neither is a fallback for the other at run time.
GEMM_MODE = [os.environ.get("IDV_GEMM", "tc")]


LSTM_WAVE = [os.environ.get("IDV_LSTM_WAVE", "1") != "0"]     # 2-layer wavefront kernel (else one launch per layer)


GATE_HOOK = [None]      # called right before the LSTM recurrence is launched (StreamPipeline staggers batches on it)


def set_gemm_mode(mode):
    if mode not in ("tc", "simt"):
        raise ValueError("mode must be 'tc' or 'simt'")
    GEMM_MODE[0] = mode


def use_split():
    return GEMM_MODE[0] == "tc"


class Planes:
    """Internal activation: fp32 [F][R][Cp], R = NB*(T+1) (row b*(T+1) is the causal zero row),
    Cp = 2*round8(C) with the real parts in [0, Ch) and the imaginary parts in [Ch, 2Ch).
    T is the ALLOCATED number of frames per utterance (fixes the row layout), Tv <= T the number of valid
    frames (non-causal layers shrink / grow it: model/net_config.py); rows of frames >= Tv are zero."""
    __slots__ = ("data", "NB", "C", "F", "T", "_cp", "split", "Tv")

    def __init__(self, data, NB, C, F, T, cp=None, split=False, Tv=None):
        """split=True: data is bf16 [2 (hi, lo)][F][R][Cp] (x ~= hi + lo), else fp32 [F][R][Cp]."""
        self.data, self.NB, self.C, self.F, self.T, self._cp, self.split = data, NB, C, F, T, cp, split
        self.Tv = T if Tv is None else int(Tv)
        if not 0 < self.Tv <= T:
            raise RuntimeError("valid frames %d outside (0, %d]" % (self.Tv, T))

    @property
    def Ch(self):
        return round8(self.C)

    @property
    def Cp(self):
        return self._cp if self._cp is not None else 2 * round8(self.C)

    @property
    def R(self):
        return self.NB * (self.T + 1)

    @property
    def plane_stride(self):
        return self.R * self.Cp


def _empty(n, device):
    return torch.empty(int(n), dtype=torch.float32, device=device)


def _empty_act(n, device, split):
    """n logical fp32 elements: fp32 [n] or bf16 [2][n]."""
    if split:
        return torch.empty(2 * int(n), dtype=torch.bfloat16, device=device)
    return torch.empty(int(n), dtype=torch.float32, device=device)


def stft(x, basis, n_fft, hop, win):
    x = lib.require_f32_cuda(x, "signal")
    if x.dim() != 2:
        raise RuntimeError("signal must be (B, L), got %s" % (tuple(x.shape),))
    B, L = x.shape
    T = L // hop + 1
    out = torch.empty((B, n_fft // 2 + 1, T, 2), dtype=torch.float32, device=x.device)
    lib.call("idv_stft_fwd", x, B, L, basis, n_fft, hop, win, out)
    return out


def istft(spec_ri, basis, wsq, n_fft, hop, win):
    """spec_ri: (B, n_fft/2+1, T, 2) fp32 -> (B, hop*(T-1))."""
    spec_ri = lib.require_f32_cuda(spec_ri, "spectrum")
    B, nb, T, _ = spec_ri.shape
    if nb != n_fft // 2 + 1:
        raise RuntimeError("spectrum has %d bins, expected %d" % (nb, n_fft // 2 + 1))
    frames = _empty(B * T * win, spec_ri.device)
    out = torch.empty((B, hop * (T - 1)), dtype=torch.float32, device=spec_ri.device)
    lib.call("idv_istft_fwd", spec_ri, B, T, basis, wsq, n_fft, hop, win, frames, out)
    return out


def _lengths_arg(lengths, B, device):
    if lengths is None:
        return None
    if not isinstance(lengths, torch.Tensor) or lengths.dtype != torch.int32 or lengths.numel() != B:
        raise RuntimeError("lengths must be an int32 tensor with one entry per utterance")
    return lengths.to(device).contiguous()


def stft_tc(x, hp, n_fft, hop, win, lengths=None, rows=None, rows_ld=0, rows_col0=0):
    """STFT on the tensor cores: frames (split bf16) -> one tap-GEMM against the windowed DFT basis whose
    epilogue writes the reference layout (B, nbins, T, 2).  lengths: int32 (B,) true sample counts of a zero-padded
    ragged batch (reflect padding at every utterance's own end).  rows: optional zero-initialised bf16
    [2][B*(T+1)][rows_ld] buffer that also receives the spectrum as split activation rows (first encoder layer)."""
    x = lib.require_f32_cuda(x, "signal")
    if x.dim() != 2:
        raise RuntimeError("signal must be (B, L), got %s" % (tuple(x.shape),))
    B, L = x.shape
    T = L // hop + 1
    R = B * T
    frames = torch.empty(2 * R * hp["kpad"], dtype=torch.bfloat16, device=x.device)
    lib.call("idv_stft_frames_split", x, B, L, n_fft, hop, win, hp["kpad"], _lengths_arg(lengths, B, x.device), frames)
    out = torch.empty((B, hp["nbins"], T, 2), dtype=torch.float32, device=x.device)
    lib.call("idv_tapgemm_tc_head", frames, hp["kpad"], 1, None, 0, 0, R, T, hp["wt"], hp["kc_max"], 1, hp["bias"],
             hp["N"], hp["units"], hp["taps"], 1, rows, rows_ld, 0, rows.numel() // 2 if rows is not None else 0, 0, 0, 0.0,
             3, hp["nbins"], 1, rows_col0, None, out, 0)
    return out


def istft_tc(spec_ri, hp, n_fft, hop, win, lengths=None):
    spec_ri = lib.require_f32_cuda(spec_ri, "spectrum")
    B, nb, T, _ = spec_ri.shape
    if nb != hp["nbins"]:
        raise RuntimeError("spectrum has %d bins, expected %d" % (nb, hp["nbins"]))
    R = B * T
    rows = torch.empty(2 * R * hp["kpad"], dtype=torch.bfloat16, device=spec_ri.device)
    lib.call("idv_spec_rows_split", spec_ri, B, nb, T, hp["kpad"], rows)
    N = hp["N"]
    frames = _empty(R * N, spec_ri.device)
    lib.call("idv_tapgemm_tc", rows, hp["kpad"], 1, None, 0, 0, R, 0, hp["wt"], hp["kc_max"], 1, hp["bias"], N,
             hp["units"], hp["taps"], 1, frames, N, R * N, 0, 0, 0, 0.0, 0)
    out = torch.empty((B, hop * (T - 1)), dtype=torch.float32, device=spec_ri.device)
    lib.call("idv_ola_fwd", frames, N, hp["wsq"], B, T, n_fft, hop, win, _lengths_arg(lengths, B, spec_ri.device), out)
    return out


STREAM_LIVE_ROWS = [os.environ.get("IDV_STREAM_LIVE_ROWS", "1") != "0"]      # A/B switch of TapGemmPack.tc_stream
STREAM_MAX_FRAMES = 16       # frames per step up to which a step's tap-GEMMs run on live rows (one unit per frame)


def tapgemm(pack, a0, a1, NB, T, zero_pad_rows=True, out_split=None, t_valid=0, out=None, first_frame=True):
    """Run one packed tap-GEMM.  a0/a1: Planes (a1 may be None).  Returns the flat output tensor
    [pack.out_planes][R][pack.out_ld] (fp32, or bf16 [2][...] when out_split).  Split inputs run on the
    tcgen05 kernel, fp32 inputs on the SIMT kernel.  t_valid: valid frames of the OUTPUT (0 = all T).
    out: write into this (static) tensor and leave its pad rows untouched (streaming state, Tp < 0 in the ABI).
    first_frame=False (composed dense + first decoder layer in a stream): frame 0 of the step is not the first frame of
    the signal, so every frame gets the regular bias."""
    R = NB * (T + 1)
    tp = ((T + 1) if zero_pad_rows else 0) if out is None else -(T + 1)
    if a0.split:
        if a1 is not None and not a1.split:
            raise RuntimeError("tap-GEMM sources must share one activation format")
        out_split = True if out_split is None else out_split
        n_out = pack.out_planes * R * pack.out_ld
        if out is None:
            out = _empty_act(n_out, a0.data.device, out_split)
        elif out.numel() != (2 * n_out if out_split else n_out):
            raise RuntimeError("static tap-GEMM output has %d elements, expected %d" % (out.numel(), n_out))
        b2 = getattr(pack, "bias_first", None)
        if STREAM_LIVE_ROWS[0] and tp <= 0 and 1 <= T <= STREAM_MAX_FRAMES and (b2 is None or not first_frame):
            # frame-streaming step (static output / no pad-row handling, a few frames): the GEMM over the live rows only
            # (a whole utterance without pad-row handling - the LSTM input projection - keeps the row layout: one unit
            # per frame would leave it NB rows per tile)
            st = pack.tc_stream(T, a0.Cp, a1.Cp if a1 is not None else 0)
            if st is not None:
                lib.call("idv_tapgemm_tc_splitk", a0.data, a0.Cp * (T + 1), a0.F, a1.data if a1 is not None else None,
                         a1.Cp * (T + 1) if a1 is not None else 0, a1.F if a1 is not None else 0, NB,
                         0, st["wt"], st["kc_max"], st["n_slots"], pack.bias, None, pack.N,
                         st["units"], st["taps"], st["n_units"], out, pack.out_ld * (T + 1), R * pack.out_ld, n_out,
                         1 if out_split else 0, 1 if pack.prelu else 0, pack.slope, 0, st["min_ksteps"])
                return out
        tc = pack.tc()
        if b2 is not None:        # layer composed with the dense map in front of it: own bias for every first frame
            if abs(tp) <= 1:
                raise RuntimeError("a composed (dense + transposed conv) pack needs the causal row layout")
            if not first_frame:
                b2 = pack.bias
            lib.call("idv_tapgemm_tc_b2", a0.data, a0.Cp, a0.F, a1.data if a1 is not None else None,
                     a1.Cp if a1 is not None else 0, a1.F if a1 is not None else 0, R,
                     tp, tc["wt"], tc["kc_max"], tc["n_slots"], pack.bias, b2, pack.N,
                     tc["units"], tc["taps"], pack.n_units, out, pack.out_ld, R * pack.out_ld, n_out,
                     1 if out_split else 0, 1 if pack.prelu else 0, pack.slope, int(t_valid))
            return out
        if tp <= 0 and tc.get("min_ksteps", 0) >= 2:          # small problems (streaming steps): split-K when tiles are few
            lib.call("idv_tapgemm_tc_splitk", a0.data, a0.Cp, a0.F, a1.data if a1 is not None else None,
                     a1.Cp if a1 is not None else 0, a1.F if a1 is not None else 0, R,
                     tp, tc["wt"], tc["kc_max"], tc["n_slots"], pack.bias, None, pack.N,
                     tc["units"], tc["taps"], pack.n_units, out, pack.out_ld, R * pack.out_ld, n_out,
                     1 if out_split else 0, 1 if pack.prelu else 0, pack.slope, int(t_valid), tc["min_ksteps"])
            return out
        lib.call("idv_tapgemm_tc", a0.data, a0.Cp, a0.F, a1.data if a1 is not None else None,
                 a1.Cp if a1 is not None else 0, a1.F if a1 is not None else 0, R,
                 tp, tc["wt"], tc["kc_max"], tc["n_slots"], tc.get("bias", pack.bias), tc.get("N", pack.N),
                 tc["units"], tc["taps"], tc.get("n_units", pack.n_units), out, pack.out_ld, R * pack.out_ld, n_out,
                 1 if out_split else 0, 1 if pack.prelu else 0, pack.slope, int(t_valid))
        return out
    if out_split:
        raise RuntimeError("the fp32 SIMT tap-GEMM writes fp32 planes only")
    if out is None:
        out = _empty(pack.out_planes * R * pack.out_ld, a0.data.device)
    lib.call("idv_tapgemm_f32",
             a0.data, a0.Cp, a0.plane_stride,
             a1.data if a1 is not None else None, a1.Cp if a1 is not None else 0,
             a1.plane_stride if a1 is not None else 0,
             R, tp,
             pack.w, pack.bias, pack.N, pack.units, pack.taps, pack.n_units,
             out, pack.out_ld, R * pack.out_ld, 1 if pack.prelu else 0, pack.slope, int(t_valid))
    return out


def enc0(stft_x, w, bias, cout, slope, out_split=False, causal=True, out=None, prev=None):
    """out / prev: streaming (static output planes whose pad rows are kept, previous STFT frame (B, Fin, 2))."""
    B, Fin, T, _ = stft_x.shape
    Fout = (Fin + 4 - 5) // 2 + 1
    Tv = T if causal else T - 1
    keep = out is not None
    if out is None:
        out = _empty_act(Fout * B * (T + 1) * 2 * cout, stft_x.device, out_split)
    lib.call("idv_enc0_fwd", stft_x, B, Fin, T, w, bias, cout, slope, out, 1 if out_split else 0,
             1 if causal else 0, Tv, prev, 1 if keep else 0)
    return Planes(out, B, cout, Fout, T, split=out_split, Tv=Tv)


def dec5_head(p, skip, w, bias, slope, mask, stft_x, predict, out_bmul, out_boff):
    if skip is not None and skip.split != p.split:
        raise RuntimeError("decoder sources must share one activation format")
    lib.call("idv_dec5_head_fwd", p.data, p.Cp, skip.data if skip is not None else None,
             skip.Cp if skip is not None else 0, 1 if p.split else 0, p.NB, p.F, p.T, w, bias, slope,
             1 if mask else 0, stft_x if mask else None, predict, out_bmul, out_boff)   # writes all p.T frames


def dec5_head_tc(hp, p, skip, mask, stft_x, predict, out_bmul, out_boff, rows=None):
    """Last decoder layer + head on the tensor-core kernel (hp = pack.pack_dec5_tc).  rows: optional bf16
    [2][NBtot*T][kpad] spectrum rows for istft_rows_tc (written next to ``predict``; padding columns stay untouched)."""
    R = p.NB * (p.T + 1)
    kpad = 0 if rows is None else rows.numel() // (2 * predict.shape[0] * predict.shape[2])
    lib.call("idv_tapgemm_tc_head", p.data, p.Cp, p.F, skip.data if skip is not None else None,
             skip.Cp if skip is not None else 0, skip.F if skip is not None else 0, R, p.T + 1,
             hp["wt"], hp["kc_max"], hp["n_slots"], hp["bias"], hp["N"], hp["units"], hp["taps"], hp["n_units"],
             rows, kpad, 0, rows.numel() // 2 if rows is not None else 0, 0, 1, hp["slope"], 2 if mask else 1,
             predict.shape[1], out_bmul, out_boff, stft_x if mask else None, predict, 0)


def istft_rows_tc(rows, B, T, hp, n_fft, hop, win, lengths=None):
    """iSTFT from spectrum rows the fused head already wrote (no idv_spec_rows_split pass): DFT tap-GEMM + overlap-add."""
    R = B * T
    N = hp["N"]
    frames = _empty(R * N, rows.device)
    lib.call("idv_tapgemm_tc", rows, hp["kpad"], 1, None, 0, 0, R, 0, hp["wt"], hp["kc_max"], 1, hp["bias"], N,
             hp["units"], hp["taps"], 1, frames, N, R * N, 0, 0, 0, 0.0, 0)
    out = torch.empty((B, hop * (T - 1)), dtype=torch.float32, device=rows.device)
    lib.call("idv_ola_fwd", frames, N, hp["wsq"], B, T, n_fft, hop, win, _lengths_arg(lengths, B, rows.device), out)
    return out


def lstm_recurrent(g, g_m_off, g_p_off, g_ld, whh, NB, T, H, want_split=False, t_valid=0):
    """Returns (hseq fp32 [4][R][H], hsplit bf16 [2][4][R][H] or None)."""
    hseq = _empty(4 * NB * (T + 1) * H, g.device)
    hsplit = _empty_act(4 * NB * (T + 1) * H, g.device, True) if want_split else None
    sync = torch.empty(2, dtype=torch.int32, device=g.device)
    lib.call("idv_lstm_recurrent_fwd", g, g_m_off, g_p_off, g_ld, whh, NB, T, H, hseq, hsplit, sync, int(t_valid))
    return hseq, hsplit


LSTM_LAYER_PAIRS = [os.environ.get("IDV_LSTM_LAYER_PAIRS", "1") != "0"]   # one-layer recurrence as CTA pairs (else one CTA per tile)


def lstm_tc_supported(H, NB, device):
    """Configuration of the one-layer-per-launch tensor-core recurrence: (gate columns per CTA, CTAs per module, kind,
    workspace bytes) with kind "pair" (idv_lstm_layer_pair_tc: CTA pairs, half of h streamed per CTA) or "tc"
    (idv_lstm_recurrent_tc), or None when the hidden size fits neither.  Batches of more than 64 utterances run as
    consecutive launches inside the entry points."""
    sms = torch.cuda.get_device_properties(device).multi_processor_count if torch.cuda.is_available() else 148
    if LSTM_LAYER_PAIRS[0]:
        cfg = lib.lstm_layer_pair_config(H)
        if cfg is not None and 2 * cfg[1] <= sms:
            return cfg[0], cfg[1], "pair", cfg[2]
    cfg = lib.lstm_tc_config(H)
    if cfg is None or 2 * cfg[1] > sms:
        return None
    return cfg[0], cfg[1], "tc", 0


def lstm_recurrent_tc(g, g_m_off, g_p_off, g_ld, wpack, NB, T, H, want_f32=True, want_split=False, t_valid=0, cfg=None):
    """Tensor-core recurrence of one layer; cfg from lstm_tc_supported (wpack packed with its first two entries).
    Returns (hseq fp32 or None, hsplit bf16 or None)."""
    n = 4 * NB * (T + 1) * H
    hseq = _empty(n, g.device) if want_f32 else None
    hsplit = _empty_act(n, g.device, True) if want_split else None
    if cfg is not None and cfg[2] == "pair":
        if g.is_cuda:
            lib.check_exclusive_device(g.device.index if g.device.index is not None else torch.cuda.current_device())
        work = torch.empty(int(cfg[3]), dtype=torch.uint8, device=g.device)
        sync = torch.empty(384, dtype=torch.int32, device=g.device)
        lib.call("idv_lstm_layer_pair_tc", g, g_m_off, g_p_off, g_ld, wpack, NB, T, H, hseq, hsplit, work, sync,
                 int(t_valid))
        return hseq, hsplit
    n_rg = (NB + 63) // 64
    hx = torch.empty(n_rg * 2 * 2 * 2 * 128 * H, dtype=torch.bfloat16, device=g.device)
    sync = torch.empty(n_rg * 2, dtype=torch.int32, device=g.device)
    lib.call("idv_lstm_recurrent_tc", g, g_m_off, g_p_off, g_ld, wpack, NB, T, H, hseq, hsplit, hx, sync,
             int(t_valid))
    return hseq, hsplit


def lstm2_wave_supported(H, NB, device):
    cfg = lib.lstm2_wave_config(H)
    if cfg is None:                       # (NB > 64: the entry point loops over chunks of 64 utterances)
        return None
    sms = torch.cuda.get_device_properties(device).multi_processor_count if torch.cuda.is_available() else 148
    return cfg if 6 * cfg[1] <= sms else None


def lstm2_wave_tc(g0, g_m_off, g_p_off, g_ld, w_hh0, w_ih1, w_hh1, bias1, NB, T, H, work_bytes, t_valid=0):
    """Both layers of the ComplexLSTM as one wavefront kernel.  Returns hseq1 fp32 [4][R][H]."""
    if g0.is_cuda:
        lib.check_exclusive_device(g0.device.index if g0.device.index is not None else torch.cuda.current_device())
    hseq = _empty(4 * NB * (T + 1) * H, g0.device)
    work = torch.empty(int(work_bytes), dtype=torch.uint8, device=g0.device)
    sync = torch.empty(384, dtype=torch.int32, device=g0.device)
    lib.call("idv_lstm2_wave_tc", g0, g_m_off, g_p_off, g_ld, w_hh0, w_ih1, w_hh1, bias1, NB, T, H, hseq, work, sync,
             int(t_valid))
    return hseq


LSTM_CLUSTER = [os.environ.get("IDV_LSTM_CLUSTER", "1") != "0"]   # cluster recurrence (weights in TMEM, DSMEM exchange)
LSTM_CLUSTER_MULTI = [os.environ.get("IDV_LSTM_CLUSTER_MULTI", "1") != "0"]   # ... also as several concurrent chunks of 16 utterances
_CLUSTER_OFF = set()                                                 # (device, H): the clusters are not co-resident here


_CLUSTER_CHUNKS = {}                                                 # (device, H, cs) -> chunks that run concurrently


def _cluster_concurrent_chunks(device, H, NB, cs):
    key = (str(device), H, cs)
    if key not in _CLUSTER_CHUNKS:
        n = None
        if torch.cuda.is_available() and torch.device(device).type == "cuda":
            with torch.cuda.device(device):
                n = lib.lstm2_cluster_concurrency(H, NB)
        if n is None:                     # no device (host-logic tests): clusters do not span GPCs (8 GPCs of ~18 SMs on a B200)
            n = 8 * (18 // cs)
        _CLUSTER_CHUNKS[key] = max(1, n // 6)
    return _CLUSTER_CHUNKS[key]


def lstm2_cluster_supported(H, NB, T, device):
    """(units per CTA, CTAs per cluster, workspace bytes) of idv_lstm2_cluster_tc when it is the faster recurrence for this
    problem, else None.  <= 16 utterances: always (1.58 / 2.64 ms against 4.0-4.25 ms of the wavefront kernel at H = 384, 4-s
    utterances).  More: the entry point runs chunks of 16 utterances, each on its own six clusters; the chunks whose clusters
    fit the device together run concurrently (clusters do not span GPCs: 5 chunks at H = 128, 1 at H = 384), so the choice is
    by estimated microseconds per time step (measured on B200, profiles/r02_lstm_chunks_vs_wave.json): a round of chunks
    2.2 + H / 200, a wavefront launch of <= 64 utterances 5.3 (H <= 128) or 7.3, of 65-128 interleaved utterances 1.3 x that."""
    if not LSTM_CLUSTER[0] or (str(device), H) in _CLUSTER_OFF:
        return None
    if lib.OPTIONS.get("gemm_dynamic_tiles", 0) or not lib.OPTIONS.get("lstm_wave_cta_pairs", 1):
        return None                       # kernels of several streams / processes share the GPU (see lib.check_exclusive_device)
    cfg = lib.lstm2_cluster_config(H, NB, T)
    if cfg is None or NB <= 16:
        return cfg
    if not LSTM_CLUSTER_MULTI[0] or cfg[2] > (8 << 30):
        return None
    n_chunks = -(-NB // 16)
    rounds = -(-n_chunks // _cluster_concurrent_chunks(device, H, NB, cfg[1]))
    full, rem = divmod(NB, 128)
    wave = (1.3 * full + (0.0 if rem == 0 else (1.0 if rem <= 64 else 1.3))) * (5.3 if H <= 128 else 7.3)
    return cfg if rounds * (2.2 + H / 200.0) < wave else None


def lstm2_cluster_next(H, NB, T, device, failed_cfg):
    """After lstm2_cluster_tc returned None for `failed_cfg`: the second cluster shape (option "lstm_cluster_alt") if the
    hidden size has one, else None (and the cluster recurrence stays off for this device and hidden size)."""
    if not lib.OPTIONS.get("lstm_cluster_alt", 0):
        lib.set_option("lstm_cluster_alt", 1)
        cfg = lib.lstm2_cluster_config(H, NB, T)
        if cfg is not None and cfg[:2] != failed_cfg[:2]:
            return cfg
    _CLUSTER_OFF.add((str(device), H))
    return None


def lstm2_cluster_tc(g0, g_m_off, g_p_off, g_ld, w_hh0, w_ih1, w_hh1, bias1, NB, T, H, work_bytes, t_valid=0):
    """Both layers of the ComplexLSTM, one thread-block cluster per (module, role) and chunk of <= 16 utterances.  Returns hseq1
    fp32 [4][R][H], or None when the clusters cannot be co-resident on this device (see lstm2_cluster_next)."""
    if g0.is_cuda:
        lib.check_exclusive_device(g0.device.index if g0.device.index is not None else torch.cuda.current_device())
    hseq = _empty(4 * NB * (T + 1) * H, g0.device)
    work = torch.empty(int(work_bytes), dtype=torch.uint8, device=g0.device)
    sync = torch.empty(128 * max(1, -(-NB // 8)), dtype=torch.int32, device=g0.device)
    ok = lib.call("idv_lstm2_cluster_tc", g0, g_m_off, g_p_off, g_ld, w_hh0, w_ih1, w_hh1, bias1, NB, T, H, hseq, work, sync,
                  int(t_valid), soft_resource=True)
    return hseq if ok is not False else None


def lstm_h1(g, g_ld, wrec, num_layers, NB, T, t_valid=0):
    """Real nn.LSTM with one hidden unit (the GAN distinguisher's head): g = layer-0 gate pre-activations [R][g_ld]
    (columns 0-3), wrec = pack.pack_lstm_h1's recurrent parameters.  Returns (NB, Tv, 1)."""
    Tv = t_valid if 0 < t_valid < T else T
    out = torch.empty((NB, Tv, 1), dtype=torch.float32, device=g.device)
    lib.call("idv_lstm_h1_fwd", g, g_ld, wrec, num_layers, NB, T, Tv, out)
    return out


def lstm_combine(hseq, NB, T, H, t_valid=0):
    Tv = t_valid if 0 < t_valid < T else T
    latent = torch.empty((NB, Tv, H, 2), dtype=torch.float32, device=hseq.device)
    lib.call("idv_lstm_combine_fwd", hseq, NB, T, H, latent, Tv)
    return latent


def reparam(latent, ch0, zdim, S, eps_r, eps_i, seed, offset, offset_dev=None, out=None, variant=0):
    """variant 1: the clamped formula of the *_fc_latent encoders (model/pvae_module.py:L2403-2450)."""
    NB, T, Htot, _ = latent.shape
    z = out if out is not None else torch.empty((NB * S, T, zdim, 2), dtype=torch.float32, device=latent.device)
    if eps_r is not None:
        eps_r = lib.require_f32_cuda(eps_r, "eps_r")
        eps_i = lib.require_f32_cuda(eps_i, "eps_i")
        if tuple(eps_r.shape) != (NB, S, T, zdim) or tuple(eps_i.shape) != (NB, S, T, zdim):
            raise RuntimeError("eps must have shape (B, S, T, zdim) = %s" % ((NB, S, T, zdim),))
    lib.call("idv_reparam_fwd", latent, NB, T, Htot, ch0, zdim, S, eps_r, eps_i, int(seed), int(offset), offset_dev,
             int(variant), z)
    return z


def latent_fused(hseq, NB, T, H, t_valid, zdim, latent_num, S, eps, seed, offset, split, offset_dev=None, zplanes_out=None):
    """ONE launch for the latent stage of the VAE encoders without heads: combine of the four LSTM streams, latent
    split, reparameterisation of every latent and the decoder's z planes (idv_latent_fwd).  eps: None (Philox) or the
    list [real_0, imag_0(, real_1, imag_1)] of (NB, S, Tv, zdim) tensors.  Returns (latent, [z_k], [Planes per sample]).
    zplanes_out: write the z planes into this static tensor and keep its pad rows (frame streaming)."""
    Tv = t_valid if 0 < t_valid < T else T
    dev = hseq.device
    latent = torch.empty((NB, Tv, H, 2), dtype=torch.float32, device=dev)
    zs = [torch.empty((NB * S, Tv, zdim, 2), dtype=torch.float32, device=dev) for _ in range(latent_num)]
    e = [None] * 4
    if eps is not None:
        if len(eps) != 2 * latent_num:
            raise RuntimeError("eps must hold %d tensors (real, imaginary draw per latent)" % (2 * latent_num))
        for i, t in enumerate(eps):
            t = lib.require_f32_cuda(t, "eps")
            if tuple(t.shape) != (NB, S, Tv, zdim):
                raise RuntimeError("eps must have shape (B, S, T, zdim) = %s" % ((NB, S, Tv, zdim),))
            e[i] = t
    Cp = 2 * round8(zdim)
    n_plane = NB * (T + 1) * Cp
    zpl = _empty_act(S * n_plane, dev, split) if zplanes_out is None else zplanes_out
    if zpl.numel() != S * n_plane * (2 if split else 1):
        raise RuntimeError("static z planes have %d elements, expected %d" % (zpl.numel(), S * n_plane * (2 if split else 1)))
    lib.call("idv_latent_fwd", hseq, NB, T, H, Tv, zdim, latent_num, S, e[0], e[1], e[2], e[3], int(seed), int(offset),
             offset_dev, latent, zs[0], zs[1] if latent_num == 2 else None, zpl, 1 if split else 0,
             0 if zplanes_out is None else 1)
    per = n_plane * (2 if split else 1)
    planes = [Planes(zpl[s * per:(s + 1) * per], NB, zdim, 1, T, split=split, Tv=Tv) for s in range(S)]
    return latent, zs, planes


def lstm_combine_planes(hseq, NB, T, H, t_valid, split):
    """lstm_combine + z_to_planes in one launch: returns (latent (NB, Tv, H, 2), Planes [1][R][2 round8(H)])."""
    Tv = t_valid if 0 < t_valid < T else T
    latent = torch.empty((NB, Tv, H, 2), dtype=torch.float32, device=hseq.device)
    data = _empty_act(NB * (T + 1) * 2 * round8(H), hseq.device, split)
    lib.call("idv_lstm_combine_planes", hseq, NB, T, H, Tv, latent, data, 1 if split else 0)
    return latent, Planes(data, NB, H, 1, T, split=split, Tv=Tv)


def bin_affine(x, scale, shift, zero_edge_imag=False, out=None):
    """x (B, F, T, 2) * scale (F, 2) + shift (F, 2) (data_mean / data_std normalisation and its inverse)."""
    x = lib.require_f32_cuda(x, "spectrum")
    B, F, T, _ = x.shape
    out = torch.empty_like(x) if out is None else out
    lib.call("idv_bin_affine", x, B, F, T, scale, shift, 1 if zero_edge_imag else 0, out)
    return out


def planes_to_user(p):
    out = torch.empty((p.NB, p.C, p.F, p.Tv, 2), dtype=torch.float32, device=p.data.device)
    lib.call("idv_planes_to_user", p.data, 1 if p.split else 0, p.NB, p.C, p.F, p.T, out, p.Tv)
    return out


def user_to_planes(x, split=False, t_alloc=None):
    """t_alloc: frames of the row layout (>= the tensor's own frame count; default = the tensor's)."""
    x = lib.require_f32_cuda(x, "activation")
    if x.dim() != 5 or x.shape[-1] != 2:
        raise RuntimeError("activation must be (B, C, F, T, 2), got %s" % (tuple(x.shape),))
    NB, C, F, Tv, _ = x.shape
    T = Tv if t_alloc is None else int(t_alloc)
    if T < Tv:
        raise RuntimeError("activation has %d frames, the row layout only %d" % (Tv, T))
    data = _empty_act(F * NB * (T + 1) * 2 * round8(C), x.device, split)
    lib.call("idv_user_to_planes", x, NB, C, F, T, data, 1 if split else 0, Tv)
    return Planes(data, NB, C, F, T, split=split, Tv=Tv)


def z_to_planes(z, NB, S, s, split=False, t_alloc=None):
    z = lib.require_f32_cuda(z, "z")
    _, Tv, zdim, _ = z.shape
    T = Tv if t_alloc is None else int(t_alloc)
    if T < Tv:
        raise RuntimeError("z has %d frames, the row layout only %d" % (Tv, T))
    data = _empty_act(NB * (T + 1) * 2 * round8(zdim), z.device, split)
    lib.call("idv_z_to_planes", z, NB, S, s, T, zdim, data, 1 if split else 0, Tv)
    return Planes(data, NB, zdim, 1, T, split=split, Tv=Tv)


def repeat_planes(p, S):
    """Rows b -> b*S + s (the unsqueeze(1).repeat(1, S, ...).view of model/pvae_module.py:L2564-2566) on activation
    planes: the skip tensors of a train-mode decoder pass with num_samples = S > 1, whose batch statistics span all
    B*S rows (device-memory copy; the eval path never materialises the repeat)."""
    Tp = p.T + 1
    lead = 2 if p.split else 1
    data = p.data.view(lead, p.F, p.NB, Tp, p.Cp).repeat_interleave(S, dim=2).reshape(-1).contiguous()
    return Planes(data, p.NB * S, p.C, p.F, p.T, cp=p._cp, split=p.split, Tv=p.Tv)


def cbn_eval_user(x, zb):
    x = lib.require_f32_cuda(x, "activation")
    B, C = x.shape[0], x.shape[1]
    inner = x[0, 0].numel() // 2
    out = torch.empty_like(x)
    lib.call("idv_cbn_eval_user", x, B, C, inner, zb, out)
    return out


# ---- ComplexBatchNormal(train=True), forward only ------------------------------------------------------------------
def _cbn_finalize(bn, acc, count, device, stats=None):
    """Batch statistics -> running-buffer update (first call copies, then EMA: complex_progress.py:L144-159) and the
    per-channel affine of the batch statistics.  ``bn`` is a modules.ComplexBatchNormal."""
    C = bn.gamma_rr.numel()
    zb = torch.empty(C * 6, dtype=torch.float32, device=device)
    first = bool(bn.init_flag)
    lib.call("idv_cbn_train_finalize", acc, float(count), C, bn.gamma_rr.detach(), bn.gamma_ri.detach(),
             bn.gamma_ii.detach(), bn.beta_r.detach(), bn.beta_i.detach(), bn.running_mean_real, bn.running_mean_imag,
             bn.Vrr, bn.Vri, bn.Vii, float(bn.momentum), 1 if first else 0, zb, stats)
    # the kernel rewrote the running buffers through raw pointers: bump their versions so every cached eval-mode fold
    # of this layer (modules._PackCache stamps are (data_ptr, _version)) is rebuilt from the new statistics
    for buf in (bn.running_mean_real, bn.running_mean_imag, bn.Vrr, bn.Vri, bn.Vii):
        torch.autograd.graph.increment_version(buf)
    if first and not bn.dis_cbn:
        bn.init_flag = False
    return zb


def cbn_train_planes(p, bn, slope):
    """In place on planes: batch statistics over (B, F, T), running-stat update, normalise (+ PReLU if slope)."""
    C = p.C
    acc = torch.empty(C * 5, dtype=torch.float64, device=p.data.device)
    lib.call("idv_cbn_stats_planes", p.data, 1 if p.split else 0, p.NB, C, p.F, p.T, acc, p.Tv)
    zb = _cbn_finalize(bn, acc, p.NB * p.F * p.Tv, p.data.device)
    lib.call("idv_cbn_apply_planes", p.data, 1 if p.split else 0, p.NB, C, p.F, p.T, zb,
             0 if slope is None else 1, 0.0 if slope is None else float(slope), p.Tv, None, 0)
    return p


def cbn_train_user(x, bn):
    """Stand-alone ComplexBatchNormal.forward(x, train=True) on the reference layout (B, C, ..., 2)."""
    x = lib.require_f32_cuda(x, "activation")
    B, C = x.shape[0], x.shape[1]
    inner = x[0, 0].numel() // 2
    acc = torch.empty(C * 5, dtype=torch.float64, device=x.device)
    lib.call("idv_cbn_stats_user", x, B, C, inner, acc)
    zb = _cbn_finalize(bn, acc, B * inner, x.device)
    out = torch.empty_like(x)
    lib.call("idv_cbn_eval_user", x, B, C, inner, zb, out)
    return out


def head_train_user(y, bn, slope, mask, stft_x, s_rep):
    """Last decoder layer in train mode: y (NB, F, T, 2) raw transposed-conv output in the reference layout ->
    CBN with batch statistics (C = 1), PReLU, optional mask head; in place."""
    NB = y.shape[0]
    inner = y[0].numel() // 2
    acc = torch.empty(5, dtype=torch.float64, device=y.device)
    lib.call("idv_cbn_stats_user", y, NB, 1, inner, acc)
    zb = _cbn_finalize(bn, acc, NB * inner, y.device)
    lib.call("idv_cbn_eval_user", y, NB, 1, inner, zb, y)
    lib.call("idv_head_user", y, inner, NB, float(slope), 1 if mask else 0, stft_x if mask else None, s_rep)
    return y
