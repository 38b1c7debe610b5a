"""torch.nn.Module surface of the reference, backed by libidv_b200.so.

Same class names, constructor / forward signatures and state_dict keys as
/root/reference/model/complex_progress.py and model/pvae_module.py (the hot-path classes listed in
SURVEY §8(b)); ``load_state_dict(reference_state_dict, strict=True)`` works on every class.  The
nn.Conv2d / nn.ConvTranspose2d / nn.LSTM / nn.Linear children exist only as parameter containers
(same names, shapes and default init as the reference); their forward is never called – all compute
goes through the C ABI (ops.py), and there is no CPU or eager fallback.

North-star aliases: ConvSTFT = STFT, ConviSTFT = ISTFT, ComplexBatchNorm = ComplexBatchNormal,
NavieComplexLSTM = ComplexLSTM.

``train=True`` runs the reference's train-mode forward (ComplexBatchNormal batch statistics and running-buffer
updates, model/complex_progress.py:L131-160).  Under ``torch.no_grad()`` that is all; with autograd enabled and
trainable parameters the encoders / decoders dispatch to train.py, whose single autograd node per model runs the
C-ABI backward kernels and writes the parameter gradients (phase 1, phase 2 and the end-to-end step; ``num_samples``
>= 1).  Both time geometries are built: the causal one (model/causal_netconfig.py: time pad 1, last column dropped, T
frames everywhere) and the non-causal one (model/net_config.py: no time pad, one frame fewer per encoder layer and
one more per decoder layer); activations keep the row layout of the STFT's T frames and carry their valid frame count
(ops.Planes.Tv).  Not built (raise NotImplementedError): the backward pass through the ``*_fc_latent`` / ``data_norm``
variants and through the non-causal net.
"""
import os

import torch
import torch.nn as nn

from . import ops, pack
from .ops import Planes

_TRAIN_MSG = "the fused reconstruction head runs its train-mode ComplexBatchNormal on one batch (out_bmul == 1)"


def _sd(module):
    """leaf-name -> tensor dict of a child container module (no copies)."""
    return dict(module.state_dict(keep_vars=True))


class _SlopeCache:
    """PReLU slopes are kernel ARGUMENTS (floats), so reading one is a device -> host copy that drains the GPU queue.  A
    training step changes every slope, and reading them block by block (13 syncs per step, each followed by that block's
    host-side weight re-packing with an empty queue) left the GPU idle for a fifth of the step.  All blocks register
    here; the first stale read refreshes EVERY stale block on that device with ONE copy."""

    def __init__(self):
        import weakref
        self.blocks = weakref.WeakSet()

    @staticmethod
    def _stamp(block):
        w = block.prelu.weight
        return (w.data_ptr(), w._version)

    def get(self, block):
        c = block.__dict__.get("_slope_cached")
        if c is not None and c[0] == self._stamp(block):
            return c[1]
        dev = block.prelu.weight.device
        stale = [b for b in self.blocks if b.prelu.weight.device == dev and
                 (b.__dict__.get("_slope_cached") is None or b.__dict__["_slope_cached"][0] != self._stamp(b))]
        if all(b is not block for b in stale):
            stale.append(block)
        vals = torch.stack([b.prelu.weight.detach().reshape(-1)[0] for b in stale]).cpu().tolist()
        for b, v in zip(stale, vals):
            b.__dict__["_slope_cached"] = (self._stamp(b), float(v))
        return block.__dict__["_slope_cached"][1]


_SLOPES = _SlopeCache()


class _PackCache:
    """Caches packed operands; rebuilt when any watched parameter/buffer changes identity or version."""

    def __init__(self):
        self._stamp = None
        self.items = {}

    def check(self, module):
        tensors = list(module.state_dict(keep_vars=True).values())
        stamp = tuple((t.data_ptr(), t._version) for t in tensors)
        if stamp != self._stamp:
            self._stamp = stamp
            self.items = {}
        return self.items


# --------------------------------------------------------------------------------------------------
# STFT / ISTFT — model/pvae_module.py:L12-42
# --------------------------------------------------------------------------------------------------
def _cached_zero_rows(owner, numel, shape_key, device):
    """Zero-initialised bf16 buffer kept on ``owner`` per (shape, device, CUDA stream): GEMM operand rows whose padding
    (pad rows, padding columns) must stay zero while the producing epilogue rewrites every live element on each call.
    One buffer per stream, so forwards in flight on different streams (pipeline.StreamPipeline) never share one; at most
    4 buffers are kept."""
    stream = torch.cuda.current_stream(device).cuda_stream if torch.device(device).type == "cuda" else 0
    key = (shape_key, str(device), stream)
    cache = owner.__dict__.setdefault("_rows_cache", {})
    buf = cache.get(key)
    if buf is None or buf.numel() != numel:
        if len(cache) >= 4:
            cache.pop(next(iter(cache)))
        buf = cache[key] = torch.zeros(numel, dtype=torch.bfloat16, device=device)
    return buf


class STFT(nn.Module):
    def __init__(self, n_fft, hop_length, win_length, device):
        super().__init__()
        self.n_fft, self.hop_length = n_fft, hop_length
        self.win_length = win_length
        self.window = torch.hann_window(self.win_length).to(device)   # plain attribute like the reference
        self._basis = None
        self.lengths = None        # int32 (B,) true sample counts of a zero-padded ragged batch (pipeline.enhance_ragged)

    def forward_with_rows(self, signal):
        """(stft_x, Planes): the spectrum in the reference layout AND as split-bf16 activation rows [1][B*(T+1)][576]
        for the first encoder layer's tap-GEMM (Encoder.forward_from_stft(rows=...)), both written by the STFT GEMM's
        epilogue.  The rows buffer is kept per (B, T, device): its pad rows / padding columns stay zero."""
        if not ops.use_split():
            raise RuntimeError("activation rows are written by the tensor-core STFT (IDV_GEMM=tc)")
        if getattr(self, "_tc", None) is None or self._tc["bias"].device != signal.device:
            self._tc = pack.pack_stft_tc(self.n_fft, self.win_length, signal.device)
        B, T = signal.shape[0], signal.shape[1] // self.hop_length + 1
        rows = _cached_zero_rows(self, 2 * B * (T + 1) * pack.ENC0_ROWS_LD, (B, T), signal.device)
        stft_x = ops.stft_tc(signal, self._tc, self.n_fft, self.hop_length, self.win_length, self.lengths, rows,
                             pack.ENC0_ROWS_LD, pack.ENC0_COL0)
        return stft_x, Planes(rows, B, 1, 1, T, cp=pack.ENC0_ROWS_LD, split=True)

    def forward(self, signal):
        if ops.use_split():                            # tensor-core DFT GEMM
            if getattr(self, "_tc", None) is None or self._tc["bias"].device != signal.device:
                self._tc = pack.pack_stft_tc(self.n_fft, self.win_length, signal.device)
            return ops.stft_tc(signal, self._tc, self.n_fft, self.hop_length, self.win_length, self.lengths)
        if self.lengths is not None:
            raise NotImplementedError("ragged batches run on the tensor-core path (IDV_GEMM=tc)")
        if self._basis is None or self._basis.device != signal.device:
            self._basis = pack.pack_stft_basis(self.n_fft, self.win_length, signal.device)
        return ops.stft(signal, self._basis, self.n_fft, self.hop_length, self.win_length)


class ISTFT(nn.Module):
    def __init__(self, n_fft, hop_length, win_length, device):
        super().__init__()
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length
        self.window = torch.hann_window(self.win_length).to(device)
        self._basis = None
        self.lengths = None        # see STFT.lengths

    def _ensure(self, device):
        if self._basis is None or self._basis[0].device != device:
            self._basis = pack.pack_istft_basis(self.n_fft, self.win_length, device)
        return self._basis

    def forward_ri(self, spec_ri):
        if ops.use_split():
            if getattr(self, "_tc", None) is None or self._tc["bias"].device != spec_ri.device:
                self._tc = pack.pack_istft_tc(self.n_fft, self.win_length, spec_ri.device)
            return ops.istft_tc(spec_ri, self._tc, self.n_fft, self.hop_length, self.win_length, self.lengths)
        if self.lengths is not None:
            raise NotImplementedError("ragged batches run on the tensor-core path (IDV_GEMM=tc)")
        basis, wsq = self._ensure(spec_ri.device)
        return ops.istft(spec_ri, basis, wsq, self.n_fft, self.hop_length, self.win_length)

    def spectrum_rows(self, B, T, device):
        """Zero-initialised split-bf16 K-major spectrum rows [2][B*T][kpad] for the fused head to fill (its padding
        columns must be zero: they meet zero weights in the DFT GEMM, and 0 * NaN garbage would not be 0).  One buffer
        per (B, T, device) is kept and reused: the head rewrites every live column on every call."""
        if getattr(self, "_tc", None) is None or self._tc["bias"].device != device:
            self._tc = pack.pack_istft_tc(self.n_fft, self.win_length, device)
        return _cached_zero_rows(self, 2 * B * T * self._tc["kpad"], (B, T), device)

    def forward_rows(self, rows, B, T):
        return ops.istft_rows_tc(rows, B, T, self._tc, self.n_fft, self.hop_length, self.win_length, self.lengths)

    def forward(self, x):
        """x: complex (B, F, T) like torch.istft's input at model/pvae_module.py:L41."""
        if not x.is_complex():
            raise RuntimeError("ISTFT.forward expects a complex (B, F, T) tensor")
        return self.forward_ri(torch.view_as_real(x).contiguous())


# --------------------------------------------------------------------------------------------------
# complex primitives — model/complex_progress.py
# --------------------------------------------------------------------------------------------------
def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class _ComplexConvBase(nn.Module):
    causal = True

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        self.conv_re = nn.Conv2d(in_channel, out_channel, kernel_size, stride=stride, padding=padding,
                                 dilation=dilation, groups=groups, bias=bias)
        self.conv_im = nn.Conv2d(in_channel, out_channel, kernel_size, stride=stride, padding=padding,
                                 dilation=dilation, groups=groups, bias=bias)
        self._cache = _PackCache()

    def _geometry(self):
        c = self.conv_re
        if _pair(c.dilation) != (1, 1) or c.groups != 1 or c.bias is None:
            raise NotImplementedError("complex conv kernels are built for dilation 1, groups 1, bias=True")
        (kh, kw), (sf, st), (pf, pt) = _pair(c.kernel_size), _pair(c.stride), _pair(c.padding)
        ok = (kw == 2 and pt == 1) if self.causal else (kw in (1, 2) and pt in (0, 1))
        if st != 1 or not ok:
            raise NotImplementedError(
                "built time geometries: causal kernel (k,2) / stride (s,1) / padding (p,1) with the last column "
                "dropped (model/complex_progress.py:L8-22) and non-causal kernel (k,1|2) / stride (s,1) / padding "
                "(p,0|1) (L24-36); got kernel %s stride %s padding %s causal=%s"
                % ((kh, kw), (sf, st), (pf, pt), self.causal))
        return kh, sf, pf

    def _time_geometry(self):
        """(time padding, change of the frame count) of this layer."""
        kw, pt = _pair(self.conv_re.kernel_size)[1], _pair(self.conv_re.padding)[1]
        return pt, 2 * pt - kw + 1 - (1 if self.causal else 0)

    def frames_out(self, frames_in):
        return frames_in + self._time_geometry()[1]

    def _packed(self, f_in, device, bn=None, slope=None):
        items = self._cache.check(self)
        key = (f_in, str(device), id(bn), slope)
        if key not in items:
            kh, sf, pf = self._geometry()
            items[key] = pack.pack_conv(self.conv_re.weight, self.conv_re.bias, self.conv_im.weight,
                                        self.conv_im.bias, bn, slope, f_in, sf, pf, device,
                                        pad_t=self._time_geometry()[0])
        return items[key]

    def run_packed(self, pk, xp, out=None, out_split=None):
        """One tap-GEMM launch of a pack built by pack.pack_conv for this layer's geometry (out: static
        streaming-state tensor, see ops.tapgemm; out_split=False: fp32 planes from split inputs)."""
        tv = self.frames_out(xp.Tv)
        if not 0 < tv <= xp.T:
            raise RuntimeError("conv output has %d frames, the row layout holds %d" % (tv, xp.T))
        out = ops.tapgemm(pk, xp, None, xp.NB, xp.T, t_valid=tv, out=out, out_split=out_split)
        split = xp.split if out_split is None else bool(out_split)
        return Planes(out, xp.NB, pk.c_out, pk.f_out, xp.T, split=split, Tv=tv)

    def forward_planes(self, xp, bn=None, slope=None):
        return self.run_packed(self._packed(xp.F, xp.data.device, bn, slope), xp)

    def forward(self, x):
        grow = max(self._time_geometry()[1], 0)
        return ops.planes_to_user(self.forward_planes(ops.user_to_planes(x, t_alloc=x.shape[3] + grow)))


class causal_complex_conv2d(_ComplexConvBase):
    """model/complex_progress.py:L8-22"""
    causal = True


class ComplexConv2d(_ComplexConvBase):
    """model/complex_progress.py:L24-36 (no frame dropped: kernel (k,2) with time padding 0 gives T-1 frames)."""
    causal = False


class _ComplexConvTransposeBase(nn.Module):
    causal = True

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, output_padding=0, dilation=1,
                 groups=1, bias=True):
        super().__init__()
        kw = dict(kernel_size=kernel_size, stride=stride, padding=padding, output_padding=output_padding,
                  groups=groups, bias=bias, dilation=dilation)
        self.tconv_re = nn.ConvTranspose2d(in_channel, out_channel, **kw)
        self.tconv_im = nn.ConvTranspose2d(in_channel, out_channel, **kw)
        self._cache = _PackCache()

    def _geometry(self):
        c = self.tconv_re
        if _pair(c.dilation) != (1, 1) or c.groups != 1 or c.bias is None or _pair(c.output_padding) != (0, 0):
            raise NotImplementedError("complex transposed conv kernels are built for dilation 1, groups 1, bias=True")
        (kh, kw), (sf, st), (pf, pt) = _pair(c.kernel_size), _pair(c.stride), _pair(c.padding)
        if st != 1 or pt != 0 or (kw != 2 if self.causal else kw not in (1, 2)):
            raise NotImplementedError(
                "built time geometries: kernel (k,2) / stride (s,1) / padding (p,0), last column dropped when causal "
                "(model/complex_progress.py:L222-250) or kept (L253-279)")
        return kh, sf, pf

    def frames_out(self, frames_in):
        return frames_in + _pair(self.tconv_re.kernel_size)[1] - 1 - (1 if self.causal else 0)

    def run_packed(self, pk, pp, skip, out=None, out_split=None):
        """One tap-GEMM launch of a pack built by pack.pack_conv_transpose for this layer (out_split=False: fp32
        planes from split inputs)."""
        tv = self.frames_out(pp.Tv)
        if not 0 < tv <= pp.T:
            raise RuntimeError("transposed conv output has %d frames, the row layout holds %d" % (tv, pp.T))
        if skip is not None and (skip.T != pp.T or skip.NB != pp.NB or skip.Tv != pp.Tv):
            raise RuntimeError("skip tensor has %d/%d frames x %d utterances, the decoder activation %d/%d x %d"
                               % (skip.Tv, skip.T, skip.NB, pp.Tv, pp.T, pp.NB))
        out = ops.tapgemm(pk, pp, skip, pp.NB, pp.T, t_valid=tv, out=out, out_split=out_split)
        split = pp.split if out_split is None else bool(out_split)
        return Planes(out, pp.NB, pk.c_out, pk.f_out, pp.T, split=split, Tv=tv)

    def _packed(self, f_in, c_p, c_skip, device, bn=None, slope=None):
        items = self._cache.check(self)
        key = (f_in, c_p, c_skip, str(device), id(bn), slope)
        if key not in items:
            kh, sf, pf = self._geometry()
            items[key] = pack.pack_conv_transpose(self.tconv_re.weight, self.tconv_re.bias, self.tconv_im.weight,
                                                  self.tconv_im.bias, bn, slope, f_in, c_p, c_skip, device, sf, pf)
        return items[key]

    def forward_planes(self, pp, skip=None, bn=None, slope=None):
        c_skip = skip.C if skip is not None else 0
        if pp.C + c_skip > self.tconv_re.in_channels:
            raise RuntimeError("transposed conv got %d+%d input channels, layer has %d"
                               % (pp.C, c_skip, self.tconv_re.in_channels))
        pk = self._packed(pp.F, pp.C, c_skip, pp.data.device, bn, slope)
        return self.run_packed(pk, pp, skip)

    def forward(self, x):
        if x.shape[1] != self.tconv_re.in_channels:
            raise RuntimeError("expected %d input channels, got %d" % (self.tconv_re.in_channels, x.shape[1]))
        return ops.planes_to_user(self.forward_planes(ops.user_to_planes(x, t_alloc=self.frames_out(x.shape[3]))))


class causal_ComplexConvTranspose2d(_ComplexConvTransposeBase):
    """model/complex_progress.py:L222-250"""
    causal = True


class ComplexConvTranspose2d(_ComplexConvTransposeBase):
    """model/complex_progress.py:L253-279 (keeps all T+1 output frames)."""
    causal = False


class ComplexBatchNormal(nn.Module):
    """model/complex_progress.py:L92-209 (eval branch; parameters/buffers identical)."""

    def __init__(self, C, H, W, momentum=0.9, dis_cbn=False):
        super().__init__()
        self.momentum = momentum
        self.gamma_rr = nn.Parameter(torch.ones(C), requires_grad=True)
        self.gamma_ri = nn.Parameter(torch.randn(C), requires_grad=True)
        self.gamma_ii = nn.Parameter(torch.ones(C), requires_grad=True)
        self.beta_r = nn.Parameter(torch.zeros(C), requires_grad=True)
        self.beta_i = nn.Parameter(torch.zeros(C), requires_grad=True)
        self.epsilon = 1e-5
        self.register_buffer('running_mean_real', torch.zeros(1, C, 1, 1))
        self.register_buffer('running_mean_imag', torch.zeros(1, C, 1, 1))
        self.register_buffer('Vrr', torch.ones(1, C, 1, 1))
        self.register_buffer('Vri', torch.zeros(1, C, 1, 1))
        self.register_buffer('Vii', torch.ones(1, C, 1, 1))
        self.init_flag = True
        self.detect_anormal = True
        self.dis_cbn = dis_cbn
        self._cache = _PackCache()

    def fold_inputs(self):
        return _sd(self)

    def forward(self, x, train=True):
        if train:
            return ops.cbn_train_user(x, self)
        items = self._cache.check(self)
        key = str(x.device)
        if key not in items:
            Z, bp = pack.cbn_fold(self.fold_inputs())
            items[key] = torch.cat((Z.reshape(-1, 4), bp), 1).to(torch.float32).contiguous().to(x.device)
        return ops.cbn_eval_user(x, items[key])


class ComplexLSTM(nn.Module):
    """model/complex_progress.py:L39-74"""

    def __init__(self, input_size, hidden_size, device, num_layers=1, bias=True, dropout=0, bidirectional=False):
        super().__init__()
        if bidirectional or not bias:
            raise NotImplementedError("ComplexLSTM kernels are built for unidirectional LSTMs with bias")
        self.num_layer = num_layers
        self.hidden_size = hidden_size
        self.device = device
        self.lstm_re = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers, bias=bias,
                               dropout=dropout, bidirectional=bidirectional)
        self.lstm_im = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers, bias=bias,
                               dropout=dropout, bidirectional=bidirectional)
        self._cache = _PackCache()

    def _packed(self, c_in, f_in, device):
        items = self._cache.check(self)
        key = (c_in, f_in, str(device))
        if key not in items:
            if c_in * f_in != self.lstm_re.input_size:
                raise RuntimeError("LSTM input size %d != C*F = %d*%d" % (self.lstm_re.input_size, c_in, f_in))
            re, im, H = _sd(self.lstm_re), _sd(self.lstm_im), self.hidden_size
            layers = [(pack.pack_lstm_inproj0(re, im, H, c_in, f_in, device), pack.pack_lstm_whh(re, im, 0, device))]
            for l in range(1, self.num_layer):
                layers.append((pack.pack_lstm_inproj1(re, im, H, device, l), pack.pack_lstm_whh(re, im, l, device)))
            items[key] = layers
        return items[key]

    def _packed_tc(self, cfg, device):
        items = self._cache.check(self)
        key = ("whh_tc", cfg[0], cfg[1], str(device))
        if key not in items:
            re, im = _sd(self.lstm_re), _sd(self.lstm_im)
            items[key] = [pack.pack_lstm_whh_tc(re, im, l, cfg[0], cfg[1], device) for l in range(self.num_layer)]
        return items[key]

    def _packed_wave(self, cfg, device):
        items = self._cache.check(self)
        key = ("wave", cfg, str(device))
        if key not in items:
            re, im = _sd(self.lstm_re), _sd(self.lstm_im)
            n, c = cfg[0], cfg[1]
            items[key] = (pack.pack_lstm_whh_tc(re, im, 0, n, c, device), pack.pack_lstm_whh_tc(re, im, 1, n, c, device, "ih"),
                          pack.pack_lstm_whh_tc(re, im, 1, n, c, device), pack.pack_lstm_bias_tc(re, im, 1, n, c, device))
        return items[key]

    def _packed_cluster(self, cfg, device):
        items = self._cache.check(self)
        key = ("cluster", cfg[0], cfg[1], str(device))
        if key not in items:
            re, im = _sd(self.lstm_re), _sd(self.lstm_im)
            u, c = cfg[0], cfg[1]
            items[key] = (pack.pack_lstm_cluster_tc(re, im, 0, u, c, device), pack.pack_lstm_cluster_tc(re, im, 1, u, c, device, "ih"),
                          pack.pack_lstm_cluster_tc(re, im, 1, u, c, device), pack.pack_lstm_cluster_bias(re, im, 1, u, c, device))
        return items[key]

    def forward_planes(self, xp, combine=True):
        """xp: Planes with C*F == input_size (feature d = c*F + f).  Returns the latent (NB, T, H, 2), or with
        combine=False the four uncombined streams ``(hseq fp32 [4][R][H], NB, T, H, Tv)`` for a fused consumer
        (ops.latent_fused / ops.lstm_combine_planes)."""
        layers = self._packed(xp.C, xp.F, xp.data.device)
        NB, T, H, Tv = xp.NB, xp.T, self.hidden_size, xp.Tv
        R = NB * (T + 1)
        src, hseq, split = xp, None, xp.split
        # two layers, batch <= 64: one wavefront kernel (layer 0 | layer-1 input projection | layer 1)
        wave = ops.lstm2_wave_supported(H, NB, xp.data.device) if (split and self.num_layer == 2 and ops.LSTM_WAVE[0]) else None
        # ... or, when faster (few utterances, or chunks of 16 that run concurrently): one thread-block cluster per (module, role), h exchanged through distributed shared memory
        clus = ops.lstm2_cluster_supported(H, NB, T, xp.data.device) if wave else None
        if wave:
            g = ops.tapgemm(layers[0][0], xp, None, NB, T, zero_pad_rows=False, out_split=False)
            if ops.GATE_HOOK[0] is not None:
                ops.GATE_HOOK[0]()
            while clus:
                w0, wi1, w1, b1 = self._packed_cluster(clus, xp.data.device)
                hseq = ops.lstm2_cluster_tc(g, 4 * H, R * 8 * H, 8 * H, w0, wi1, w1, b1, NB, T, H, clus[2], t_valid=Tv)
                # not co-resident: the second cluster shape if there is one, else the wavefront kernel below
                clus = ops.lstm2_cluster_next(H, NB, T, xp.data.device, clus) if hseq is None else None
            if hseq is None:                # (no clusters, or they are not co-resident on this device)
                w0, wi1, w1, b1 = self._packed_wave(wave, xp.data.device)
                hseq = ops.lstm2_wave_tc(g, 4 * H, R * 8 * H, 8 * H, w0, wi1, w1, b1, NB, T, H, wave[2], t_valid=Tv)
            return ops.lstm_combine(hseq, NB, T, H, Tv) if combine else (hseq, NB, T, H, Tv)
        # tensor-core recurrence when the planes are split-bf16 and the cooperative grid fits the device;
        # otherwise the fp32 SIMT recurrence (any batch size)
        cfg = ops.lstm_tc_supported(H, NB, xp.data.device) if split else None
        whh_tc = self._packed_tc(cfg, xp.data.device) if cfg else None
        for l, (inproj, whh) in enumerate(layers):
            g = ops.tapgemm(inproj, src, None, NB, T, zero_pad_rows=False, out_split=False)
            if l == 0 and ops.GATE_HOOK[0] is not None:
                ops.GATE_HOOK[0]()
            last = l + 1 == len(layers)
            more = split and not last                  # the next layer's tensor-core in-proj reads split h
            offs = (4 * H, R * 8 * H, 8 * H) if l == 0 else (2 * R * 4 * H, R * 4 * H, 4 * H)
            if cfg:
                hseq, hsp = ops.lstm_recurrent_tc(g, offs[0], offs[1], offs[2], whh_tc[l], NB, T, H,
                                                  want_f32=last, want_split=more, t_valid=Tv, cfg=cfg)
            else:
                hseq, hsp = ops.lstm_recurrent(g, offs[0], offs[1], offs[2], whh, NB, T, H, want_split=more,
                                               t_valid=Tv)
            # [4 streams][R][H]: plane = stream, row stride H
            if not last:
                src = Planes(hsp, NB, H, 4, T, cp=H, split=True, Tv=Tv) if split else \
                    Planes(hseq, NB, H, 4, T, cp=H, Tv=Tv)
        return ops.lstm_combine(hseq, NB, T, H, Tv) if combine else (hseq, NB, T, H, Tv)

    def forward(self, x):
        """x: (T, B, D, 2) -> (T, B, H, 2)"""
        if x.dim() != 4 or x.shape[-1] != 2:
            raise RuntimeError("ComplexLSTM expects (T, B, D, 2)")
        if self.hidden_size % 4:
            raise NotImplementedError("ComplexLSTM kernels need hidden_size %% 4 == 0")
        xu = x.permute(1, 2, 0, 3).unsqueeze(2).contiguous()          # (B, D, 1, T, 2)
        lat = self.forward_planes(ops.user_to_planes(xu))             # (B, T, H, 2)
        return lat.permute(1, 0, 2, 3).contiguous()


class ComplexDense(nn.Module):
    """model/complex_progress.py:L77-89"""

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.linear_read = nn.Linear(in_channel, out_channel)
        self.linear_imag = nn.Linear(in_channel, out_channel)
        self._cache = _PackCache()

    def _packed(self, c_out, f_out, device):
        items = self._cache.check(self)
        key = (c_out, f_out, str(device))
        if key not in items:
            if c_out * f_out != self.linear_read.out_features:
                raise RuntimeError("dense out_features %d != C*F = %d*%d" % (self.linear_read.out_features, c_out, f_out))
            items[key] = pack.pack_dense(self.linear_read.weight, self.linear_read.bias, self.linear_imag.weight,
                                         self.linear_imag.bias, c_out, f_out, device)
        return items[key]

    def forward_planes(self, zp, c_out, f_out, out=None):
        pk = self._packed(c_out, f_out, zp.data.device)
        out = ops.tapgemm(pk, zp, None, zp.NB, zp.T, t_valid=zp.Tv, out=out)
        return Planes(out, zp.NB, c_out, f_out, zp.T, split=zp.split, Tv=zp.Tv)

    def forward(self, x):
        """x: (..., D, 2) -> (..., out, 2)"""
        lead = x.shape[:-2]
        D = x.shape[-2]
        M = 1
        for s in lead:
            M *= s
        xu = x.reshape(1, M, D, 2).permute(0, 2, 1, 3).unsqueeze(2).contiguous()       # (1, D, 1, M, 2)
        outp = self.forward_planes(ops.user_to_planes(xu), self.linear_read.out_features, 1)
        y = ops.planes_to_user(outp)                                                      # (1, out, 1, M, 2)
        return y[0, :, 0].permute(1, 0, 2).reshape(*lead, self.linear_read.out_features, 2).contiguous()


# --------------------------------------------------------------------------------------------------
# Encoder / Decoder blocks — model/pvae_module.py:L45-93
# --------------------------------------------------------------------------------------------------
class Encoder(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, stride, chw, padding=None, causal=False):
        super().__init__()
        if padding is None:
            padding = [int((i - 1) / 2) for i in kernel_size]
        cls = causal_complex_conv2d if causal else ComplexConv2d
        self.conv = cls(in_channel=in_channel, out_channel=out_channel, kernel_size=kernel_size, stride=stride,
                        padding=padding)
        self.bn = ComplexBatchNormal(chw[0], chw[1], chw[2])
        self.prelu = nn.PReLU()
        self._cache = _PackCache()
        _SLOPES.blocks.add(self)

    def _slope(self):
        return _SLOPES.get(self)

    def forward_from_stft(self, stft_x, train=False, out=None, prev=None, raw_only=False, rows=None):
        """First layer (in_channel == 1): reads the user-layout STFT (B, F, T, 2) directly.  out / prev: streaming
        state (ops.enc0).  raw_only: return the conv output before ComplexBatchNormal / PReLU (training forward).
        rows: the STFT's activation rows (STFT.forward_with_rows) - the causal eval layer then runs as a tap-GEMM on the
        tensor cores (pack.pack_enc0_tc) instead of the SIMT kernel."""
        items = self._cache.check(self)
        if rows is not None and not train and not raw_only and out is None and self.conv.causal:
            key = ("enc0_tc", stft_x.shape[1], str(stft_x.device))
            if key not in items:
                self.conv._geometry()
                c = self.conv
                items[key] = pack.pack_enc0_tc(c.conv_re.weight, c.conv_re.bias, c.conv_im.weight, c.conv_im.bias,
                                               self.bn.fold_inputs(), self._slope(), stft_x.shape[1], stft_x.device)
            pk = items[key]
            data = ops.tapgemm(pk, rows, None, rows.NB, rows.T, t_valid=rows.T)
            return Planes(data, rows.NB, pk.c_out, pk.f_out, rows.T, split=True)
        key = ("enc0", bool(train), str(stft_x.device))
        if key not in items:
            self.conv._geometry()
            c = self.conv
            items[key] = pack.pack_enc0(c.conv_re.weight, c.conv_re.bias, c.conv_im.weight, c.conv_im.bias,
                                        None if train else self.bn.fold_inputs(), None if train else self._slope(),
                                        stft_x.device)
        w, b, cout, slope = items[key]
        if not self.conv.causal and self.conv._time_geometry() != (0, -1):
            raise NotImplementedError("first non-causal encoder layer: kernel (5,2) with time padding 0 expected")
        out = ops.enc0(stft_x, w, b, cout, slope, out_split=ops.use_split() and not raw_only, causal=self.conv.causal,
                       out=out, prev=prev)
        if raw_only:                                   # fp32: the batch statistics and the backward pass need y - mean
            return out                                 # at full precision (|mean| can be >> the channel's std)
        return ops.cbn_train_planes(out, self.bn, self._slope()) if train else out

    def forward_planes(self, xp, train=False, out=None, raw_only=False):
        items = self._cache.check(self)          # invalidates the child's fold when bn / prelu change
        key = ("raw" if train else "fold", xp.F, str(xp.data.device))
        if key not in items:
            kh, sf, pf = self.conv._geometry()
            c = self.conv
            items[key] = pack.pack_conv(c.conv_re.weight, c.conv_re.bias, c.conv_im.weight, c.conv_im.bias,
                                        None if train else self.bn.fold_inputs(), None if train else self._slope(),
                                        xp.F, sf, pf, xp.data.device, pad_t=c._time_geometry()[0])
        out = self.conv.run_packed(items[key], xp, out, out_split=False if raw_only else None)
        if raw_only:
            return out
        return ops.cbn_train_planes(out, self.bn, self._slope()) if train else out

    def forward(self, x, train):
        grow = max(self.conv._time_geometry()[1], 0)
        return ops.planes_to_user(self.forward_planes(ops.user_to_planes(x, t_alloc=x.shape[3] + grow), train))


class Decoder(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, stride, chw, padding=None, causal=False, if_bn=True):
        super().__init__()
        cls = causal_ComplexConvTranspose2d if causal else ComplexConvTranspose2d
        self.transconv = cls(in_channel=in_channel, out_channel=out_channel, kernel_size=kernel_size,
                             stride=stride, padding=padding)
        self.bn = ComplexBatchNormal(chw[0], chw[1], chw[2])
        self.prelu = nn.PReLU()
        self.if_bn = if_bn
        self._cache = _PackCache()
        _SLOPES.blocks.add(self)

    def _slope(self):
        return _SLOPES.get(self)

    def _fold(self):
        return (self.bn.fold_inputs(), self._slope()) if self.if_bn else (None, None)

    def forward_planes(self, pp, skip=None, train=False, out=None, raw_only=False):
        """raw_only: the transposed-conv output before ComplexBatchNormal / PReLU as fp32 planes (training forward)."""
        train = bool(train) and self.if_bn
        items = self._cache.check(self)
        c_skip = skip.C if skip is not None else 0
        key = ("raw" if train else "fold", pp.F, pp.C, c_skip, str(pp.data.device))
        if key not in items:
            kh, sf, pf = self.transconv._geometry()
            t = self.transconv
            if pp.C + c_skip > t.tconv_re.in_channels:
                raise RuntimeError("decoder layer got %d+%d input channels, has %d"
                                   % (pp.C, c_skip, t.tconv_re.in_channels))
            bn, slope = (None, None) if train else self._fold()
            items[key] = pack.pack_conv_transpose(t.tconv_re.weight, t.tconv_re.bias, t.tconv_im.weight,
                                                  t.tconv_im.bias, bn, slope, pp.F, pp.C, c_skip,
                                                  pp.data.device, sf, pf)
        if raw_only:
            if not train:
                raise RuntimeError("raw_only is the training forward (train=True with ComplexBatchNormal)")
            return self.transconv.run_packed(items[key], pp, skip, out, out_split=False)
        out = self.transconv.run_packed(items[key], pp, skip, out)
        return ops.cbn_train_planes(out, self.bn, self._slope()) if train else out

    def forward_after_dense(self, dense, zp, c_p, f_in, skip=None, out=None, first_frame=True):
        """ComplexDense ``dense`` + this (first, causal) decoder layer as ONE tap-GEMM on the z planes ``zp``
        (pack.pack_dense_conv_transpose: the two maps are composed at pack time; eval mode).  c_p, f_in: channels /
        planes of the dense output (model/pvae_module.py:L2085-2088).  Returns this layer's output planes.
        out / first_frame: frame streaming (static output planes; a step that does not start the signal)."""
        if not self.transconv.causal or not zp.split:
            raise RuntimeError("the dense + first-layer composition runs on the causal tensor-core path")
        items = self._cache.check(self)
        c_skip = skip.C if skip is not None else 0
        dstamp = tuple((t.data_ptr(), t._version) for t in dense.state_dict(keep_vars=True).values())
        key = ("dense_fused", dstamp, c_p, f_in, c_skip, str(zp.data.device))
        if key not in items:
            for k in [k for k in items if k[0] == "dense_fused"]:          # an older dense version's pack
                del items[k]
            kh, sf, pf = self.transconv._geometry()
            t = self.transconv
            if c_p + c_skip > t.tconv_re.in_channels or c_p * f_in != dense.linear_read.out_features:
                raise RuntimeError("dense output %d x %d (+ %d skip channels) does not fit this layer" % (c_p, f_in, c_skip))
            bn, slope = self._fold()
            items[key] = pack.pack_dense_conv_transpose(
                dense.linear_read.weight, dense.linear_read.bias, dense.linear_imag.weight, dense.linear_imag.bias,
                t.tconv_re.weight, t.tconv_re.bias, t.tconv_im.weight, t.tconv_im.bias, bn, slope, f_in, c_p, c_skip,
                zp.data.device, sf, pf)
        pk = items[key]
        if skip is not None and (skip.T != zp.T or skip.NB != zp.NB or skip.Tv != zp.Tv):
            raise RuntimeError("skip tensor has %d/%d frames x %d utterances, z %d/%d x %d"
                               % (skip.Tv, skip.T, skip.NB, zp.Tv, zp.T, zp.NB))
        out = ops.tapgemm(pk, zp, skip, zp.NB, zp.T, t_valid=zp.Tv, out=out, first_frame=first_frame)
        return Planes(out, zp.NB, pk.c_out, pk.f_out, zp.T, split=True, Tv=zp.Tv)

    def head_on_tensor_cores(self, pp, skip):
        kcs = [pp.Cp] + ([skip.Cp] if skip is not None else [])
        return pp.split and all(k % 64 == 0 for k in kcs)

    def forward_head(self, pp, skip, mask, stft_x, predict, out_bmul, out_boff, train=False, raw_only=False, rows=None):
        """Last layer (out_channel == 1) fused with the reconstruction head; writes ``predict``.  train=True: the raw
        transposed conv is written first, then CBN with batch statistics + PReLU (+ mask head) run in place.
        rows: ISTFT.spectrum_rows buffer the tensor-core head also fills (only with head_on_tensor_cores, eval)."""
        train = bool(train) and self.if_bn
        if train and out_bmul != 1:
            raise NotImplementedError(_TRAIN_MSG)
        if self.transconv.frames_out(pp.Tv) != pp.T or predict.shape[2] != pp.T:
            raise RuntimeError("the reconstruction head writes all %d frames of the row layout; the last decoder "
                               "layer produces %d" % (pp.T, self.transconv.frames_out(pp.Tv)))
        if skip is not None and (skip.T != pp.T or skip.Tv != pp.Tv):
            raise RuntimeError("skip tensor frames %d/%d != decoder activation frames %d/%d"
                               % (skip.Tv, skip.T, pp.Tv, pp.T))
        items = self._cache.check(self)
        c_skip = skip.C if skip is not None else 0
        key = ("head", train, pp.C, c_skip, str(pp.data.device))
        if key not in items:
            self.transconv._geometry()
            t = self.transconv
            bn, slope = (None, None) if train else self._fold()
            items[key] = pack.pack_dec5(t.tconv_re.weight, t.tconv_re.bias, t.tconv_im.weight, t.tconv_im.bias,
                                        bn, slope, pp.C, c_skip, pp.data.device)
        w, b, slope = items[key]
        kcs = [pp.Cp] + ([skip.Cp] if skip is not None else [])
        fused_mask = mask and not train          # train: raw output first (slope 1 = identity), head applied after CBN
        if pp.split and all(k % 64 == 0 for k in kcs):
            tkey = ("head_tc", train, pp.C, c_skip, pp.F, str(pp.data.device))
            if tkey not in items:
                items[tkey] = pack.pack_dec5_tc(w, b, slope, pp.F, kcs, pp.data.device)
            ops.dec5_head_tc(items[tkey], pp, skip, fused_mask, stft_x, predict, out_bmul, out_boff,
                             rows if not train else None)
        else:
            if rows is not None:
                raise RuntimeError("spectrum rows are written by the tensor-core head only")
            ops.dec5_head(pp, skip, w, b, slope, fused_mask, stft_x, predict, out_bmul, out_boff)
        if train and not raw_only:
            ops.head_train_user(predict, self.bn, self._slope(), mask, stft_x, 1)

    def forward(self, x, train=True):
        if x.shape[1] != self.transconv.tconv_re.in_channels:
            raise RuntimeError("expected %d input channels" % self.transconv.tconv_re.in_channels)
        return ops.planes_to_user(self.forward_planes(
            ops.user_to_planes(x, t_alloc=self.transconv.frames_out(x.shape[3])), None, train))


# --------------------------------------------------------------------------------------------------
# skip list: encoder outputs stay in planes; reference-layout tensors are materialised on access
# --------------------------------------------------------------------------------------------------
class SkipList(list):
    """``skiper`` of the encoder return tuple (model/pvae_module.py:L2236-2240).  Behaves like the
    reference's list of (B, C, F, T, 2) tensors (converted lazily, cached), while the decoders read the
    planes directly (no torch.cat, no repeat: SURVEY §2.1 K8)."""

    def __init__(self, planes):
        super().__init__([None] * len(planes))
        self.planes = list(planes)

    def _get(self, i):
        v = list.__getitem__(self, i)
        if v is None:
            v = ops.planes_to_user(self.planes[i])
            list.__setitem__(self, i, v)
        return v

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._get(j) for j in range(*i.indices(len(self)))]
        return self._get(i if i >= 0 else len(self) + i)

    def __iter__(self):
        return (self._get(i) for i in range(len(self)))


def _skip_planes(skiper, idx, split, t_alloc=None):
    if isinstance(skiper, SkipList) and skiper.planes[idx].split == split and \
            (t_alloc is None or skiper.planes[idx].T == t_alloc):
        return skiper.planes[idx]
    return ops.user_to_planes(skiper[idx], split=split, t_alloc=t_alloc)


def _build_encoders(net_params, causal):
    ch, ks = net_params["encoder_channels"], net_params["encoder_kernel_sizes"]
    st, pd, chw = net_params["encoder_strides"], net_params["encoder_paddings"], net_params["encoder_chw"]
    return [Encoder(in_channel=ch[i], out_channel=ch[i + 1], kernel_size=ks[i], stride=st[i], padding=pd[i],
                    chw=chw[i], causal=causal) for i in range(len(ch) - 1)]


def _build_decoders(net_params, causal, skip_to_use, use_sc=True):
    en_ch, de_ch = net_params["encoder_channels"], net_params["decoder_channels"]
    ks, st, pd = net_params["decoder_kernel_sizes"], net_params["decoder_strides"], net_params["decoder_paddings"]
    chw = net_params["decoder_chw"]
    out = []
    for i in range(len(de_ch) - 1):
        cin = de_ch[i] + (en_ch[len(en_ch) - 1 - i] if (use_sc and i in skip_to_use) else 0)
        out.append(Decoder(in_channel=cin, out_channel=de_ch[i + 1], kernel_size=ks[i], stride=st[i],
                           padding=pd[i], chw=chw[i], causal=causal))
    return out


def _run_encoder_stack(encoders, stft_x, train=False, rows=None):
    planes = [encoders[0].forward_from_stft(stft_x, train, rows=rows)]
    for enc in encoders[1:]:
        planes.append(enc.forward_planes(planes[-1], train))
    return planes


_philox_calls = [0]
ENC0_TC = [os.environ.get("IDV_ENC0_TC", "1") != "0"]        # first encoder layer as a tap-GEMM on STFT activation rows (A/B switch)
FUSED_SPEC_ROWS = [os.environ.get("IDV_FUSED_SPEC_ROWS", "1") != "0"]   # head epilogue writes the iSTFT GEMM's rows (A/B switch)
FUSED_DENSE = [os.environ.get("IDV_FUSED_DENSE", "1") != "0"]     # ComplexDense composed into the first decoder layer (A/B switch)
FUSED_LATENT = [os.environ.get("IDV_FUSED_LATENT", "1") != "0"]      # one idv_latent_fwd launch instead of lstm_combine + reparam + z_to_planes (A/B switch)


def _z_planes(z, B, S, s, split, t_alloc):
    """Decoder input planes of sample s of z (B*S, T, zdim, 2).  When z is the (unmodified) z_speech an encoder of
    this package returned, the planes idv_latent_fwd already wrote are used; any other tensor is converted."""
    tag = getattr(z, "_idv_planes", None)
    if tag is not None and tag[1] == z._version and len(tag[0]) == S:
        p = tag[0][s]
        if p.split == split and p.T == t_alloc and p.NB == B and p.data.device == z.device:
            return p
    return ops.z_to_planes(z, B, S, s, split=split, t_alloc=t_alloc)


def philox_seed():
    """Seed of the on-device Philox draws: torch's seed mixed with the rank, so the ranks of a data-parallel job draw
    different eps for their different utterances (every rank usually sets the same torch seed)."""
    seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        seed ^= ((torch.distributed.get_rank() + 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    return seed


def _next_philox():
    """(seed, draw counter): every draw advances the counter by one; the kernels turn a counter into its own block of
    Philox outputs (4 x 32 bit per element), so draws never share a uniform."""
    _philox_calls[0] += 1
    return philox_seed(), _philox_calls[0]


# --------------------------------------------------------------------------------------------------
# VAE encoders — model/pvae_module.py:L1791-1914, L2131-2268
# --------------------------------------------------------------------------------------------------
def _norm_consts(mean, std):
    """(scale, shift) of the data_mean / data_std normalisation and of its inverse, each (F, 2) fp32:
    (x - mean) / (std + 1e-6) = x * scale + shift (model/pvae_module.py:L367-368); std * y + mean (L483-484)."""
    m = mean.detach().to(torch.float32).reshape(-1, 2)
    sd = std.detach().to(torch.float32).reshape(-1, 2)
    scale = 1.0 / (sd + 1e-6)
    return (scale.contiguous(), (-m * scale).contiguous()), (sd.contiguous(), m.contiguous())


class _VaeEncoderBase(nn.Module):
    def _init_common(self, net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, latent_num,
                     heads=None, data_mean=None, data_std=None, channels=None, chw=None, lstm_in=None):
        """heads: names of the ComplexDense(zdim, zdim) latent heads of the *_fc_latent variants in the order
        (mean, logvar, delta) per latent (the LSTM then has zdim hidden units and there is no unused ``dense``);
        channels / chw / lstm_in: overrides of the channel-doubling variants."""
        self.device = device
        self.causal = causal
        self.latent_num = latent_num
        self.stft = STFT(n_fft, hop_len, win_length=win_length, device=device)
        self._heads = list(heads) if heads else None
        if self._heads is None:
            self.dense = ComplexDense(zdim, net_params["dense"][1])          # unused, present in the state_dict
        else:
            for name in self._heads:
                setattr(self, name, ComplexDense(zdim, zdim))
            self._head_cache = _PackCache()
        self.zdim = zdim
        self.num_samples = num_samples
        if channels is None:
            encoders = _build_encoders(net_params, causal)
        else:
            ks, st, pd = net_params["encoder_kernel_sizes"], net_params["encoder_strides"], net_params["encoder_paddings"]
            encoders = [Encoder(in_channel=channels[i], out_channel=channels[i + 1], kernel_size=ks[i], stride=st[i],
                                padding=pd[i], chw=chw[i], causal=causal) for i in range(len(channels) - 1)]
        lstm_dims = list(net_params["lstm_dim"]) if lstm_in is None else [lstm_in] + list(net_params["lstm_dim"][1:])
        hidden = int(zdim if self._heads is not None else 3 * zdim * latent_num)
        lstms = [ComplexLSTM(input_size=lstm_dims[i], hidden_size=hidden,
                             num_layers=net_params["lstm_layer_num"], device=device)
                 for i in range(len(lstm_dims) - 1)]
        self.encoders = nn.ModuleList(encoders)
        self.lstms = nn.ModuleList(lstms)
        self.epsilon = 1e-6
        self.register_buffer("data_mean", data_mean)
        self.register_buffer("data_std", data_std)
        self.datanorm = data_mean is not None and data_std is not None

    def _apply_heads(self, lat):
        """(B, T, zdim, 2) LSTM output -> (B, T, 3 * latent_num * zdim, 2): all ComplexDense heads as ONE tap-GEMM
        (model/pvae_module.py:L2478-2494; heads concatenated in (mean, logvar, delta) order per latent)."""
        B, T, H, _ = lat.shape
        mods = [getattr(self, n) for n in self._heads]
        n_out = len(mods) * self.zdim
        items = self._head_cache.check(nn.ModuleList(mods))
        key = str(lat.device)
        if key not in items:
            cat = lambda f: torch.cat([f(m) for m in mods])
            items[key] = pack.pack_dense(cat(lambda m: m.linear_read.weight), cat(lambda m: m.linear_read.bias),
                                         cat(lambda m: m.linear_imag.weight), cat(lambda m: m.linear_imag.bias),
                                         n_out, 1, lat.device)
        zp = ops.z_to_planes(lat, B, 1, 0, split=ops.use_split(), t_alloc=T)
        out = ops.tapgemm(items[key], zp, None, B, T, out_split=False if zp.split else None)
        u = ops.planes_to_user(Planes(out, B, n_out, 1, T, split=False))             # (B, n_out, 1, T, 2)
        return u[:, :, 0].permute(0, 2, 1, 3).contiguous()

    def _encode(self, x, train, eps):
        if len(self.lstms) != 1:
            raise NotImplementedError("one ComplexLSTM stage expected (lstm_dim has two entries)")
        token = step = None
        if train and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training step: the latent carries a grad_fn whose backward runs the C-ABI backward kernels (train.py)
            if self._heads is not None or self.datanorm:
                raise NotImplementedError("the backward pass is built for the encoders without latent heads / data_norm "
                                          "(nsvae_pvae_dccrn_encoder_twophase, pvae_dccrn_encoder_skip_prepare)")
            from . import train as _train
            latent_g, stft_x, planes, token, step = _train.encoder_train_forward(self, x)
            top, latent = planes[-1], latent_g.detach()
        else:
            latent_g = None
            rows = None
            if ENC0_TC[0] and ops.use_split() and self.causal and not train and not self.datanorm:
                stft_x, rows = self.stft.forward_with_rows(x)          # + activation rows for the first layer's tap-GEMM
            else:
                stft_x = self.stft(x)
            if self.datanorm:                                          # model/pvae_module.py:L367-371
                key = (self.data_mean._version, self.data_std._version, str(stft_x.device))
                if getattr(self, "_norm_key", None) != key:
                    self._norm_key, self._norm = key, _norm_consts(self.data_mean, self.data_std)[0]
                stft_x = ops.bin_affine(stft_x, self._norm[0], self._norm[1], zero_edge_imag=True, out=stft_x)
            planes = _run_encoder_stack(self.encoders, stft_x, train, rows)
            top = planes[-1]
            if self._heads is None and FUSED_LATENT[0]:
                # ONE launch: combine of the four LSTM streams, latent split, reparameterisation of every latent and the
                # z planes the decoder's ComplexDense reads (they ride on z_speech: see _z_planes)
                hs = self.lstms[0].forward_planes(top, combine=False)
                seed, off = (0, 0) if eps is not None else _next_philox()
                latent, zs, zpl = ops.latent_fused(*hs, self.zdim, self.latent_num, self.num_samples, eps, seed, off,
                                                   top.split)
                zs[0]._idv_planes = (zpl, zs[0]._version)
                skiper = SkipList(planes)
                skiper.grad_token, skiper.train_step = None, None
                return stft_x, skiper, latent, zs, top.C, top.F
            latent = self.lstms[0].forward_planes(top)                 # (B, T, hidden, 2)
            if self._heads is not None:
                latent = self._apply_heads(latent)                     # (B, T, 3*zdim*latent_num, 2)
        z, S = self.zdim, self.num_samples
        variant = 1 if self._heads is not None else 0
        zs = []
        for k in range(self.latent_num):
            if eps is not None:
                er, ei, seed, off = eps[2 * k], eps[2 * k + 1], 0, 0
            else:
                er = ei = None
                seed, off = _next_philox()
            if latent_g is not None and S == 1:
                zs.append(_train.reparam_train(latent_g, 3 * z * k, z, er, ei))    # differentiable z
            else:
                zs.append(ops.reparam(latent, 3 * z * k, z, S, er, ei, seed, off, variant=variant))
        skiper = SkipList(planes)
        skiper.grad_token, skiper.train_step = token, step
        return stft_x, skiper, (latent if latent_g is None else latent_g), zs, top.C, top.F

    def _forward12(self, x, train, eps):
        """The NSVAE 12-tuple (model/pvae_module.py:L2268)."""
        stft_x, skiper, lat, zs, C, F = self._encode(x, train, eps)
        z = self.zdim
        miu_s, ls_s, de_s = lat[:, :, 0:z, :], lat[:, :, z:2 * z, :], lat[:, :, 2 * z:3 * z, :]
        if self.latent_num == 1:
            return zs[0], miu_s, ls_s, de_s, None, None, None, None, skiper, C, F, stft_x
        miu_n, ls_n, de_n = lat[:, :, 3 * z:4 * z, :], lat[:, :, 4 * z:5 * z, :], lat[:, :, 5 * z:6 * z, :]
        return zs[0], miu_s, ls_s, de_s, zs[1], miu_n, ls_n, de_n, skiper, C, F, stft_x

    def _forward8(self, x, train, eps):
        """The CVAE 8-tuple (model/pvae_module.py:L1914)."""
        stft_x, skiper, lat, zs, C, F = self._encode(x, train, eps)
        z = self.zdim
        return zs[0], lat[:, :, 0:z, :], lat[:, :, z:2 * z, :], lat[:, :, 2 * z:, :], skiper, C, F, stft_x

    def reparameterization(self, miu, log_sigma, delta, num_samples, eps=None):
        """model/pvae_module.py:L2177-2231; eps = (eps_real, eps_imag) of shape (B, S, T, zdim) or None."""
        latent = torch.cat((miu, log_sigma, delta), dim=2).contiguous()
        er, ei = eps if eps is not None else (None, None)
        seed, off = (0, 0) if eps is not None else _next_philox()
        return ops.reparam(latent, 0, miu.shape[2], num_samples, er, ei, seed, off,
                           variant=1 if self._heads is not None else 0)


def _check_latent_num(latent_num):
    if latent_num not in (1, 2):
        raise ValueError("latent_num must be 1 or 2")


_NS_HEADS = ("speech_dense_mean", "speech_dense_logvar", "speech_dense_delta",
             "noise_dense_mean", "noise_dense_logvar", "noise_dense_delta")
_CVAE_HEADS = ("dense_mean", "dense_logvar", "dense_delta")


class nsvae_pvae_dccrn_encoder_twophase(_VaeEncoderBase):
    """model/pvae_module.py:L2131-2268.  Extra keyword ``eps``: list of supplied N(0,1) tensors
    (B, S, T, zdim) in draw order [speech_real, speech_imag(, noise_real, noise_imag)]; default = Philox."""

    def __init__(self, net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, latent_num):
        super().__init__()
        _check_latent_num(latent_num)
        self._init_common(net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, latent_num)

    def forward(self, x, train=True, eps=None):
        return self._forward12(x, train, eps)


class nsvae_dccrn_encoder_original(nsvae_pvae_dccrn_encoder_twophase):
    """model/pvae_module.py:L930-1076: the same computation as nsvae_pvae_dccrn_encoder_twophase."""


class nsvae_pvae_dccrn_encoder_twophase_fc_latent(_VaeEncoderBase):
    """model/pvae_module.py:L2353-2503: LSTM with zdim hidden units, ComplexDense heads for (mu, log sigma, delta) of
    the speech (and noise) latent, clamped reparameterisation (L2403-2450)."""

    def __init__(self, net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, latent_num):
        super().__init__()
        _check_latent_num(latent_num)
        self._init_common(net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, latent_num,
                          heads=_NS_HEADS[:3 * latent_num])

    def forward(self, x, train=True, eps=None):
        return self._forward12(x, train, eps)


class nsvae_dccrn_encoder_original_fc_latent(nsvae_pvae_dccrn_encoder_twophase_fc_latent):
    """model/pvae_module.py:L1077-1235: the same computation as nsvae_pvae_dccrn_encoder_twophase_fc_latent."""


class nsvae_dccrn_encoder_double_channel(_VaeEncoderBase):
    """model/pvae_module.py:L1236-1393: every encoder layer has twice the channels (the LSTM input too)."""

    def __init__(self, net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, latent_num):
        super().__init__()
        _check_latent_num(latent_num)
        ch = list(net_params["encoder_channels"])
        channels = [ch[0]] + [2 * c for c in ch[1:]]
        chw = [tuple(2 * t for t in c) for c in net_params["encoder_chw"]]
        self._init_common(net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, latent_num,
                          channels=channels, chw=chw, lstm_in=2 * net_params["lstm_dim"][0])

    def forward(self, x, train=True, eps=None):
        return self._forward12(x, train, eps)


class nsvae_dccrn_encoder_adapt_channel(_VaeEncoderBase):
    """model/pvae_module.py:L1394-1555: the encoder layers whose output is used as a skip tensor have twice the
    channels.  Like the reference, the constructor doubles the entries of ``net_params`` IN PLACE (L1411-1414)."""

    def __init__(self, net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, latent_num, skip_to_use):
        super().__init__()
        _check_latent_num(latent_num)
        ch, chw = net_params["encoder_channels"], net_params["encoder_chw"]
        for idx, c_num in enumerate(ch[1:]):
            if (len(ch) - 2 - idx) in skip_to_use:
                ch[idx + 1] = c_num * 2
                chw[idx] = tuple(t * 2 for t in chw[idx])
        lstm_in = 2 * net_params["lstm_dim"][0] if 0 in skip_to_use else net_params["lstm_dim"][0]
        self._init_common(net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, latent_num,
                          channels=list(ch), chw=list(chw), lstm_in=lstm_in)

    def forward(self, x, train=True, eps=None):
        return self._forward12(x, train, eps)


class pvae_dccrn_encoder_skip_prepare(_VaeEncoderBase):
    """model/pvae_module.py:L1791-1914"""

    def __init__(self, net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples):
        super().__init__()
        self._init_common(net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, 1)

    def forward(self, x, train=True, eps=None):
        return self._forward8(x, train, eps)


class pvae_dccrn_encoder_prob_skip(pvae_dccrn_encoder_skip_prepare):
    """model/pvae_module.py:L1556-1680: the same computation as pvae_dccrn_encoder_skip_prepare."""


class pvae_dccrn_encoder_skip_prepare_fc_latent(_VaeEncoderBase):
    """model/pvae_module.py:L1917-2043"""

    def __init__(self, net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples):
        super().__init__()
        self._init_common(net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, 1,
                          heads=_CVAE_HEADS)

    def forward(self, x, train=True, eps=None):
        return self._forward8(x, train, eps)


class pvae_dccrn_encoder(_VaeEncoderBase):
    """model/pvae_module.py:L259-394: the CVAE encoder with the optional data_mean / data_std normalisation."""

    def __init__(self, net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, data_mean=None,
                 data_std=None):
        super().__init__()
        self._init_common(net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, 1,
                          data_mean=data_mean, data_std=data_std)

    def forward(self, x, train=True, eps=None):
        return self._forward8(x, train, eps)


class pvae_dccrn_encoder_no_skip(_VaeEncoderBase):
    """model/pvae_module.py:L523-660 (data_mean / data_std are positional here)"""

    def __init__(self, net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, data_mean, data_std):
        super().__init__()
        self._init_common(net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, 1,
                          data_mean=data_mean, data_std=data_std)

    def forward(self, x, train=True, eps=None):
        return self._forward8(x, train, eps)


class pvae_dccrn_encoder_no_skip_fc_latent(_VaeEncoderBase):
    """model/pvae_module.py:L662-803"""

    def __init__(self, net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, data_mean, data_std):
        super().__init__()
        self._init_common(net_params, causal, device, zdim, n_fft, hop_len, win_length, num_samples, 1,
                          heads=_CVAE_HEADS, data_mean=data_mean, data_std=data_std)

    def forward(self, x, train=True, eps=None):
        return self._forward8(x, train, eps)


# --------------------------------------------------------------------------------------------------
# VAE decoders — model/pvae_module.py:L2045-2122, L2505-2619
# --------------------------------------------------------------------------------------------------
class _VaeDecoderBase(nn.Module):
    def _init_common(self, net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                     skip_to_use, use_sc, data_mean=None, data_std=None):
        if recon_type not in ("real_imag", "mask"):
            raise ValueError("recon_type must be 'real_imag' or 'mask'")
        self.device = device
        self.causal = causal
        self.num_samples = num_samples
        self.zdim = zdim
        self.recon_type = recon_type
        self.skip_to_use = skip_to_use
        self.dense = ComplexDense(zdim, net_params["dense"][1])
        self.use_sc = use_sc
        self.decoders = nn.ModuleList(_build_decoders(net_params, causal, skip_to_use, use_sc))
        self.istft = ISTFT(n_fft, hop_len, win_length=win_length, device=device)
        self.keep_decoder_outputs = False
        self.register_buffer("data_mean", data_mean)
        self.register_buffer("data_std", data_std)
        self.datanorm = data_mean is not None and data_std is not None

    def _decode(self, stft_x, z, skiper, C, F, train, real_skips, mask, self_skip=False):
        """self_skip: every skip slot is fed the layer's own input (pvae_dccrn_decoder_prob_skip, skip_prob = 2)."""
        BS, T, zdim, D = z.shape
        S = self.num_samples
        autograd = train and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if BS % S:
            raise RuntimeError("z batch %d is not a multiple of num_samples %d" % (BS, S))
        B = BS // S
        # train-mode forward with num_samples > 1: the batch statistics of every ComplexBatchNormal span all B*S rows
        # (model/pvae_module.py:L2550-2567), so the samples run as ONE batch with the skip tensors / noisy STFT repeated
        # per sample (row b*S + s <- utterance b) instead of the eval path's per-sample passes
        repeat = S if (train and S != 1 and not autograd) else 1
        if repeat > 1:
            B, S = BS, 1
        n = len(self.decoders)
        # frames of the row layout = frames of the reconstructed spectrum (the non-causal layers add one each)
        t_alloc = T
        for dec in self.decoders:
            t_alloc = dec.transconv.frames_out(t_alloc)
        skips, split = {}, ops.use_split()
        if real_skips or self_skip:
            for i in range(n):
                if self.use_sc and i in self.skip_to_use:
                    skips[i] = "self" if self_skip else _skip_planes(skiper, len(skiper) - i - 1, split, t_alloc)
                    if repeat > 1 and not self_skip:
                        skips[i] = ops.repeat_planes(skips[i], repeat)
        n_bins = F
        for _ in range(n):
            n_bins = 2 * n_bins - 1          # kernel 5 / stride 2 / pad 2 transposed conv
        predict = torch.empty((BS, n_bins, t_alloc, 2), dtype=torch.float32, device=z.device)
        if mask:
            stft_x = ops.lib.require_f32_cuda(stft_x, "stft_x")
            if repeat > 1:
                stft_x = stft_x.repeat_interleave(repeat, 0)
            if tuple(stft_x.shape[1:]) != (n_bins, t_alloc, 2):
                raise RuntimeError("stft_x %s does not match the reconstructed spectrum (B, %d, %d, 2)"
                                   % (tuple(stft_x.shape), n_bins, t_alloc))
        self.decoder_outputs = []
        rows = None          # the head also writes the iSTFT GEMM's operand rows when nothing touches predict in between
        if FUSED_SPEC_ROWS[0] and split and not train and not self.datanorm:
            rows = self.istft.spectrum_rows(BS, t_alloc, z.device)
        if autograd:
            # training step: (recon_sig, predict) carry a grad_fn whose backward runs the C-ABI backward kernels
            if self.datanorm:
                raise NotImplementedError("the backward pass is built for the decoders without data_norm")
            from . import train as _train
            recon_sig, predict = _train.decoder_train_forward(self, stft_x if mask else None, z, skiper, skips, C, F, mask)
            return recon_sig, torch.view_as_complex(predict)
        for s in range(S):
            zp = _z_planes(z, B, S, s, split, t_alloc)
            first = 0
            if FUSED_DENSE[0] and split and self.causal and not train and not self_skip and n > 1:
                # dense + decoders[0] as one launch on the z planes (composed at pack time)
                p = self.decoders[0].forward_after_dense(self.dense, zp, C, F, skips.get(0))
                if S == 1:
                    self.decoder_outputs.append(p)
                first = 1
            else:
                p = self.dense.forward_planes(zp, C, F)
            for i in range(first, n - 1):
                p = self.decoders[i].forward_planes(p, p if self_skip and i in skips else skips.get(i), train)
                if S == 1:
                    self.decoder_outputs.append(p)
            last_skip = p if self_skip and (n - 1) in skips else skips.get(n - 1)
            if rows is not None and not self.decoders[n - 1].head_on_tensor_cores(p, last_skip):
                rows = None
            self.decoders[n - 1].forward_head(p, last_skip, mask, stft_x if mask else None, predict, S, s, train,
                                              rows=rows)
        # model/pvae_module.py:L2090,L2099: the per-layer outputs (B*S, C, F, T, 2) are kept on the module.  Here they
        # stay activation planes and are converted on access (SkipList); with num_samples > 1 the sample passes run
        # one after the other on (B, ...) planes, so the list is only kept for num_samples == 1.
        self.decoder_outputs = SkipList(self.decoder_outputs) if S == 1 else []
        if self.datanorm:                                              # model/pvae_module.py:L483-484, L507-510
            key = (self.data_mean._version, self.data_std._version, str(predict.device))
            if getattr(self, "_norm_key", None) != key:
                self._norm_key, self._norm = key, _norm_consts(self.data_mean, self.data_std)[1]
            ops.bin_affine(predict, self._norm[0], self._norm[1], out=predict)
        recon_sig = self.istft.forward_rows(rows, BS, t_alloc) if rows is not None else self.istft.forward_ri(predict)
        return recon_sig, torch.view_as_complex(predict)


class pvae_dccrn_decoder_skip_prepare(_VaeDecoderBase):
    """model/pvae_module.py:L2045-2122: the skip slots are fed zeros (L2092-2097) – the skip half of
    every K loop is skipped instead of multiplying zeros."""

    def __init__(self, net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                 skip_to_use):
        super().__init__()
        self._init_common(net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                          skip_to_use, True)
        if recon_type != "real_imag":
            raise NotImplementedError("pvae_dccrn_decoder_skip_prepare only defines recon_type='real_imag' "
                                      "(model/pvae_module.py:L2117-2120)")

    def forward(self, stft_x, z, skiper, C, F, train=True):
        return self._decode(stft_x, z, skiper, C, F, train, real_skips=False, mask=False)


class pvae_dccrn_decoder(_VaeDecoderBase):
    """model/pvae_module.py:L396-521: real skip tensors (repeated num_samples times), optional data_norm."""

    def __init__(self, net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                 skip_to_use, resynthesis=False, data_mean=None, data_std=None):
        super().__init__()
        self._init_common(net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                          skip_to_use, True, data_mean, data_std)
        self.resynthesis = resynthesis
        if resynthesis:
            raise NotImplementedError("resynthesis=True calls self.stft, which this class does not have in the "
                                      "reference either (model/pvae_module.py:L488)")

    def forward(self, stft_x, z, skiper, C, F, train=True):
        return self._decode(stft_x, z, skiper, C, F, train, real_skips=True, mask=(self.recon_type == 'mask'))


class pvae_dccrn_decoder_no_skip(_VaeDecoderBase):
    """model/pvae_module.py:L805-928: no skip inputs at all, optional data_norm / resynthesis."""

    def __init__(self, net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                 resynthesis=False, data_mean=None, data_std=None):
        super().__init__()
        self._init_common(net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                          [], False, data_mean, data_std)
        self.resynthesis = resynthesis
        self.stft = STFT(n_fft, hop_len, win_length=win_length, device=device)

    def forward(self, stft_x, z, skiper, C, F, train=True):
        sig, predict = self._decode(stft_x, z, skiper, C, F, train, real_skips=False, mask=(self.recon_type == 'mask'))
        if self.resynthesis:
            predict = torch.view_as_complex(self.stft(sig))
        return sig, predict


class pvae_dccrn_decoder_prob_skip(_VaeDecoderBase):
    """model/pvae_module.py:L1681-1788: the CVAE decoder whose skip connections are dropped at random while training.
    Every forward draws ``torch.rand(1)`` from torch's global CPU generator like the reference (L1729; also in eval, so
    the generator advances identically); with train=True and a draw >= 0.5 the skip slots are fed zeros (skip_prob = 1)
    or the layer's own input (skip_prob = 2), otherwise - and always with train=False - the real skip tensors, repeated
    ``num_samples`` times.  Only recon_type 'real_imag' defines an output in the reference (L1783-1786).
    Extra keyword ``sc_flag``: force the draw's outcome (True = real skips) for parity runs."""

    def __init__(self, net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                 skip_to_use, skip_prob):
        super().__init__()
        self._init_common(net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                          skip_to_use, True)
        if recon_type != "real_imag":
            raise NotImplementedError("pvae_dccrn_decoder_prob_skip only defines recon_type='real_imag' "
                                      "(model/pvae_module.py:L1783-1786)")
        if skip_prob not in (1, 2):
            raise ValueError("skip_prob must be 1 (zero skips) or 2 (the layer's own input as its skip) - the reference "
                             "leaves zero_flag undefined otherwise (model/pvae_module.py:L1691-1694)")
        self.skip_prob = skip_prob
        self.zero_flag = skip_prob == 1

    def forward(self, stft_x, z, skiper, C, F, train=True, sc_flag=None):
        draw = torch.rand(1)
        if sc_flag is None:
            sc_flag = bool(draw[0] < 0.5) if train else True
        if sc_flag:
            return self._decode(stft_x, z, skiper, C, F, train, real_skips=True, mask=False)
        return self._decode(stft_x, z, skiper, C, F, train, real_skips=False, mask=False, self_skip=not self.zero_flag)


class nsvae_pvae_dccrn_decoder_twophase(_VaeDecoderBase):
    """model/pvae_module.py:L2505-2619"""

    def __init__(self, net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                 use_sc, skip_to_use, resynthesis):
        super().__init__()
        self._init_common(net_params, causal, device, num_samples, zdim, n_fft, hop_len, win_length, recon_type,
                          skip_to_use, use_sc)
        self.resynthesis = resynthesis
        self.stft = STFT(n_fft, hop_len, win_length=win_length, device=device)

    def forward(self, stft_x, z, skiper, C, F, train=True, pad='zero'):
        if pad not in ('zero', 'sig'):
            raise ValueError("pad must be 'zero' or 'sig'")
        sig, predict = self._decode(stft_x, z, skiper, C, F, train, real_skips=(pad == 'sig'),
                                    mask=(self.recon_type == 'mask'))
        if self.resynthesis:
            predict = torch.view_as_complex(self.stft(sig))
        return sig, predict


# --------------------------------------------------------------------------------------------------
# supervised DCCRN — model/pvae_module.py:L96-255
# --------------------------------------------------------------------------------------------------
class standard_DCCRN(nn.Module):
    def __init__(self, net_params, causal, device, skip_to_use):
        super().__init__()
        self.device = device
        self.causal = causal
        self.dense = ComplexDense(net_params["dense"][0], net_params["dense"][1])
        self.skip_to_use = skip_to_use
        lstm_dims = net_params["lstm_dim"]
        lstms = [ComplexLSTM(input_size=lstm_dims[i], hidden_size=lstm_dims[i + 1],
                             num_layers=net_params["lstm_layer_num"], device=device)
                 for i in range(len(lstm_dims) - 1)]
        self.encoders = nn.ModuleList(_build_encoders(net_params, causal))
        self.lstms = nn.ModuleList(lstms)
        self.decoders = nn.ModuleList(_build_decoders(net_params, causal, skip_to_use, True))
        self.linear = ComplexConv2d(in_channel=1, out_channel=1, kernel_size=1, stride=1)   # unused, in state_dict
        self.detect_anormal = True

    def forward_spec(self, stft_x, train, mask, rows=None, enc_rows=None):
        """stft_x (B, F, T, 2) -> predict (B, F, T, 2): decoder output, optionally through the mask head.
        rows: ISTFT.spectrum_rows buffer the fused head fills as well (eval, tensor-core path)."""
        stft_x = ops.lib.require_f32_cuda(stft_x, "stft_x")
        planes = _run_encoder_stack(self.encoders, stft_x, train, enc_rows)
        top = planes[-1]
        hs = self.lstms[0].forward_planes(top, combine=False)
        lat, zp = ops.lstm_combine_planes(*hs, top.split)               # (B, T, H, 2) + the dense layer's input plane
        if not train:
            self.latent = lat                                           # model/pvae_module.py:L187-188
        B, T = lat.shape[0], top.T
        n = len(self.decoders)
        first = 0
        if FUSED_DENSE[0] and top.split and self.causal and not train and n > 1:
            p = self.decoders[0].forward_after_dense(self.dense, zp, top.C, top.F,
                                                     planes[n - 1] if 0 in self.skip_to_use else None)
            first = 1
        else:
            p = self.dense.forward_planes(zp, top.C, top.F)
        for i in range(first, n - 1):
            p = self.decoders[i].forward_planes(p, planes[n - 1 - i] if i in self.skip_to_use else None, train)
        predict = torch.empty((B, stft_x.shape[1], T, 2), dtype=torch.float32, device=stft_x.device)
        last_skip = planes[0] if (n - 1) in self.skip_to_use else None
        self.rows_written = rows is not None and not train and self.decoders[n - 1].head_on_tensor_cores(p, last_skip)
        self.decoders[n - 1].forward_head(p, last_skip, mask, stft_x if mask else None, predict, 1, 0, train,
                                          rows=rows if self.rows_written else None)
        return predict

    def forward(self, x, train=True):
        """x: (B, 1, F, T, 2) -> (B, 1, F, T, 2) like model/pvae_module.py:L174-198."""
        return self.forward_spec(x[:, 0].contiguous(), train, mask=False).unsqueeze(1)


class DCCRN_(nn.Module):
    def __init__(self, n_fft, hop_len, net_params, causal, device, win_length, skip_to_use, recon_type, resynthesis,
                 data_mean, data_std):
        super().__init__()
        self.stft = STFT(n_fft, hop_len, win_length=win_length, device=device)
        self.std_DCCRN = standard_DCCRN(net_params, causal, device=device, skip_to_use=skip_to_use)
        self.istft = ISTFT(n_fft, hop_len, win_length=win_length, device=device)
        self.recon_type = recon_type
        self.resynthesis = resynthesis
        self.register_buffer("data_mean", data_mean)
        self.register_buffer("data_std", data_std)
        self.datanorm = self.data_mean is not None and self.data_std is not None
        if recon_type not in ("mask", "real_imag"):
            raise ValueError("recon_type must be 'mask' or 'real_imag'")

    def forward(self, signal, train=True):
        enc_rows = None
        if ENC0_TC[0] and ops.use_split() and self.std_DCCRN.causal and not train and not self.datanorm:
            stft_x, enc_rows = self.stft.forward_with_rows(signal)
        else:
            stft_x = self.stft(signal)
        if self.datanorm:                                              # model/pvae_module.py:L217-221, L236-239, L248-249
            key = (self.data_mean._version, self.data_std._version, str(stft_x.device))
            if getattr(self, "_norm_key", None) != key:
                self._norm_key, self._norm = key, _norm_consts(self.data_mean, self.data_std)
            stft_x = ops.bin_affine(stft_x, self._norm[0][0], self._norm[0][1], zero_edge_imag=True, out=stft_x)
        rows = None
        if FUSED_SPEC_ROWS[0] and ops.use_split() and not train and not self.datanorm:
            rows = self.istft.spectrum_rows(stft_x.shape[0], stft_x.shape[2], stft_x.device)
        predict = self.std_DCCRN.forward_spec(stft_x, train, mask=(self.recon_type == 'mask'), rows=rows, enc_rows=enc_rows)
        if self.datanorm:
            ops.bin_affine(predict, self._norm[1][0], self._norm[1][1], out=predict)
        if rows is not None and self.std_DCCRN.rows_written:
            clean = self.istft.forward_rows(rows, stft_x.shape[0], stft_x.shape[2])
        else:
            clean = self.istft.forward_ri(predict)
        if self.resynthesis:
            predict = self.stft(clean)
        return clean, torch.view_as_complex(predict)


# --------------------------------------------------------------------------------------------------
# GAN discriminator of train_second_phase_adversarial.py — model/pvae_module.py:L2271-2350
# --------------------------------------------------------------------------------------------------
class dis_Encoder(Encoder):
    """model/pvae_module.py:L2271-2293: Encoder whose ComplexBatchNormal keeps ``init_flag`` set (dis_cbn=True: the
    running statistics are overwritten by every train-mode batch instead of averaged)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride, chw, padding=None, causal=False):
        super().__init__(in_channel, out_channel, kernel_size, stride, chw, padding, causal)
        self.bn = ComplexBatchNormal(chw[0], chw[1], chw[2], dis_cbn=True)


class distinguisher(nn.Module):
    """model/pvae_module.py:L2294-2350: STFT -> 6 dis_Encoder blocks -> a REAL nn.LSTM over the flattened (C, F, re/im)
    features with ONE hidden unit -> (B, T, 1) score per frame.  The encoder stack runs on the tap-GEMM kernels; the
    LSTM's input projection (2560 -> 4 gates) is one more tap-GEMM (N = 32, 4 live columns) and its scalar recurrence
    one small kernel (idv_lstm_h1_fwd, one thread per utterance)."""

    def __init__(self, net_params, causal, device, zdim, n_fft, hop_len, win_length):
        super().__init__()
        self.device = device
        self.causal = causal
        self.stft = STFT(n_fft, hop_len, win_length=win_length, device=device)
        ch, ks = net_params["encoder_channels"], net_params["encoder_kernel_sizes"]
        st, pd, chw = net_params["encoder_strides"], net_params["encoder_paddings"], net_params["encoder_chw"]
        self.encoders = nn.ModuleList([dis_Encoder(in_channel=ch[i], out_channel=ch[i + 1], kernel_size=ks[i], stride=st[i],
                                                   padding=pd[i], chw=chw[i], causal=causal) for i in range(len(ch) - 1)])
        lstm_dims = net_params["lstm_dim"]
        self.lstms = nn.ModuleList([nn.LSTM(input_size=lstm_dims[i] * 2, hidden_size=1,
                                            num_layers=net_params["lstm_layer_num"], device=device)
                                    for i in range(len(lstm_dims) - 1)])
        self.epsilon = 1e-6
        self._cache = _PackCache()

    def forward(self, x, train=True):
        if len(self.lstms) != 1:
            raise NotImplementedError("one LSTM stage expected (lstm_dim has two entries)")
        if train and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("the distinguisher is built forward-only (eval and train-mode statistics); its "
                                      "backward pass (train_second_phase_adversarial.py) is not - call it under "
                                      "torch.no_grad() or freeze its parameters")
        stft_x = self.stft(x)
        top = _run_encoder_stack(self.encoders, stft_x, train)[-1]
        lstm = self.lstms[0]
        if top.C * top.F * 2 != lstm.input_size:
            raise RuntimeError("LSTM input size %d != 2*C*F = 2*%d*%d" % (lstm.input_size, top.C, top.F))
        items = self._cache.check(lstm)
        key = (top.C, top.F, str(top.data.device))
        if key not in items:
            items[key] = pack.pack_lstm_h1(_sd(lstm), lstm.num_layers, top.C, top.F, top.data.device)
        inproj, wrec = items[key]
        g = ops.tapgemm(inproj, top, None, top.NB, top.T, zero_pad_rows=False, out_split=False)      # [1][R][32]
        return ops.lstm_h1(g, inproj.out_ld, wrec, lstm.num_layers, top.NB, top.T, top.Tv)           # (B, T, 1)


# north-star aliases
ConvSTFT = STFT
ConviSTFT = ISTFT
ComplexBatchNorm = ComplexBatchNormal
NavieComplexLSTM = ComplexLSTM
