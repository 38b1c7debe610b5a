"""Frame-streaming enhancement with carried state (BASELINE config 5: the causal DCCRN-VAE, many concurrent streams,
hop-synchronous).

The reference only has whole-utterance forwards; a stream is the same causal computation cut into steps of ``k``
frames with the state a causal network needs carried between steps:

  * STFT  (model/pvae_module.py:L21-27): the last win-hop input samples (frame t covers x[hop*t - win/2, hop*t + win/2));
    the start of the stream reproduces torch.stft's reflect padding;
  * causal complex conv / transposed conv (model/complex_progress.py:L16-22, L244-250): x[t-1] of every layer input =
    the causal pad row of the plane set, refreshed by ``idv_carry_rows`` after every step;
  * ComplexLSTM (complex_progress.py:L58-74): (h, c) of both layers of the four (module, part) passes;
  * iSTFT (pvae_module.py:L38-42): the overlap-add tail of win-hop partial sums.

Every step runs the same C-ABI kernels as the whole-utterance path (tap-GEMMs on k-frame planes) plus the small state
kernels of csrc/stream.cu, and can be replayed as one CUDA graph.  A step's output is delayed by win/2 + hop samples
(the STFT look-ahead plus one hop of framing) – ``output_delay``.  In the interior the stream equals the
whole-utterance forward (tests/test_streaming*.py); the last frames of an utterance differ because the reference
reflect-pads the END of the signal, which a stream cannot know.
"""
import ctypes

import torch

from . import lib, modules, ops, pack
from .ops import Planes
from .pack import round8


def _carry_table(bufs, device):
    """bufs: list of (tensor, n_planes, NB, Tp).  Device table of idv_carry_t + the python-side description."""
    recs = (lib.CarryEntry * len(bufs))()
    entries = []
    for i, (t, n_planes, NB, Tp) in enumerate(bufs):
        plane_bytes = t.numel() * t.element_size() // n_planes
        row_bytes = plane_bytes // (NB * Tp)
        if row_bytes % 16 or row_bytes > 4096 or row_bytes * NB * Tp != plane_bytes:
            raise RuntimeError("carry rows must be multiples of 16 bytes, at most 4096")
        recs[i] = lib.CarryEntry(t.data_ptr(), n_planes, plane_bytes, row_bytes, NB, Tp, Tp - 1)
        entries.append((t, n_planes, NB, Tp, Tp - 1))
    raw = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8).clone().to(device)
    raw._entries = entries
    return raw


class StreamingEnhancer:
    """Streams ``n_streams`` signals through encoder ``enc`` (nsvae_pvae_dccrn_encoder_twophase or
    pvae_dccrn_encoder_skip_prepare, causal, num_samples == 1) and decoder ``dec`` (pvae_dccrn_decoder_skip_prepare, or
    nsvae_pvae_dccrn_decoder_twophase with ``pad``), ``frames_per_step`` frames per step.

        se = StreamingEnhancer(enc, dec, n_streams=128)
        se.prime(x[:, :hop])                     # the first hop samples only fill the look-ahead
        y = se.step(x[:, hop:hop + hop*k])       # (n_streams, hop*k); y[:, i] is output sample se.out_pos + i
    """

    def __init__(self, enc, dec, n_streams, frames_per_step=1, pad="sig", device="cuda", use_graph=True):
        if not ops.use_split():
            raise RuntimeError("streaming runs on the tensor-core path (IDV_GEMM=tc)")
        if not (enc.causal and dec.causal):
            raise NotImplementedError("frame streaming needs the causal network (model/causal_netconfig.py)")
        if enc.num_samples != 1 or dec.num_samples != 1:
            raise NotImplementedError("frame streaming draws one latent sample per frame (num_samples == 1)")
        if len(enc.lstms) != 1 or enc.lstms[0].num_layer != 2:
            raise NotImplementedError("frame streaming is built for the 2-layer ComplexLSTM of the shipped configs")
        self.enc, self.dec = enc, dec
        self.NB, self.k = int(n_streams), int(frames_per_step)
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        st = enc.stft
        self.n_fft, self.hop, self.win = st.n_fft, st.hop_length, st.win_length
        self.real_skips = isinstance(dec, modules.nsvae_pvae_dccrn_decoder_twophase) and pad == "sig" and dec.use_sc
        self.mask = dec.recon_type == "mask"
        self.output_delay = self.win // 2 + self.hop
        self.use_graph = bool(use_graph) and self.device.type == "cuda"
        self._alloc()

    # ------------------------------------------------------------------------------------------ state
    def _planes(self, C, F):
        NB, k = self.NB, self.k
        data = torch.zeros(2 * F * NB * (k + 1) * 2 * round8(C), dtype=torch.bfloat16, device=self.device)
        return Planes(data, NB, C, F, k, split=True)

    def _alloc(self):
        NB, k, dev = self.NB, self.k, self.device
        hl = self.win - self.hop
        nb = self.n_fft // 2 + 1
        self.hist = torch.zeros(NB, hl, device=dev)
        self.ola = torch.zeros(NB, hl, device=dev)
        self.stft_prev = torch.zeros(NB, nb, 2, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self.x_in = torch.zeros(NB, self.hop * k, device=dev)
        self.y_out = torch.zeros(NB, self.hop * k, device=dev)
        self.eps_in, self._eps_zero, self.spec_rows = None, None, None
        # encoder outputs (also the skip tensors), dense output, decoder outputs 0..n-2
        f, self.enc_bufs = nb, []
        for e in self.enc.encoders:
            c = e.conv.conv_re
            f = (f + 2 * c.padding[0] - c.kernel_size[0]) // c.stride[0] + 1
            self.enc_bufs.append(self._planes(c.out_channels, f))
        top = self.enc_bufs[-1]
        self.C, self.F = top.C, top.F
        # the decoder's first layer runs composed with the dense layer on the z planes (pack.pack_dense_conv_transpose): the
        # carried x[t-1] of that layer is z of the previous frame
        self.fused_dense = modules.FUSED_DENSE[0] and len(self.dec.decoders) > 1
        self.dense_buf = self._planes(self.enc.zdim if self.fused_dense else self.C, 1 if self.fused_dense else self.F)
        f, self.dec_bufs = self.F, []
        for d in list(self.dec.decoders)[:-1]:
            f = 2 * f - 1
            self.dec_bufs.append(self._planes(d.transconv.tconv_re.out_channels, f))
        H = self.enc.lstms[0].hidden_size
        self.H = H
        self.c_state = [torch.zeros(4 * NB * H, device=dev) for _ in range(2)]
        self.h_state = [torch.zeros(2 * 4 * NB * H, dtype=torch.bfloat16, device=dev) for _ in range(2)]
        self.hseq = torch.zeros(4 * NB * (k + 1) * H, device=dev)
        carried = self.enc_bufs + [self.dense_buf] + self.dec_bufs
        self.carry = _carry_table([(p.data, 2 * p.F, NB, k + 1) for p in carried], dev)
        self.n_carry = len(carried)
        self.steps = 0            # completed steps (frames processed = k * steps)
        self.primed = False
        self._graph = None
        self._lstm_packs = None
        self._mods, self._stamp = None, None

    def reset(self):
        """Forget all streams (state back to the start of a signal).  A captured graph stays valid."""
        for t in [self.hist, self.ola, self.stft_prev, self.hseq] + self.c_state + self.h_state + \
                [p.data for p in self.enc_bufs + [self.dense_buf] + self.dec_bufs]:
            t.zero_()
        self.steps, self.primed = 0, False

    @property
    def out_pos(self):
        """Whole-utterance output index of the first sample the NEXT step() returns (negative = pre-roll)."""
        return self.hop * self.k * self.steps - self.win // 2

    # ------------------------------------------------------------------------------------------ one step
    def _lstm(self):
        lstm = self.enc.lstms[0]
        if self._lstm_packs is None:
            re, im = dict(lstm.lstm_re.state_dict()), dict(lstm.lstm_im.state_dict())
            self._lstm_packs = [pack.pack_lstm_step(re, im, self.H, l, self.device) for l in range(2)]
        return lstm, self._lstm_packs

    def _step_impl(self, base, t0):
        NB, k, H = self.NB, self.k, self.H
        enc, dec = self.enc, self.dec
        hop, win, n_fft = self.hop, self.win, self.n_fft
        st = enc.stft
        if getattr(st, "_tc", None) is None or st._tc["bias"].device != self.device:
            st._tc = pack.pack_stft_tc(n_fft, win, self.device)
        hp = st._tc
        # ---- STFT of the k new frames
        frames = torch.empty(2 * NB * k * hp["kpad"], dtype=torch.bfloat16, device=self.device)
        lib.call("idv_stream_frames_split", self.hist, self.x_in, NB, k, int(base), hop, win, hp["kpad"], frames)
        stft_x = torch.empty((NB, hp["nbins"], k, 2), dtype=torch.float32, device=self.device)
        lib.call("idv_tapgemm_tc_head", frames, hp["kpad"], 1, None, 0, 0, NB * k, k, hp["wt"], hp["kc_max"], 1,
                 hp["bias"], hp["N"], hp["units"], hp["taps"], 1, None, 0, 0, 0, 0, 0, 0.0, 3, hp["nbins"], 1, 0, None,
                 stft_x, 0)
        # ---- encoder stack (state = pad rows of the static plane sets)
        p = enc.encoders[0].forward_from_stft(stft_x, False, out=self.enc_bufs[0].data, prev=self.stft_prev)
        for i in range(1, len(enc.encoders)):
            p = enc.encoders[i].forward_planes(p, False, out=self.enc_bufs[i].data)
        # ---- ComplexLSTM, one time step per frame on the carried (h, c)
        lstm, (pk0, pk1) = self._lstm()
        layers = lstm._packed(p.C, p.F, self.device)
        R = NB * (k + 1)
        g0 = ops.tapgemm(layers[0][0], p, None, NB, k, zero_pad_rows=False, out_split=False)
        hp0 = _HPlanes(self.h_state[0], NB, H)
        hp1 = _HPlanes(self.h_state[1], NB, H)
        for f in range(k):
            gr = ops.tapgemm(pk0, hp0, None, NB, 0, zero_pad_rows=False, out_split=False)
            lib.call("idv_lstm_cell_step", g0, 4 * H, R * 8 * H, 8 * H, gr, NB, H, k, f, self.c_state[0],
                     self.h_state[0], None)
            g1 = ops.tapgemm(pk1, hp0, hp1, NB, 0, zero_pad_rows=False, out_split=False)
            lib.call("idv_lstm_cell_step", None, 0, 0, 0, g1, NB, H, k, f, self.c_state[1], self.h_state[1], self.hseq)
        # ---- combine + reparameterisation + z planes in one launch (only the speech latent feeds the decoder; a supplied
        # eps covers the speech latent, the noise latent of a two-latent encoder then draws with eps = 0)
        zd, ln = enc.zdim, enc.latent_num
        if self.eps_in is not None:
            eps = list(self.eps_in[:2])
            if ln == 2:
                if self._eps_zero is None or self._eps_zero.shape != eps[0].shape:
                    self._eps_zero = torch.zeros_like(eps[0])
                eps += [self._eps_zero, self._eps_zero]
            _, _, zpl = ops.latent_fused(self.hseq, NB, k, H, k, zd, ln, 1, eps, 0, 0, True,
                                         zplanes_out=self.dense_buf.data if self.fused_dense else None)
        else:
            _, _, zpl = ops.latent_fused(self.hseq, NB, k, H, k, zd, ln, 1, None, self._seed, 0, True,
                                         offset_dev=self.counter,
                                         zplanes_out=self.dense_buf.data if self.fused_dense else None)
        n = len(dec.decoders)
        skips = {}
        if self.real_skips:
            for i in range(n):
                if i in dec.skip_to_use:
                    skips[i] = self.enc_bufs[n - 1 - i]
        first = 0
        if self.fused_dense:
            # dense + decoders[0] as one tap-GEMM on the z planes; only the very first frame of a signal lacks the dense
            # bias behind its x[t-1] tap (zero padding of the dense OUTPUT in the reference)
            q = dec.decoders[0].forward_after_dense(dec.dense, zpl[0], self.C, self.F, skips.get(0),
                                                    out=self.dec_bufs[0].data, first_frame=(t0 == 0))
            first = 1
        else:
            q = dec.dense.forward_planes(zpl[0], self.C, self.F, out=self.dense_buf.data)
        for i in range(first, n - 1):
            q = dec.decoders[i].forward_planes(q, skips.get(i), False, out=self.dec_bufs[i].data)
        predict = torch.empty((NB, hp["nbins"], k, 2), dtype=torch.float32, device=self.device)
        # ---- last layer + head; its epilogue also writes the K-major split rows of the synthesis GEMM
        ist = dec.istft
        if getattr(ist, "_tc", None) is None or ist._tc["bias"].device != self.device:
            ist._tc = pack.pack_istft_tc(n_fft, win, self.device)
        ip = ist._tc
        if self.spec_rows is None:          # padding columns stay zero: the head rewrites every live column per step
            self.spec_rows = torch.zeros(2 * NB * k * ip["kpad"], dtype=torch.bfloat16, device=self.device)
        rows = self.spec_rows
        last = dec.decoders[n - 1]
        fused_rows = last.head_on_tensor_cores(q, skips.get(n - 1))
        last.forward_head(q, skips.get(n - 1), self.mask, stft_x if self.mask else None, predict, 1, 0, False,
                          rows=rows if fused_rows else None)
        if not fused_rows:
            lib.call("idv_spec_rows_split", predict, NB, hp["nbins"], k, ip["kpad"], rows)
        # ---- iSTFT: synthesis frames + carried overlap-add
        N = ip["N"]
        fr = torch.empty(NB * k * N, dtype=torch.float32, device=self.device)
        lib.call("idv_tapgemm_tc", rows, ip["kpad"], 1, None, 0, 0, NB * k, 0, ip["wt"], ip["kc_max"], 1, ip["bias"], N,
                 ip["units"], ip["taps"], 1, fr, N, NB * k * N, 0, 0, 0, 0.0, 0)
        # ---- everything that only updates carried state or emits the output, in one launch: overlap-add, x[t-1] of every
        # layer input to the pad rows, sample history, last STFT frame, noise counter
        lib.call("idv_stream_tail", self.carry, self.n_carry, self.counter, self.hist, self.x_in, stft_x, hp["nbins"],
                 self.stft_prev, fr, N, ip["wsq"], self.ola, int(t0), self.y_out, NB, k, hop, win)
        self.predict = predict

    # ------------------------------------------------------------------------------------------ public
    def prime(self, x0):
        """First ``hop`` samples of every stream (they only fill the STFT look-ahead; no frame is complete yet)."""
        if self.primed:
            raise RuntimeError("stream already primed; call reset() first")
        x0 = lib.require_f32_cuda(x0, "x0")
        if tuple(x0.shape) != (self.NB, self.hop):
            raise RuntimeError("prime expects (%d, %d) samples" % (self.NB, self.hop))
        self.hist[:, -self.hop:].copy_(x0)
        self._seed = modules.philox_seed()
        self.primed = True

    def step(self, x, eps=None):
        """x: (n_streams, hop*k) new samples -> (n_streams, hop*k) output samples starting at ``out_pos`` (before the
        call).  eps: optional (eps_real, eps_imag) of shape (n_streams, 1, k, zdim) instead of on-device Philox."""
        if not self.primed:
            raise RuntimeError("call prime() with the first hop samples before step()")
        x = lib.require_f32_cuda(x, "x")
        if tuple(x.shape) != (self.NB, self.hop * self.k):
            raise RuntimeError("step expects (%d, %d) samples" % (self.NB, self.hop * self.k))
        self._check_weights()
        self.x_in.copy_(x)
        if eps is not None:
            if self.eps_in is None:
                self.eps_in = [torch.empty_like(lib.require_f32_cuda(e, "eps")) for e in eps]
                self._graph = None
            for dst, src in zip(self.eps_in, eps):
                dst.copy_(src)
        elif self.eps_in is not None:
            self.eps_in, self._graph = None, None
        t0 = self.k * self.steps
        base = self.hop * t0 - self.win // 2
        steady = base >= 0 and self.hop * t0 >= self.win           # no reflect, interior window envelope
        if self.use_graph and steady:
            if self._graph is None:
                self._step_impl(base, t0)                          # builds every pack outside the capture
                self._capture()
            else:
                self._graph.replay()
        else:
            self._step_impl(base, t0)
        self.steps += 1
        return self.y_out.clone()

    def _check_weights(self):
        """The LSTM step packs and the captured graph bake weight-pack pointers in: drop both when any parameter or
        buffer of the two models changed identity or version (load_state_dict, an optimiser step, .to()) - the same
        (data_ptr, _version) stamp modules._PackCache uses for the per-layer packs.  Parameters and buffers are looked
        up through their modules on every call, so re-bound tensors are seen too."""
        if self._mods is None:
            self._mods = list(self.enc.modules()) + list(self.dec.modules())
        stamp = tuple((t.data_ptr(), t._version) for m in self._mods for d in (m._parameters, m._buffers)
                      for t in d.values() if t is not None)
        if stamp != self._stamp:
            if self._stamp is not None:
                self._lstm_packs, self._graph = None, None
            self._stamp = stamp

    def _capture(self):
        """Capture one steady-state step (base >= 0: no reflect; t0 large: interior envelope) as a CUDA graph.  The
        eager step that just ran already advanced the state; the capture itself does not execute."""
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step_impl(1 << 40, 1 << 20)
        self._graph = g

    def enhance(self, x, eps=None):
        """Convenience driver: stream a whole batch of signals x (n_streams, L) and return the samples that a stream
        can produce, aligned with the whole-utterance output: y[:, o] for o in [0, n_out).  eps: optional
        (eps_real, eps_imag) of shape (n_streams, 1, T, zdim) indexed by global frame."""
        self.reset()
        hop, k = self.hop, self.k
        L = x.shape[1]
        n_steps = (L - hop) // (hop * k)
        self.prime(x[:, :hop].contiguous())
        outs = []
        for j in range(n_steps):
            lo = hop + j * hop * k
            e = None
            if eps is not None:
                e = [t[:, :, j * k:(j + 1) * k].contiguous() for t in eps]
            outs.append(self.step(x[:, lo:lo + hop * k].contiguous(), e))
        y = torch.cat(outs, 1)
        return y[:, self.win // 2:]                                  # drop the pre-roll (o < 0)


class _HPlanes(Planes):
    """Carried LSTM output h as a tap-GEMM source: 4 planes (streams) x NB rows x H channels, no pad rows."""

    def __init__(self, data, NB, H):
        Planes.__init__(self, data, NB, H, 4, 1, cp=H, split=True)

    @property
    def R(self):
        return self.NB
