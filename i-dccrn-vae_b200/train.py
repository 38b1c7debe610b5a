"""Training step of the NSVAE encoder (phase 1 of the reference: i_dccrn_vae/nsvae_dccrn/train_nsvae.py:L472-566 -
frozen clean / noise encoders, noisy encoder ``train=True``, closed-form KL loss, ``loss.backward()``, Adam).

``EncoderTrainStep`` runs the train-mode forward of ``nsvae_pvae_dccrn_encoder_twophase`` keeping what the backward
needs (raw conv outputs, batch statistics, LSTM outputs of both layers) and back-propagates a gradient of the latent
(B, T, 3*zdim*latent_num, 2) into ``.grad`` of every encoder parameter.  ``encoder(x, train=True)`` uses it through
one ``torch.autograd.Function`` when autograd is recording, so the reference's own loss code and
``loss.backward()`` / ``optimizer.step()`` work unchanged; autograd is only the plumbing that hands the latent
gradient over - every FLOP is a C-ABI kernel:

  * complex conv data gradients and all LSTM matrix products: the tcgen05 tap-GEMM with transposed weights;
  * weight gradients: the same kernel as a GEMM over the ROW dimension on transposed activation copies
    (``idv_planes_transpose_split``), K = rows x planes, split across enough units to fill the GPU;
  * ComplexBatchNormal(train) + PReLU backward, BPTT cell math, reductions: csrc/backward.cu.

Not differentiated: ``z`` (the phase-1 loss only uses mu / log sigma / delta, model/nsvae_loss.py:L275-328) and the
skip tensors (``w_resi = 0`` in the shipped configs); the unused ``dense.*`` gets no gradient, like the reference.
"""
import os

import torch

from . import lib, ops, pack
from .ops import Planes
from .pack import round8


BPTT_GRAPH = [os.environ.get("IDV_BPTT_GRAPH", "1") != "0"]     # replay the BPTT loop as a CUDA graph


def _rpad(R):
    return (R + 63) // 64 * 64


def _zeros(n, device, dtype=torch.float32):
    return torch.zeros(int(n), dtype=dtype, device=device)


def _transpose_split(data, split, F, R, Cp, shift):
    rp = _rpad(R)
    out = torch.empty(2 * F * Cp * rp, dtype=torch.bfloat16, device=data.device)
    lib.call("idv_planes_transpose_split", data, 1 if split else 0, F, R, Cp, rp, shift, out)
    return out


def _to_split(x):
    out = torch.empty(2 * x.numel(), dtype=torch.bfloat16, device=x.device)
    lib.call("idv_f32_to_split", x, x.numel(), out)
    return out


def _wgrad_gemm(a0, a1, a_planes, rows, wt, n_slots, N, rpad, units, taps, n_units):
    """out[unit][rows][N] fp32 = sum over the unit's taps of aT[plane] (rows x rpad) . wT[slot]^T (rpad x N)."""
    out = torch.empty(n_units * rows * N, dtype=torch.float32, device=wt.device)
    bias = _zeros(N, wt.device)
    lib.call("idv_tapgemm_tc", a0, rpad, a_planes, a1, rpad if a1 is not None else 0, a_planes if a1 is not None else 0,
             rows, 0, wt, rpad, n_slots, bias, N, units, taps, n_units, out, N, rows * N, 0, 0, 0, 0.0, 0)
    return out.view(n_units, rows, N)


class _HRows(Planes):
    """[4 streams][NB rows][cp] split planes without pad rows (one time step of the four LSTM passes)."""

    def __init__(self, data, NB, cp):
        Planes.__init__(self, data, NB, cp, 4, 1, cp=cp, split=True)

    @property
    def R(self):
        return self.NB


class EncoderTrainStep:
    def __init__(self, enc):
        if not enc.causal:
            raise NotImplementedError("the backward pass is built for the causal network (model/causal_netconfig.py)")
        if len(enc.lstms) != 1 or enc.lstms[0].num_layer != 2:
            raise NotImplementedError("the backward pass is built for the 2-layer ComplexLSTM of the shipped configs")
        if not ops.use_split():
            raise RuntimeError("training runs on the tensor-core path (IDV_GEMM=tc)")
        self.enc = enc
        self.saved = None
        self._packs = {}
        self.skip_grads = {}
        self._bptt_state = {}

    def add_skip_grads(self, dskips, n_dec):
        """Gradients of the skip tensors from a decoder's backward (decoder layer i reads encoder layer n_dec-1-i:
        model/pvae_module.py:L2556-2567); added to the layer's output gradient in ``backward``."""
        for i, g in dskips.items():
            k = n_dec - 1 - i
            if k in self.skip_grads:
                lib.call("idv_axpy", self.skip_grads[k], g, 1.0, g.numel())
            else:
                self.skip_grads[k] = g

    # ---------------------------------------------------------------------------------------------- forward
    def forward(self, x):
        self.skip_grads = {}
        with pack.on_device():                 # the weights change every step: repack on the GPU, not on the host
            return self._forward(x)

    def _forward(self, x):
        """x (B, L) -> latent (B, T, 3*zdim*latent_num, 2); keeps the tensors of the backward pass."""
        enc = self.enc
        stft_x = enc.stft(x)
        dev = stft_x.device
        sv = {"stft_x": stft_x, "layers": []}
        p = None
        for i, e in enumerate(enc.encoders):
            raw = e.forward_from_stft(stft_x, True, raw_only=True) if i == 0 else e.forward_planes(p, True, raw_only=True)
            C = raw.C
            acc = torch.empty(C * 5, dtype=torch.float64, device=dev)
            lib.call("idv_cbn_stats_planes", raw.data, 0, raw.NB, C, raw.F, raw.T, acc, raw.Tv)
            stats = torch.empty(C * 5, dtype=torch.float32, device=dev)
            zb = ops._cbn_finalize(e.bn, acc, raw.NB * raw.F * raw.Tv, dev, stats)
            slope = e._slope()
            # raw stays fp32 for the backward pass; the normalised activation goes straight into the split format of
            # the next layer's tensor-core GEMM
            act = torch.empty(2 * raw.data.numel(), dtype=torch.bfloat16, device=dev)
            lib.call("idv_cbn_apply_planes", raw.data, 0, raw.NB, C, raw.F, raw.T, zb, 1, slope, raw.Tv, act, 1)
            a = Planes(act, raw.NB, C, raw.F, raw.T, split=True, Tv=raw.Tv)
            sv["layers"].append({"x": p, "raw": raw, "stats": stats, "zb": zb, "slope": slope, "act": a})
            p = a
        lstm = enc.lstms[0]
        NB, T, H = p.NB, p.T, lstm.hidden_size
        R = NB * (T + 1)
        layers = lstm._packed(p.C, p.F, dev)
        cfg = ops.lstm_tc_supported(H, NB, dev)
        if cfg is None:
            raise NotImplementedError("training needs the tensor-core LSTM recurrence (H %% 64 == 0, batch fits the GPU)")
        whh_tc = lstm._packed_tc(cfg, dev)
        g = ops.tapgemm(layers[0][0], p, None, NB, T, zero_pad_rows=False, out_split=False)
        h0, h0s = ops.lstm_recurrent_tc(g, 4 * H, R * 8 * H, 8 * H, whh_tc[0], NB, T, H, want_f32=True, want_split=True,
                                        cfg=cfg)
        src = Planes(h0s, NB, H, 4, T, cp=H, split=True)
        g = ops.tapgemm(layers[1][0], src, None, NB, T, zero_pad_rows=False, out_split=False)
        h1, h1s = ops.lstm_recurrent_tc(g, 2 * R * 4 * H, R * 4 * H, 4 * H, whh_tc[1], NB, T, H, want_f32=True,
                                        want_split=True, cfg=cfg)
        # pad rows of the h planes are h(-1) = 0 for the recompute of the gates (the recurrence leaves them unwritten)
        for t in (h0, h1):
            t.view(4, NB, T + 1, H)[:, :, 0].zero_()
        for t in (h0s, h1s):
            t.view(2, 4, NB, T + 1, H)[:, :, :, 0].zero_()
        latent = ops.lstm_combine(h1, NB, T, H)
        sv.update(top=p, h0=h0, h0s=h0s, h1=h1, h1s=h1s, NB=NB, T=T, H=H)
        self.saved = sv
        return latent, stft_x, [l["act"] for l in sv["layers"]]

    # ---------------------------------------------------------------------------------------------- backward
    def _grad(self, param, value):
        value = value.to(param.dtype).reshape(param.shape)
        param.grad = value if param.grad is None else param.grad + value

    def backward(self, dlatent):
        """Accumulates dL/dparam into ``.grad`` of every encoder parameter given dL/dlatent."""
        sv = self.saved
        if sv is None:
            raise RuntimeError("backward() without a train-mode forward")
        dlatent = lib.require_f32_cuda(dlatent, "dlatent")
        with pack.on_device():
            g = self._lstm_backward(dlatent)
            for i in reversed(range(len(self.enc.encoders))):
                gs = self.skip_grads.pop(i, None)
                if gs is not None:
                    lib.call("idv_axpy", g, gs, 1.0, g.numel())
                    del gs
                g = self._conv_backward(i, g)
        self.saved = None

    # ---- LSTM ----
    def _lstm_pack(self, key, fn):
        if key not in self._packs:
            self._packs[key] = fn()
        return self._packs[key]

    def _lstm_backward(self, dlatent):
        sv = self.saved
        lstm = self.enc.lstms[0]
        NB, T, H, top = sv["NB"], sv["T"], sv["H"], sv["top"]
        dev = dlatent.device
        R = NB * (T + 1)
        rp = _rpad(R)
        re, im = dict(lstm.lstm_re.state_dict(keep_vars=True)), dict(lstm.lstm_im.state_dict(keep_vars=True))
        ver = tuple(t._version for t in list(re.values()) + list(im.values()))
        if self._packs.get("lstm_ver") != ver:
            self._packs = {"lstm_ver": ver}
        dH = torch.empty(4 * R * H, dtype=torch.float32, device=dev)
        lib.call("idv_lstm_combine_bwd", dlatent, NB, T, H, dH, T)
        h_split = {0: Planes(sv["h0s"], NB, H, 4, T, cp=H, split=True), 1: Planes(sv["h1s"], NB, H, 4, T, cp=H, split=True)}
        g_top = None
        for layer in (1, 0):
            below = h_split[0] if layer == 1 else top
            pk_g = self._lstm_pack(("gates", layer), lambda: pack.pack_lstm_gates(re, im, H, layer, top.C, top.F, dev))
            sk = pack.bptt_split_k(H)
            pk_hh = self._lstm_pack(("hh", layer), lambda: pack.pack_lstm_dgrad(re, im, H, layer, "hh", dev, split_k=sk))
            pk_ih = self._lstm_pack(("ih", layer), lambda: pack.pack_lstm_dgrad(re, im, H, layer, "ih", dev, top.C, top.F))
            # gate pre-activations of every step, cell states
            P = ops.tapgemm(pk_g, below, h_split[layer], NB, T, zero_pad_rows=False, out_split=False)
            cst = torch.empty(4 * R * H, dtype=torch.float32, device=dev)
            lib.call("idv_lstm_scan_c", P, NB, T, H, cst, T)
            # BPTT: one cell kernel + one dh = dP W_hh tap-GEMM per step, replayed as ONE CUDA graph per layer
            dP = self._bptt(layer, P, cst, dH, pk_hh, NB, T, H)
            del P, cst
            dPs = Planes(_to_split(dP), NB, 4 * H, 4, T, cp=4 * H, split=True)
            # gradient of the layer input
            gin = ops.tapgemm(pk_ih, dPs, None, NB, T, zero_pad_rows=True, out_split=False)
            # weight gradients: GEMMs over the row dimension
            dPT0 = _transpose_split(dP, False, 4, R, 4 * H, 0)
            dPT1 = _transpose_split(dP, False, 4, R, 4 * H, 1)
            hT = _transpose_split(sv["h%d" % layer], False, 4, R, H, 0)
            u, tp, n = self._lstm_pack(("wg", rp), lambda: pack.wgrad_lstm_tables(rp, dev))
            d_hh = _wgrad_gemm(dPT1, None, 4, 4 * H, hT, 4, H, rp, u, tp, n)                    # [2][4H][H]
            if layer == 1:
                belowT = _transpose_split(sv["h0"], False, 4, R, H, 0)
                d_ih = _wgrad_gemm(dPT0, None, 4, 4 * H, belowT, 4, H, rp, u, tp, n)
            else:
                ch = top.Cp // 2
                topT = _transpose_split(top.data, True, top.F, R, top.Cp, 0)                  # [F][2ch][rp] = slots f*2+p
                u0, tp0, n0 = self._lstm_pack(("wg0", rp), lambda: pack.wgrad_lstm_tables(rp, dev, top.F))
                o = _wgrad_gemm(dPT0, None, 4, 4 * H, topT, 2 * top.F, ch, rp, u0, tp0, n0)   # [2*F][4H][ch]
                d_ih = o.view(2, top.F, 4 * H, ch)[:, :, :, :top.C].permute(0, 2, 3, 1).reshape(2, 4 * H, top.C * top.F)
            db = _zeros(2 * 4 * H, dev)
            for s in range(4):
                lib.call("idv_colsum_add", dP[s * R * 4 * H:(s + 1) * R * 4 * H], R, 4 * H, 4 * H,
                         db[(s >> 1) * 4 * H:((s >> 1) + 1) * 4 * H])
            for m, mod in enumerate((lstm.lstm_re, lstm.lstm_im)):
                self._grad(getattr(mod, "weight_hh_l%d" % layer), d_hh[m])
                self._grad(getattr(mod, "weight_ih_l%d" % layer), d_ih[m])
                self._grad(getattr(mod, "bias_ih_l%d" % layer), db[m * 4 * H:(m + 1) * 4 * H])
                self._grad(getattr(mod, "bias_hh_l%d" % layer), db[m * 4 * H:(m + 1) * 4 * H].clone())
            if layer == 1:
                dH = gin                                               # [4][R][H]: gradient of h0
            else:
                g_top = gin                                            # planes [F][R][2ch]: gradient of the encoder output
        return g_top

    def _bptt(self, layer, P, cst, dH, pk_hh, NB, T, H):
        """Back-propagation through time of one nn.LSTM layer (4 streams x NB rows): 2T - 1 dependent launches.  The
        loop is launch-bound (a step is a few microseconds of GPU work), so it is captured once per (layer, shape) into
        a CUDA graph over static buffers and replayed; the first call of a shape runs eagerly (lazy CUDA
        initialisation must not happen under capture).  Returns dP fp32 [4][R][4H] (static, pad rows zero)."""
        dev = P.device
        R = NB * (T + 1)
        key = (layer, NB, T, H, str(dev))
        st = self._bptt_state.get(key)
        tc = pk_hh.tc()
        if st is None:
            st = self._bptt_state[key] = {
                "P": torch.empty_like(P), "cst": torch.empty_like(cst), "dH": torch.empty_like(dH),
                "dP": _zeros(4 * R * 4 * H, dev), "dP_step": torch.empty(2 * 4 * NB * 4 * H, dtype=torch.bfloat16, device=dev),
                "dc": torch.empty(4 * NB * H, dtype=torch.float32, device=dev),
                "dh_rec": torch.empty(pk_hh.out_planes // 2 * 4 * NB * H, dtype=torch.float32, device=dev),
                "wt": torch.empty_like(tc["wt"]), "units": tc["units"].clone(), "taps": tc["taps"].clone(),
                "bias": pk_hh.bias.clone(), "graph": None, "calls": 0}
        st["P"].copy_(P)
        st["cst"].copy_(cst)
        st["dH"].copy_(dH)
        st["wt"].copy_(tc["wt"])                    # the weights are re-packed every optimiser step: static copy

        parts = pk_hh.out_planes // 2

        def loop():
            dh = None
            for t in range(T - 1, -1, -1):
                lib.call("idv_lstm_cell_bwd_step", st["P"], st["cst"], st["dH"], dh, st["dc"], NB, T, H, t,
                         1 if dh is None else 0, parts, st["dP"], st["dP_step"])
                if t:
                    # dP_step [4 streams][NB][4H] read as [2 modules][2 NB rows][4H]: one unit per (module, K slice)
                    lib.call("idv_tapgemm_tc", st["dP_step"], 4 * H, 2, None, 0, 0, 2 * NB, 0, st["wt"], tc["kc_max"],
                             tc["n_slots"], st["bias"], pk_hh.N, st["units"], st["taps"], pk_hh.n_units, st["dh_rec"],
                             pk_hh.out_ld, 2 * NB * pk_hh.out_ld, pk_hh.out_planes * 2 * NB * pk_hh.out_ld, 0, 0, 0.0, 0)
                    dh = st["dh_rec"]
        st["calls"] += 1
        if not BPTT_GRAPH[0] or st["calls"] == 1 or not P.is_cuda:
            loop()
        else:
            if st["graph"] is None:
                n0 = lib.LAUNCHES[0]
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    loop()
                st["graph"], st["launches"] = g, lib.LAUNCHES[0] - n0
                lib.LAUNCHES[0] = n0
            st["graph"].replay()
            lib.LAUNCHES[0] += st["launches"]
        return st["dP"]

    # ---- conv + ComplexBatchNormal + PReLU ----
    def _conv_backward(self, i, g):
        """g: fp32 planes = dL/d(output of encoder layer i).  Returns dL/d(input planes) (None for layer 0)."""
        sv = self.saved["layers"][i]
        e = self.enc.encoders[i]
        raw, bn = sv["raw"], e.bn
        NB, C, F, T = raw.NB, raw.C, raw.F, raw.T
        dev = g.device
        R = NB * (T + 1)
        acc = torch.empty(C * 8, dtype=torch.float64, device=dev)
        lib.call("idv_cbn_bwd_reduce", raw.data, 0, g, 0, NB, C, F, T, sv["stats"], sv["zb"], sv["slope"], acc, raw.Tv)
        coef = torch.empty(C * 10, dtype=torch.float32, device=dev)
        dpar = [_zeros(C, dev) for _ in range(5)]
        dslope = _zeros(1, dev, torch.float64)
        lib.call("idv_cbn_bwd_finalize", acc, float(NB * F * raw.Tv), C, sv["stats"], bn.gamma_rr.detach(),
                 bn.gamma_ri.detach(), bn.gamma_ii.detach(), coef, dpar[0], dpar[1], dpar[2], dpar[3], dpar[4], dslope)
        for prm, val in zip((bn.gamma_rr, bn.gamma_ri, bn.gamma_ii, bn.beta_r, bn.beta_i), dpar):
            self._grad(prm, val)
        self._grad(e.prelu.weight, dslope.to(torch.float32))
        first = i == 0
        dy = torch.empty(raw.data.numel() * (1 if first else 2), dtype=torch.float32 if first else torch.bfloat16,
                         device=dev)
        lib.call("idv_cbn_bwd_apply", raw.data, 0, g, 0, NB, C, F, T, sv["stats"], sv["zb"], coef, sv["slope"], dy,
                 0 if first else 1, raw.Tv)
        c = e.conv
        wr, wi = c.conv_re.weight, c.conv_im.weight
        cout, cin, kh, kw = wr.shape
        # a bias in front of a batch-statistics normalisation has an exactly zero gradient
        self._grad(c.conv_re.bias, torch.zeros_like(c.conv_re.bias))
        self._grad(c.conv_im.bias, torch.zeros_like(c.conv_im.bias))
        if first:
            stft_x = self.saved["stft_x"]
            dW = torch.empty(20 * 2 * cout, dtype=torch.float32, device=dev)
            lib.call("idv_enc0_wgrad", stft_x, dy, NB, stft_x.shape[1], T, cout, 1, dW)
            d = dW.view(10, 2, 2 * cout)
            d_re = d[:, 0, :cout] + d[:, 1, cout:]
            d_im = d[:, 0, cout:] - d[:, 1, :cout]
            f = lambda m: m.reshape(kh, kw, 1, cout).permute(3, 2, 0, 1).contiguous()
            self._grad(wr, f(d_re))
            self._grad(wi, f(d_im))
            return None
        xin = sv["x"]
        kh_, sf, pf = c._geometry()
        pt = c._time_geometry()[0]
        key = ("dgrad", i, wr._version, wi._version)
        if key not in self._packs:
            self._packs[key] = pack.pack_conv_dgrad(wr, wi, xin.F, sf, pf, pt, dev)
        dyp = Planes(dy, NB, C, F, T, split=True)
        gin = ops.tapgemm(self._packs[key], dyp, None, NB, T, zero_pad_rows=True, out_split=False)
        # weight gradient: K = rows x output planes, split over enough units to give every SM a tile
        rp = _rpad(R)
        rows, N = 2 * round8(cout), xin.Cp
        tiles = kh * kw * ((rows + 127) // 128) * max(1, N // 256)
        groups = max(1, min(F, -(-160 // tiles)))
        tk = ("wgrad", i, rp, groups)
        if tk not in self._packs:
            self._packs[tk] = pack.wgrad_conv_tables(xin.F, F, kh, kw, sf, pf, pt, rp, groups, dev)
        u, tp, n, groups = self._packs[tk]
        dyT0 = _transpose_split(dy, True, F, R, raw.Cp, 0)
        dyT1 = _transpose_split(dy, True, F, R, raw.Cp, 1)
        xT = _transpose_split(xin.data, True, xin.F, R, xin.Cp, 0)
        o = _wgrad_gemm(dyT0, dyT1, F, rows, xT, xin.F, N, rp, u, tp, n)             # [taps*groups][rows][N]
        dwt = o.view(kh * kw, groups, rows, N).sum(1)
        d_re, d_im = pack.unfold_conv_wgrad(dwt, kh, kw, cin, cout)
        self._grad(wr, d_re)
        self._grad(wi, d_im)
        return gin


# Callables(list of parameters) invoked when a model's backward node has written ALL of its parameter gradients for this
# backward pass: FlatAdam starts the all-reduce of that model's slice of the gradient bucket there, so the collective
# runs on the NCCL stream while the rest of the backward pass (the other model) still computes.
GRADS_FINAL_HOOKS = []


def _notify_grads_final(params):
    for h in list(GRADS_FINAL_HOOKS):
        h(params)


class _EncoderTrainFn(torch.autograd.Function):
    """Autograd plumbing: forward = EncoderTrainStep.forward, backward hands dL/dlatent to EncoderTrainStep.backward,
    which writes the parameter gradients itself (the parameters are inputs only so that autograd calls backward)."""

    @staticmethod
    def forward(ctx, step, x, *params):
        latent, stft_x, acts = step.forward(x)
        ctx.step = step
        ctx.n = len(params)
        step._aux = (stft_x, acts)
        # ordering token: a decoder that back-propagates into the skip tensors takes it as an input, so autograd runs
        # the decoder's backward (which hands the skip gradients to ``step``) before this node's
        return latent, torch.zeros((), device=latent.device)

    @staticmethod
    def backward(ctx, dlatent, dtoken):
        ctx.step.backward(dlatent.contiguous())
        _notify_grads_final([p for p in ctx.step.enc.parameters() if p.requires_grad])
        return (None, None) + (None,) * ctx.n


class _ReparamFn(torch.autograd.Function):
    """z = reparameterization(mu, log sigma, delta) (model/pvae_module.py:L2177-2231, num_samples = 1) with the
    gradient of the latent from idv_reparam_bwd (end-to-end training: the decoder's loss reaches the encoder through z)."""

    @staticmethod
    def forward(ctx, latent, ch0, zdim, eps_r, eps_i):
        lat = lib.require_f32_cuda(latent.detach(), "latent")
        z = ops.reparam(lat, ch0, zdim, 1, eps_r, eps_i, 0, 0)
        ctx.save_for_backward(lat, eps_r, eps_i)
        ctx.ch0, ctx.zdim = ch0, zdim
        return z

    @staticmethod
    def backward(ctx, dz):
        lat, eps_r, eps_i = ctx.saved_tensors
        NB, T, Htot, _ = lat.shape
        dlat = torch.zeros_like(lat)
        lib.call("idv_reparam_bwd", lat, NB, T, Htot, ctx.ch0, ctx.zdim, eps_r, eps_i,
                 lib.require_f32_cuda(dz, "gradient of z"), dlat)
        return dlat, None, None, None, None


def reparam_train(latent, ch0, zdim, eps_r, eps_i):
    """Differentiable z for the training step; eps drawn with torch.randn when not supplied (like the reference's
    randn_like: the backward needs the same draw)."""
    NB, T = latent.shape[0], latent.shape[1]
    if eps_r is None:
        eps_r = torch.randn((NB, 1, T, zdim), dtype=torch.float32, device=latent.device)
        eps_i = torch.randn((NB, 1, T, zdim), dtype=torch.float32, device=latent.device)
    eps_r, eps_i = lib.require_f32_cuda(eps_r, "eps_r"), lib.require_f32_cuda(eps_i, "eps_i")
    return _ReparamFn.apply(latent, ch0, zdim, eps_r, eps_i)


def encoder_train_forward(enc, x):
    """latent (with a grad_fn), stft_x, encoder activation planes, ordering token, train step of one train-mode forward."""
    step = getattr(enc, "_train_step", None)
    if step is None:
        step = enc._train_step = EncoderTrainStep(enc)
    params = [p for p in enc.parameters() if p.requires_grad]
    latent, token = _EncoderTrainFn.apply(step, x, *params)
    stft_x, acts = step._aux
    return latent, stft_x, acts, token, step


# ------------------------------------------------------------------------------------------------------------------
# decoder (phase 2 of the reference: i_dccrn_vae/nsvae_dccrn/train_second_phase_decoder.py:L376-433 - frozen NSVAE
# encoder, decoder train=True, SI-SNR / spectral reconstruction losses, backward through iSTFT, reconstruction head,
# ComplexBatchNormal(train) + PReLU, complex transposed convs with skip concat, ComplexDense)
# ------------------------------------------------------------------------------------------------------------------
class DecoderTrainStep:
    """Train-mode forward of a VAE decoder keeping what the backward needs, and the backward from the gradients of
    (recon_sig, predict) into ``.grad`` of every decoder parameter; optionally also the gradients of ``z`` and of the
    skip tensors (end-to-end step).  Same kernel plan as the encoder: data gradients of the
    transposed convs = the tcgen05 tap-GEMM with transposed weights (the strided conv the layer is the adjoint of),
    weight gradients = the tap-GEMM over the row dimension on transposed copies; the Cout = 1 last layer, whose output
    gradient has only 2 channels, has SIMT kernels (idv_dec5_dgrad / idv_dec5_wgrad)."""

    def __init__(self, dec):
        if not dec.causal:
            raise NotImplementedError("the backward pass is built for the causal network (model/causal_netconfig.py)")
        if not ops.use_split():
            raise RuntimeError("training runs on the tensor-core path (IDV_GEMM=tc)")
        self.dec = dec
        self.saved = None
        self._packs = {}

    # ---------------------------------------------------------------------------------------------- forward
    def forward(self, stft_x, z, skips, C, F, mask):
        with pack.on_device():
            return self._forward(stft_x, z, skips, C, F, mask)

    def _forward(self, stft_x, z, skips, C, F, mask):
        """z (B*S, T, zdim, 2), skips {layer: Planes of B utterances, or "self"}.  Returns recon_sig (B*S, L), predict
        (B*S, F, T, 2).  With num_samples = S > 1 (train_second_phase_decoder.sh:L6 runs --num_samples 2) the S samples of
        an utterance are rows b*S + s of ONE batch, exactly like the reference (model/pvae_module.py:L2550-2567): the
        batch statistics of every ComplexBatchNormal span all B*S rows, the skip tensors and the noisy STFT of the mask
        head are repeated per sample (row b*S + s reads utterance b).  "self" = the layer's own input as its skip tensor
        (pvae_dccrn_decoder_prob_skip with skip_prob = 2, L1757-1758)."""
        dec = self.dec
        z = lib.require_f32_cuda(z, "z")
        B, T, zdim, _ = z.shape                     # B = utterances x samples from here on
        S = dec.num_samples
        if B % S:
            raise RuntimeError("z batch %d is not a multiple of num_samples %d" % (B, S))
        dev = z.device
        n = len(dec.decoders)
        if S > 1:
            skips = {i: (sk if isinstance(sk, str) else ops.repeat_planes(sk, S)) for i, sk in skips.items()}
            if mask:
                stft_x = lib.require_f32_cuda(stft_x, "stft_x").repeat_interleave(S, 0)
        zp = ops.z_to_planes(z, B, 1, 0, split=True, t_alloc=T)
        p = dec.dense.forward_planes(zp, C, F)
        sv = {"zp": zp, "dense_out": p, "layers": [], "B": B, "T": T, "C": C, "F": F, "mask": mask, "S": S}
        _skip = lambda i, cur: cur if isinstance(skips.get(i), str) else skips.get(i)
        for i in range(n - 1):
            d = dec.decoders[i]
            raw = d.forward_planes(p, _skip(i, p), True, raw_only=True)
            Cc = raw.C
            acc = torch.empty(Cc * 5, dtype=torch.float64, device=dev)
            lib.call("idv_cbn_stats_planes", raw.data, 0, raw.NB, Cc, raw.F, raw.T, acc, raw.Tv)
            stats = torch.empty(Cc * 5, dtype=torch.float32, device=dev)
            zb = ops._cbn_finalize(d.bn, acc, raw.NB * raw.F * raw.Tv, dev, stats)
            slope = d._slope()
            act = torch.empty(2 * raw.data.numel(), dtype=torch.bfloat16, device=dev)
            lib.call("idv_cbn_apply_planes", raw.data, 0, raw.NB, Cc, raw.F, raw.T, zb, 1, slope, raw.Tv, act, 1)
            a = Planes(act, raw.NB, Cc, raw.F, raw.T, split=True, Tv=raw.Tv)
            sv["layers"].append({"x": p, "skip": _skip(i, p), "raw": raw, "stats": stats, "zb": zb, "slope": slope})
            p = a
        d = dec.decoders[n - 1]
        n_bins = 2 * p.F - 1
        raw5 = torch.empty((B, n_bins, T, 2), dtype=torch.float32, device=dev)
        d.forward_head(p, _skip(n - 1, p), False, None, raw5, 1, 0, train=True, raw_only=True)
        inner = n_bins * T
        acc = torch.empty(5, dtype=torch.float64, device=dev)
        lib.call("idv_cbn_stats_user", raw5, B, 1, inner, acc)
        stats = torch.empty(5, dtype=torch.float32, device=dev)
        zb = ops._cbn_finalize(d.bn, acc, B * inner, dev, stats)
        predict = torch.empty_like(raw5)
        lib.call("idv_cbn_eval_user", raw5, B, 1, inner, zb, predict)
        if mask:
            stft_x = lib.require_f32_cuda(stft_x, "stft_x")
        lib.call("idv_head_user", predict, inner, B, float(d._slope()), 1 if mask else 0, stft_x if mask else None, 1)
        sv["head"] = {"x": p, "skip": _skip(n - 1, p), "raw": raw5, "stats": stats, "zb": zb, "slope": d._slope(),
                      "stft_x": stft_x if mask else None}
        self.saved = sv
        recon_sig = dec.istft.forward_ri(predict)
        return recon_sig, predict

    # ---------------------------------------------------------------------------------------------- backward
    def _grad(self, param, value):
        value = value.to(param.dtype).reshape(param.shape)
        param.grad = value if param.grad is None else param.grad + value

    def backward(self, d_sig, d_pred=None, want_dz=False, want_dskip=False):
        """d_sig (B, L) / d_pred (B, F, T, 2): gradients of recon_sig / predict (either may be None).  Accumulates
        the parameter gradients; returns (dz or None, {layer: gradient planes of the skip tensor})."""
        if self.saved is None:
            raise RuntimeError("backward() without a train-mode forward")
        with pack.on_device():
            out = self._backward(d_sig, d_pred, want_dz, want_dskip)
        self.saved = None
        return out

    def _bn_backward(self, d, raw_data, g, NB, C, F, T, stats, zb, slope, dy_split):
        """ComplexBatchNormal(train) + PReLU backward of decoder block ``d`` on planes; returns dy (gradient of the raw
        transposed-conv output) and writes the parameter gradients."""
        dev = g.device
        bn = d.bn
        acc = torch.empty(C * 8, dtype=torch.float64, device=dev)
        lib.call("idv_cbn_bwd_reduce", raw_data, 0, g, 0, NB, C, F, T, stats, zb, slope, acc, T)
        coef = torch.empty(C * 10, dtype=torch.float32, device=dev)
        dpar = [_zeros(C, dev) for _ in range(5)]
        dslope = _zeros(1, dev, torch.float64)
        lib.call("idv_cbn_bwd_finalize", acc, float(NB * F * T), C, stats, bn.gamma_rr.detach(), bn.gamma_ri.detach(),
                 bn.gamma_ii.detach(), coef, dpar[0], dpar[1], dpar[2], dpar[3], dpar[4], dslope)
        for prm, val in zip((bn.gamma_rr, bn.gamma_ri, bn.gamma_ii, bn.beta_r, bn.beta_i), dpar):
            self._grad(prm, val)
        self._grad(d.prelu.weight, dslope.to(torch.float32))
        dy = torch.empty(raw_data.numel() * (2 if dy_split else 1), dtype=torch.bfloat16 if dy_split else torch.float32,
                         device=dev)
        lib.call("idv_cbn_bwd_apply", raw_data, 0, g, 0, NB, C, F, T, stats, zb, coef, slope, dy, 1 if dy_split else 0, T)
        t = d.transconv
        # a bias in front of a batch-statistics normalisation has an exactly zero gradient
        self._grad(t.tconv_re.bias, torch.zeros_like(t.tconv_re.bias))
        self._grad(t.tconv_im.bias, torch.zeros_like(t.tconv_im.bias))
        return dy

    def _backward(self, d_sig, d_pred, want_dz, want_dskip):
        sv, dec = self.saved, self.dec
        B, T = sv["B"], sv["T"]
        hd = sv["head"]
        dev = hd["raw"].device
        n = len(dec.decoders)
        n_bins = hd["raw"].shape[1]
        dskips = {}
        # ---- iSTFT adjoint: overlap-add adjoint, then the DFT GEMM with the transposed synthesis basis
        ist = dec.istft
        if d_sig is not None:
            d_sig = lib.require_f32_cuda(d_sig, "gradient of recon_sig")
            key = ("istft_adj", str(dev))
            if key not in self._packs:
                self._packs[key] = pack.pack_istft_adjoint_tc(ist.n_fft, ist.win_length, dev)
            hp = self._packs[key]
            R0 = B * T
            dframes = torch.empty(R0 * hp["kpad"], dtype=torch.float32, device=dev)
            lib.call("idv_ola_bwd", d_sig, hp["wsq"], B, T, ist.n_fft, ist.hop_length, ist.win_length, hp["kpad"], dframes)
            dfs = _to_split(dframes)
            drows = torch.empty(R0 * hp["N"], dtype=torch.float32, device=dev)
            lib.call("idv_tapgemm_tc", dfs, hp["kpad"], 1, None, 0, 0, R0, 0, hp["wt"], hp["kc_max"], 1, hp["bias"],
                     hp["N"], hp["units"], hp["taps"], 1, drows, hp["N"], R0 * hp["N"], 0, 0, 0, 0.0, 0)
            ld = hp["N"]
        else:
            ld = 2 * n_bins
            drows = _zeros(B * T * ld, dev)
        if d_pred is not None:
            d_pred = lib.require_f32_cuda(d_pred, "gradient of predict")
        # ---- reconstruction head + last layer (Cout = 1)
        d5 = dec.decoders[n - 1]
        x5, sk5 = hd["x"], hd["skip"]
        R = B * (T + 1)
        yp = torch.empty(n_bins * R * 16, dtype=torch.float32, device=dev)
        gp = torch.empty(n_bins * R * 16, dtype=torch.float32, device=dev)
        lib.call("idv_head_bwd", hd["raw"], hd["zb"], float(hd["slope"]), 1 if sv["mask"] else 0, hd["stft_x"], drows, ld,
                 d_pred, B, n_bins, T, yp, gp)
        del drows
        dy5 = self._bn_backward(d5, yp, gp, B, 1, n_bins, T, hd["stats"], hd["zb"], float(hd["slope"]), False)
        del yp, gp
        t5 = d5.transconv
        c_skip = sk5.C if sk5 is not None else 0
        w10, _, _ = pack.pack_dec5(t5.tconv_re.weight, t5.tconv_re.bias, t5.tconv_im.weight, t5.tconv_im.bias, None, None,
                                   x5.C, c_skip, dev)
        ktot = x5.Cp + (sk5.Cp if sk5 is not None else 0)
        dW = _zeros(10 * ktot * 2, dev)
        g = torch.empty(x5.F * R * x5.Cp, dtype=torch.float32, device=dev)
        lib.call("idv_dec5_dgrad", dy5, w10, ktot, 0, x5.Cp, x5.F, B, T, g)
        lib.call("idv_dec5_wgrad", x5.data, 1, dy5, ktot, 0, x5.Cp, x5.F, B, T, dW)
        if sk5 is not None:
            lib.call("idv_dec5_wgrad", sk5.data, 1, dy5, ktot, x5.Cp, sk5.Cp, sk5.F, B, T, dW)
            if want_dskip or sk5 is x5:
                gs = torch.empty(sk5.F * R * sk5.Cp, dtype=torch.float32, device=dev)
                lib.call("idv_dec5_dgrad", dy5, w10, ktot, x5.Cp, sk5.Cp, sk5.F, B, T, gs)
                if sk5 is x5:                        # the layer's input was also its skip tensor: both gradients are dL/dx
                    lib.call("idv_axpy", g, gs, 1.0, g.numel())
                else:
                    dskips[n - 1] = self._sum_samples(gs, sk5)
        del dy5
        d_re, d_im = pack.unfold_dec5_wgrad(dW.view(10, ktot, 2), x5.C, c_skip, t5.tconv_re.weight.shape[0])
        self._grad(t5.tconv_re.weight, d_re)
        self._grad(t5.tconv_im.weight, d_im)
        # ---- layers n-2 .. 0
        for i in reversed(range(n - 1)):
            g, gs = self._tconv_backward(i, g, want_dskip)
            if gs is not None:
                dskips[i] = gs
        # ---- ComplexDense: g = gradient of its output planes [F][R][2 ch_c]
        dz = self._dense_backward(g, want_dz)
        return dz, dskips

    def _tconv_backward(self, i, g, want_dskip):
        sv = self.saved["layers"][i]
        d = self.dec.decoders[i]
        raw, xin, skip = sv["raw"], sv["x"], sv["skip"]
        NB, C, F, T = raw.NB, raw.C, raw.F, raw.T
        dev = g.device
        R = NB * (T + 1)
        dy = self._bn_backward(d, raw.data, g, NB, C, F, T, sv["stats"], sv["zb"], sv["slope"], True)
        t = d.transconv
        wr, wi = t.tconv_re.weight, t.tconv_im.weight
        cin_tot, cout, kh, kw = wr.shape
        kh_, sf, pf = t._geometry()
        dyp = Planes(dy, NB, C, F, T, split=True)
        c_skip = skip.C if skip is not None else 0
        key = ("dgrad", i, wr._version, wi._version)
        if key not in self._packs:
            self._packs[key] = pack.pack_convT_dgrad(wr, wi, 0, xin.C, xin.F, sf, pf, dev)
        gin = ops.tapgemm(self._packs[key], dyp, None, NB, T, zero_pad_rows=True, out_split=False)
        gskip = None
        if skip is not None and (want_dskip or skip is xin):
            key = ("dgrad_skip", i, wr._version, wi._version)
            if key not in self._packs:
                self._packs[key] = pack.pack_convT_dgrad(wr, wi, xin.C, c_skip, xin.F, sf, pf, dev)
            gskip = ops.tapgemm(self._packs[key], dyp, None, NB, T, zero_pad_rows=True, out_split=False)
            if skip is xin:                          # self skip: both halves of the concatenated input are the same tensor
                lib.call("idv_axpy", gin, gskip, 1.0, gin.numel())
                gskip = None
            else:
                gskip = self._sum_samples(gskip, skip)
        # weight gradients: K = rows x output planes, one GEMM per source (weight rows [p | skip] like torch.cat)
        rp = _rpad(R)
        rows = 2 * round8(cout)
        dyT0 = _transpose_split(dy, True, F, R, raw.Cp, 0)
        dyT1 = _transpose_split(dy, True, F, R, raw.Cp, 1)
        d_re, d_im = torch.zeros_like(wr), torch.zeros_like(wi)
        for (src, c0) in ((xin, 0), (skip, xin.C)):
            if src is None:
                continue
            N = src.Cp
            tiles = kh * kw * ((rows + 127) // 128) * max(1, N // 256)
            groups = max(1, min(F, -(-160 // tiles)))
            tk = ("wgrad", i, rp, groups, src.F)
            if tk not in self._packs:
                self._packs[tk] = pack.wgrad_conv_tables(src.F, F, kh, kw, sf, pf, 0, rp, groups, dev, transposed=True)
            u, tp, nu, ng = self._packs[tk]
            xT = _transpose_split(src.data, True, src.F, R, src.Cp, 0)
            o = _wgrad_gemm(dyT0, dyT1, F, rows, xT, src.F, N, rp, u, tp, nu)
            dwt = o.view(kh * kw, ng, rows, N).sum(1)
            g_re, g_im = pack.unfold_conv_wgrad(dwt, kh, kw, src.C, cout, transposed=True)
            d_re[c0:c0 + src.C] = g_re
            d_im[c0:c0 + src.C] = g_im
            del xT, o
        self._grad(wr, d_re)
        self._grad(wi, d_im)
        return gin, gskip

    def _sum_samples(self, g, planes):
        """Gradient planes [F][B*S*(T+1)][Cp] of a skip tensor that was repeated per sample -> [F][B*(T+1)][Cp]: the
        encoder's skip tensor of utterance b feeds rows b*S .. b*S + S - 1."""
        S = self.saved["S"]
        if S == 1:
            return g
        Tp = planes.T + 1
        return g.view(planes.F, planes.NB // S, S, Tp, planes.Cp).sum(2).reshape(-1).contiguous()

    def _dense_backward(self, g, want_dz):
        sv, dense = self.saved, self.dec.dense
        zp, p0 = sv["zp"], sv["dense_out"]
        NB, T, C, F = p0.NB, p0.T, p0.C, p0.F
        dev = g.device
        R = NB * (T + 1)
        rp = _rpad(R)
        ch_c, ch_z = round8(C), round8(zp.C)
        zdim = zp.C
        wr, wi = dense.linear_read.weight, dense.linear_imag.weight
        # bias gradients: column sums of every output plane
        db = _zeros(F * 2 * ch_c, dev)
        for f in range(F):
            lib.call("idv_colsum_add", g[f * R * 2 * ch_c:(f + 1) * R * 2 * ch_c], R, 2 * ch_c, 2 * ch_c,
                     db[f * 2 * ch_c:(f + 1) * 2 * ch_c])
        dbv = db.view(F, 2, ch_c)[:, :, :C]
        self._grad(dense.linear_read.bias, dbv[:, 0].t().reshape(-1))
        self._grad(dense.linear_imag.bias, dbv[:, 1].t().reshape(-1))
        gT = _transpose_split(g, False, F, R, 2 * ch_c, 0)
        zT = _transpose_split(zp.data, True, 1, R, 2 * ch_z, 0)
        tk = ("wg_dense", rp, F)
        if tk not in self._packs:
            self._packs[tk] = pack.wgrad_dense_tables(F, rp, dev)
        u, tp, nu = self._packs[tk]
        o = _wgrad_gemm(gT, None, F, 2 * ch_c, zT, 2, ch_z, rp, u, tp, nu).view(F, 2, 2, ch_c, ch_z)   # [f][part][half]
        self._grad(wr, o[:, 0, 0, :C, :zdim].permute(1, 0, 2).reshape(C * F, zdim))
        self._grad(wi, o[:, 1, 1, :C, :zdim].permute(1, 0, 2).reshape(C * F, zdim))
        if not want_dz:
            return None
        key = ("dgrad_dense", wr._version, wi._version)
        if key not in self._packs:
            self._packs[key] = pack.pack_dense_dgrad(wr, wi, C, F, dev)
        gs = Planes(_to_split(g), NB, C, F, T, split=True)
        dzp = ops.tapgemm(self._packs[key], gs, None, NB, T, zero_pad_rows=True, out_split=False)    # [1][R][2 ch_z]
        dz = dzp.view(NB, T + 1, 2, ch_z)[:, 1:, :, :zdim].permute(0, 1, 3, 2).contiguous()           # (B, T, zdim, 2)
        return dz


class _DecoderTrainFn(torch.autograd.Function):
    """Autograd plumbing of the decoder: forward = DecoderTrainStep.forward; backward receives the gradients of
    (recon_sig, predict) and runs DecoderTrainStep.backward, which writes the parameter gradients itself.  ``token`` is
    the encoder's ordering token (a scalar produced by the encoder's autograd node) when the skip tensors need
    gradients: autograd then runs this node before the encoder's, and the skip gradients are handed over directly."""

    @staticmethod
    def forward(ctx, step, stft_x, z, token, skips, C, F, mask, enc_step, *params):
        recon_sig, predict = step.forward(stft_x, z, skips, C, F, mask)
        ctx.step, ctx.enc_step, ctx.n, ctx.dev = step, enc_step, len(params), z.device
        ctx.set_materialize_grads(False)
        return recon_sig, predict

    @staticmethod
    def backward(ctx, d_sig, d_pred):
        want_dz = ctx.needs_input_grad[2]
        want_dskip = ctx.enc_step is not None and ctx.needs_input_grad[3]
        dz, dskips = ctx.step.backward(None if d_sig is None else d_sig.contiguous(),
                                       None if d_pred is None else d_pred.contiguous(), want_dz, want_dskip)
        if want_dskip:
            ctx.enc_step.add_skip_grads(dskips, len(ctx.step.dec.decoders))
        _notify_grads_final([p for p in ctx.step.dec.parameters() if p.requires_grad])
        token_grad = torch.zeros((), device=ctx.dev) if ctx.needs_input_grad[3] else None
        return (None, None, dz, token_grad, None, None, None, None, None) + (None,) * ctx.n


def decoder_train_forward(dec, stft_x, z, skiper, skips, C, F, mask):
    """(recon_sig, predict (B, F, T, 2)) with a grad_fn.  skiper: the encoder's SkipList (carries the encoder's train
    step / ordering token when the encoder ran under autograd), skips: {layer: Planes}."""
    step = getattr(dec, "_train_step", None)
    if step is None:
        step = dec._train_step = DecoderTrainStep(dec)
    params = [p for p in dec.parameters() if p.requires_grad]
    token = getattr(skiper, "grad_token", None) if skips else None
    enc_step = getattr(skiper, "train_step", None) if token is not None else None
    if token is None:
        token = torch.zeros((), device=z.device)
    return _DecoderTrainFn.apply(step, stft_x, z, token, skips, C, F, mask, enc_step, *params)


# ------------------------------------------------------------------------------------------------------------------
# optimiser + gradient all-reduce (data-parallel training: SURVEY §8(e))
# ------------------------------------------------------------------------------------------------------------------
class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics (train_nsvae.py:L200: lr from the config,
    weight_decay = 0.001) on ONE flat fp32 buffer: the parameters that receive gradients are re-pointed into a flat
    tensor, their gradients are gathered into a flat bucket (one NCCL all-reduce per step when a process group is
    given: mean over ranks, DDP semantics), the update is one ``idv_adam_step`` launch per run of parameters that share
    a step count (normally one).

    It IS a ``torch.optim.Optimizer``: ``param_groups`` (one group; ``lr`` is read from it every step, so torch's lr
    schedulers drive it), ``zero_grad``, and ``state_dict()`` / ``load_state_dict()`` in torch.optim.Adam's own format
    (per parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so the reference's checkpoints of optimiser state
    (train_nsvae.py:L318-330) load into it and its checkpoints load into ``torch.optim.Adam``.  Like torch.optim.Adam a
    parameter whose ``.grad`` is None in a step is skipped in that step (no update, no moment decay, no step count):
    the flat layout is rebuilt when the set of parameters with gradients changes."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None,
                 world_size=1, overlap=True):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FlatAdam holds one parameter group")
        self.group, self.world = process_group, world_size
        self.flat = None
        self.live = []                 # parameters in the flat layout, in param_groups order
        # overlapped all-reduce: slices of the bucket whose gradients a backward node declared final (GRADS_FINAL_HOOKS)
        # are reduced asynchronously while the backward pass continues; step() reduces what is left and waits
        self.overlap = bool(overlap) and process_group is not None and world_size > 1
        self._inflight = []            # (begin, end, work handle, [(param, grad tensor, version)])
        self.overlapped_elements = 0   # bucket elements whose all-reduce was started during a backward pass (diagnostic)
        if self.overlap:
            import weakref
            ref = weakref.ref(self)

            def hook(params, _ref=ref):
                me = _ref()
                if me is None:
                    GRADS_FINAL_HOOKS.remove(hook)
                else:
                    me._on_grads_final(params)
            GRADS_FINAL_HOOKS.append(hook)
        self._steps = {}               # id(param) -> number of updates applied
        self._pending = None           # (m, v) per parameter from load_state_dict, applied at the next layout build

    # ---- convenience views of the single group (the round-1 attribute names)
    @property
    def params(self):
        return self.param_groups[0]["params"]

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        self.param_groups[0]["lr"] = v

    @property
    def step_count(self):
        return max(self._steps.values()) if self._steps else 0

    def _moments(self):
        """{id(param): (exp_avg, exp_avg_sq)} views of the current layout (+ moments loaded but not laid out yet)."""
        out = dict(self._pending or {})
        if self.flat is not None:
            for p, (off, k) in zip(self.live, self.views):
                out[id(p)] = (self.m[off:off + k].view(p.shape), self.v[off:off + k].view(p.shape))
        return out

    def _build(self, live):
        old = self._moments()
        n = sum(p.numel() for p in live)
        dev = live[0].device
        flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.gflat = torch.zeros(n, dtype=torch.float32, device=dev)
        m = torch.zeros(n, dtype=torch.float32, device=dev)
        v = torch.zeros(n, dtype=torch.float32, device=dev)
        off, views = 0, []
        for p in live:
            k = p.numel()
            flat[off:off + k].copy_(p.data.reshape(-1))
            if id(p) in old:
                m[off:off + k].copy_(old[id(p)][0].reshape(-1))
                v[off:off + k].copy_(old[id(p)][1].reshape(-1))
            views.append((off, k))
            off += k
        in_new = {id(p) for p in live}
        for p in self.live:                          # parameters that leave the layout get storage of their own back
            if id(p) not in in_new:
                p.data = p.data.clone()
        # moments of parameters outside the new layout (left it, or loaded and never laid out) wait in _pending
        pending = {k: (mv[0].clone(), mv[1].clone()) for k, mv in old.items() if k not in in_new}
        for p, (off, k) in zip(live, views):
            p.data = flat[off:off + k].view(p.shape)
        self.flat, self.m, self.v, self.live, self.views = flat, m, v, list(live), views
        self._pending = pending or None

    @torch.no_grad()
    def _on_grads_final(self, params):
        """A backward node finished writing the gradients of ``params``: if they form one contiguous run of the flat
        layout, gather them into the bucket and start its all-reduce now (async, on the collective's own stream)."""
        if self.flat is None or not self.live:
            return
        index = {id(p): i for i, p in enumerate(self.live)}
        idx = sorted(index[id(p)] for p in params if id(p) in index and p.grad is not None)
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            return
        a, b = self.views[idx[0]][0], self.views[idx[-1]][0] + self.views[idx[-1]][1]
        if any(not (e <= a or s >= b) for s, e, _, _ in self._inflight):
            return
        torch.cat([self.live[i].grad.reshape(-1) for i in idx], out=self.gflat[a:b])       # one gather for the run
        stamp = [(self.live[i], self.live[i].grad, self.live[i].grad._version) for i in idx]
        work = torch.distributed.all_reduce(self.gflat[a:b], group=self.group, async_op=True)
        self._inflight.append((a, b, work, stamp))
        self.overlapped_elements += b - a

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return loss
        rebuilt = self.flat is None or len(live) != len(self.live) or any(a is not b for a, b in zip(live, self.live))
        inflight, self._inflight = self._inflight, []
        for _, _, work, _ in inflight:
            work.wait()
        # a slice reduced early is only valid if nothing touched those gradients afterwards (a module used twice in one
        # graph accumulates into .grad after its first backward node ran) and the layout is still the same
        valid = [(a, b) for a, b, _, stamp in inflight
                 if not rebuilt and all(p.grad is g and g._version == v for p, g, v in stamp)]
        if rebuilt:
            self._build(live)
        done = lambda off: any(a <= off < b for a, b in valid)
        i = 0
        while i < len(self.live):                    # one gather (torch.cat) per run of parameters still to be copied
            if done(self.views[i][0]):
                i += 1
                continue
            j = i
            while j + 1 < len(self.live) and not done(self.views[j + 1][0]):
                j += 1
            a, b = self.views[i][0], self.views[j][0] + self.views[j][1]
            torch.cat([p.grad.reshape(-1) for p in self.live[i:j + 1]], out=self.gflat[a:b])
            i = j + 1
        if self.group is not None and self.world > 1:
            # reduce the runs of the bucket that were not reduced during the backward pass
            n, pos = self.gflat.numel(), 0
            for a, b in sorted(valid) + [(n, n)]:
                if a > pos:
                    torch.distributed.all_reduce(self.gflat[pos:a], group=self.group)
                pos = max(pos, b)
            self.gflat.div_(self.world)
        g = self.param_groups[0]
        # one launch per run of consecutive parameters with the same step count (bias corrections differ otherwise)
        i = 0
        while i < len(self.live):
            st = self._steps.get(id(self.live[i]), 0)
            j = i
            while j + 1 < len(self.live) and self._steps.get(id(self.live[j + 1]), 0) == st:
                j += 1
            a, b = self.views[i][0], self.views[j][0] + self.views[j][1]
            lib.call("idv_adam_step", self.flat[a:b], self.gflat[a:b], self.m[a:b], self.v[a:b], b - a, float(g["lr"]),
                     float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), st + 1)
            i = j + 1
        for p in self.live:                          # the kernel wrote the parameters behind autograd's back:
            self._steps[id(p)] = self._steps.get(id(p), 0) + 1
            torch.autograd.graph.increment_version(p)    # bump the versions so the weight-pack caches rebuild
        return loss

    def state_dict(self):
        """torch.optim.Adam's format: {'state': {index: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}."""
        mom = self._moments()
        state = {}
        for i, p in enumerate(self.params):
            if id(p) in mom and self._steps.get(id(p), 0) > 0:
                state[i] = {"step": torch.tensor(float(self._steps[id(p)])), "exp_avg": mom[id(p)][0].clone(),
                            "exp_avg_sq": mom[id(p)][1].clone()}
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g.update(amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                 params=list(range(len(self.params))))
        return {"state": state, "param_groups": [g]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("optimizer state has %d group(s) / %d parameters, this optimiser 1 / %d"
                             % (len(groups), len(groups[0]["params"]), len(self.params)))
        if groups[0].get("amsgrad") or groups[0].get("maximize"):
            raise NotImplementedError("FlatAdam implements plain Adam (amsgrad / maximize off)")
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in groups[0]:
                self.param_groups[0][k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]
        index = {pid: i for i, pid in enumerate(groups[0]["params"])}
        pending, self._steps = {}, {}
        for pid, st in sd["state"].items():
            p = self.params[index[pid] if pid in index else int(pid)]
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError("optimizer state of parameter %s has shape %s, the parameter %s"
                                 % (pid, tuple(st["exp_avg"].shape), tuple(p.shape)))
            pending[id(p)] = (st["exp_avg"].detach().to(p.device, torch.float32).clone(),
                              st["exp_avg_sq"].detach().to(p.device, torch.float32).clone())
            self._steps[id(p)] = int(float(st["step"]))
        if self.flat is not None:                    # keep the current layout, overwrite its moments
            for p, (off, k) in zip(self.live, self.views):
                if id(p) in pending:
                    self.m[off:off + k].copy_(pending[id(p)][0].reshape(-1))
                    self.v[off:off + k].copy_(pending[id(p)][1].reshape(-1))
                else:
                    self.m[off:off + k].zero_()
                    self.v[off:off + k].zero_()
            pending = {k: v for k, v in pending.items() if all(id(p) != k for p in self.live)}
        self._pending = pending or None
