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
import torch

from . import lib, ops, pack
from .ops import Planes
from .pack import round8


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

    # ---------------------------------------------------------------------------------------------- forward
    def forward(self, x):
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
            # out of place and into the split format in one pass would need a second kernel variant: normalise in
            # fp32 (raw is kept for the backward pass), then split for the next layer's tensor-core GEMM
            act32 = torch.empty_like(raw.data)
            lib.call("idv_cbn_apply_planes", raw.data, 0, raw.NB, C, raw.F, raw.T, zb, 1, slope, raw.Tv, act32)
            act = _to_split(act32)
            del act32
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
        h0, h0s = ops.lstm_recurrent_tc(g, 4 * H, R * 8 * H, 8 * H, whh_tc[0], NB, T, H, want_f32=True, want_split=True)
        src = Planes(h0s, NB, H, 4, T, cp=H, split=True)
        g = ops.tapgemm(layers[1][0], src, None, NB, T, zero_pad_rows=False, out_split=False)
        h1, h1s = ops.lstm_recurrent_tc(g, 2 * R * 4 * H, R * 4 * H, 4 * H, whh_tc[1], NB, T, H, want_f32=True,
                                        want_split=True)
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
            pk_hh = self._lstm_pack(("hh", layer), lambda: pack.pack_lstm_dgrad(re, im, H, layer, "hh", dev))
            pk_ih = self._lstm_pack(("ih", layer), lambda: pack.pack_lstm_dgrad(re, im, H, layer, "ih", dev, top.C, top.F))
            # gate pre-activations of every step, cell states
            P = ops.tapgemm(pk_g, below, h_split[layer], NB, T, zero_pad_rows=False, out_split=False)
            cst = torch.empty(4 * R * H, dtype=torch.float32, device=dev)
            lib.call("idv_lstm_scan_c", P, NB, T, H, cst, T)
            # BPTT: one cell kernel + one dh = dP W_hh tap-GEMM per step
            dP = _zeros(4 * R * 4 * H, dev)
            dP_step = torch.empty(2 * 4 * NB * 4 * H, dtype=torch.bfloat16, device=dev)
            dc = torch.empty(4 * NB * H, dtype=torch.float32, device=dev)
            step_planes = _HRows(dP_step, NB, 4 * H)
            dh_rec = None
            for t in range(T - 1, -1, -1):
                lib.call("idv_lstm_cell_bwd_step", P, cst, dH, dh_rec, dc, NB, T, H, t, 1 if dh_rec is None else 0, dP,
                         dP_step)
                if t:
                    dh_rec = ops.tapgemm(pk_hh, step_planes, None, NB, 0, zero_pad_rows=False, out_split=False)
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


class _EncoderTrainFn(torch.autograd.Function):
    """Autograd plumbing: forward = EncoderTrainStep.forward, backward hands dL/dlatent to EncoderTrainStep.backward,
    which writes the parameter gradients itself (the parameters are inputs only so that autograd calls backward)."""

    @staticmethod
    def forward(ctx, step, x, *params):
        latent, stft_x, acts = step.forward(x)
        ctx.step = step
        ctx.n = len(params)
        step._aux = (stft_x, acts)
        return latent

    @staticmethod
    def backward(ctx, dlatent):
        ctx.step.backward(dlatent.contiguous())
        return (None, None) + (None,) * ctx.n


def encoder_train_forward(enc, x):
    """latent (with a grad_fn), stft_x, encoder activation planes of one train-mode forward."""
    step = getattr(enc, "_train_step", None)
    if step is None:
        step = enc._train_step = EncoderTrainStep(enc)
    params = [p for p in enc.parameters() if p.requires_grad]
    latent = _EncoderTrainFn.apply(step, x, *params)
    stft_x, acts = step._aux
    return latent, stft_x, acts


# ------------------------------------------------------------------------------------------------------------------
# optimiser + gradient all-reduce (data-parallel training: SURVEY §8(e))
# ------------------------------------------------------------------------------------------------------------------
class FlatAdam:
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics (train_nsvae.py:L200: lr from the config,
    weight_decay = 0.001) on ONE flat fp32 buffer: the parameters that receive gradients are re-pointed into a flat
    tensor, their gradients are gathered into a flat bucket (one NCCL all-reduce per step when a process group is
    given: mean over ranks, DDP semantics), the update is one ``idv_adam_step`` launch."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None,
                 world_size=1):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.group, self.world = process_group, world_size
        self.step_count = 0
        self.flat = None

    def _build(self):
        # parameters without a gradient after the first backward (the unused dense.*) are left out, like
        # torch.optim.Adam skips parameters whose .grad is None
        self.live = [p for p in self.params if p.grad is not None]
        n = sum(p.numel() for p in self.live)
        dev = self.live[0].device
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.gflat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        self.views = []
        for p in self.live:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            self.views.append((off, k))
            off += k

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def step(self):
        if self.flat is None:
            self._build()
        for p, (off, k) in zip(self.live, self.views):
            self.gflat[off:off + k].copy_(p.grad.reshape(-1))
        if self.group is not None and self.world > 1:
            torch.distributed.all_reduce(self.gflat, group=self.group)
            self.gflat.div_(self.world)
        self.step_count += 1
        lib.call("idv_adam_step", self.flat, self.gflat, self.m, self.v, self.flat.numel(), float(self.lr),
                 float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.wd), self.step_count)
        for p in self.live:                          # the kernel wrote the parameters behind autograd's back:
            torch.autograd.graph.increment_version(p)    # bump the versions so the weight-pack caches rebuild
