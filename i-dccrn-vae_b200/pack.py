"""Host-side weight pre-packing (done once per weight version, never per forward).

Turns reference-layout parameters into the operands of the kernels in include/idv.h:
  * complex (transposed) conv  -> block-real tap-GEMM weights with ComplexBatchNormal(eval) folded in
    (SURVEY §9 V3/V4/V5): out = Z (W x + b - mu) + beta = (Z W) x + (Z (b - mu) + beta);
  * nn.LSTM input projections and ComplexDense -> tap-GEMM weights;
  * STFT / iSTFT windowed DFT bases.
All folding is done in float64 and rounded once to float32.
"""
import math
import os

import torch

EPS_CBN = 1e-5


def round8(c):
    return (c + 7) // 8 * 8


_TABLES = {}


def dev_table(rows, device):
    """int32 [n][6] unit / tap table on ``device``.  Tables depend on the layer geometry only, not on the weights, but a
    training step re-packs every layer - and a host -> device copy from pageable memory blocks the host until every
    kernel queued before it has run (188 such copies made one training step host-bound).  Identical tables are therefore
    uploaded once per device and shared."""
    import array
    flat = array.array("i", [int(v) for r in rows for v in r])
    key = (str(device), flat.tobytes())
    t = _TABLES.get(key)
    if t is None:
        if len(_TABLES) > 4096:
            _TABLES.clear()
        t = _TABLES[key] = torch.tensor(flat, dtype=torch.int32).reshape(-1, 6).to(device)
    return t


_ZEROS = {}


def dev_zeros(n, device):
    """Shared read-only fp32 zero vector (bias of the gradient packs) - no host -> device copy per pack."""
    key = (str(device), int(n))
    if key not in _ZEROS:
        _ZEROS[key] = torch.zeros(int(n), dtype=torch.float32, device=device)
    return _ZEROS[key]


PAIR_PLANES = [os.environ.get("IDV_PAIR_PLANES", "1") != "0"]   # narrow (transposed) convs: 256 / N output planes per tensor-core tile (TapGemmPack._tc_paired)


class TapGemmPack:
    """Operands of one idv_tapgemm_* launch (weights, bias, unit/tap tables)."""

    def __init__(self, w, bias, units, taps, N, out_planes, out_ld, prelu, slope, device):
        self.w = w.to(device=device, dtype=torch.float32).contiguous()          # (no copy / no sync when already there)
        self.bias = bias.to(device=device, dtype=torch.float32).contiguous()
        self.units = dev_table(units, device)
        self.taps = dev_table(taps, device)
        self.n_units = len(units)
        self.N = N
        self.out_planes = out_planes
        self.out_ld = out_ld
        self.prelu = bool(prelu)
        self.slope = float(slope)
        self._tc = None
        self._units_l, self._taps_l = [list(u) for u in units], [list(t) for t in taps]

    def tc_eligible(self):
        n_ok = self.N == 32 or self.N % 64 == 0
        return n_ok and all(t[4] % 64 == 0 and t[3] % 8 == 0 for t in self._taps_l) and self.out_ld % 8 == 0

    def tc(self):
        """Operands of idv_tapgemm_tc: K-major bf16 hi/lo weight slots [2][slots][N][kc_max], taps with
        w_off -> slot, units with reserved -> number of 64-wide K steps."""
        if self._tc is None:
            if not self.tc_eligible():
                raise RuntimeError("tap-GEMM (N=%d) does not fit the tensor-core kernel's shape rules" % self.N)
            if getattr(self, "pair_planes", False):
                dev = self.w.device
                w = self.w.detach() if PACK_ON_DEVICE[0] else self.w.detach().cpu()
                self._tc = self._tc_paired(w, dev)
            else:
                self._tc = self.tc_plain()
        return self._tc

    def tc_plain(self):
        """The one-output-plane-per-unit tables (what tc() returns unless the pack computes several planes per tile)."""
        if self.__dict__.get("_tc_plain_cache") is None:
            if not self.tc_eligible():
                raise RuntimeError("tap-GEMM (N=%d) does not fit the tensor-core kernel's shape rules" % self.N)
            dev = self.w.device
            w = self.w.detach() if PACK_ON_DEVICE[0] else self.w.detach().cpu()
            kc_max = max(t[4] for t in self._taps_l)
            # slot ids in (K width, weight offset) order: the slots of one K width whose blocks follow one another in
            # ``w`` (the usual case: tap blocks of one source) are then filled by ONE strided view instead of one copy
            # per slot - a training step re-packs every layer, and every tiny torch op is ~15 us of host time
            keys = sorted({(t[4], t[5]) for t in self._taps_l})
            slots = {(w_off, kc): i for i, (kc, w_off) in enumerate(keys)}
            taps = [[t[0], t[1], t[2], t[3], t[4], slots[(t[5], t[4])]] for t in self._taps_l]
            wt = torch.zeros(len(slots), self.N, kc_max, dtype=torch.float32, device=w.device)
            i = 0
            while i < len(keys):
                kc, o0 = keys[i]
                blk = kc * self.N
                j = i + 1
                while j < len(keys) and keys[j] == (kc, o0 + (j - i) * blk):
                    j += 1
                wt[i:j, :, :kc] = w.as_strided((j - i, kc, self.N), (blk, self.N, 1), w.storage_offset() + o0).transpose(1, 2)
                i = j
            hi = wt.to(torch.bfloat16)
            lo = (wt - hi.to(torch.float32)).to(torch.bfloat16)
            units = []
            for u in self._units_l:
                ks = sum(t[4] // 64 for t in self._taps_l[u[0]:u[0] + u[1]])
                units.append([u[0], u[1], u[2], u[3], u[4], ks])
            self._tc_plain_cache = dict(wt=torch.stack((hi, lo)).contiguous().to(dev), kc_max=kc_max, n_slots=len(slots),
                                        taps=dev_table(taps, dev), units=dev_table(units, dev),
                                        min_ksteps=min(u[5] for u in units), taps_l=taps, units_l=units)
        return self._tc_plain_cache

    def tc_stream(self, k, cp0, cp1):
        """Tables of the same tap-GEMM for a frame-streaming step of k frames, on LIVE rows only.  A streaming plane set
        holds k + 1 rows per stream (row 0 = the carried x[t-1], rows 1..k = the new frames); read as ONE row of
        (k + 1) * cp channels per stream it makes the time tap a channel offset: output frame j of unit u becomes its own
        unit whose taps read channel offset ch_off + (j + 1 - dt) * cp at dt = 0 and write column offset
        out_ch_off + (j + 1) * out_ld of an output row of (k + 1) * out_ld columns.  The GEMM then has n_streams rows
        instead of n_streams * (k + 1): no tile row is spent on pad rows (half of all rows at k = 1) and the weights
        stream through half as many tiles.  Same memory layouts, same arithmetic.  Built from the one-plane-per-unit
        tables (tc_plain) also for the packs that compute several output planes per tile in the batched path.  Returns
        the full operand set (wt, kc_max, n_slots, units, taps, n_units, min_ksteps) or None."""
        if self.N > self.out_ld:
            return None
        cache = self.__dict__.setdefault("_tc_stream", {})
        key = (k, cp0, cp1)
        if key not in cache:
            tc = self.tc_plain()
            taps_tc = tc["taps_l"]
            units_tc = tc["units_l"]
            cps = (cp0, cp1)
            taps, units = [], []
            for u in units_tc:
                if any(not 0 <= taps_tc[i][2] <= 1 for i in range(u[0], u[0] + u[1])):
                    cache[key] = None                      # a time tap outside the carried frame
                    return None
                for j in range(k):
                    begin = len(taps)
                    for i in range(u[0], u[0] + u[1]):
                        t = taps_tc[i]
                        taps.append([t[0], t[1], 0, t[3] + (j + 1 - t[2]) * cps[t[0]], t[4], t[5]])
                    units.append([begin, u[1], u[2], u[3] + (j + 1) * self.out_ld, u[4], u[5]])
            dev = tc["wt"].device
            cache[key] = dict(units=dev_table(units, dev), taps=dev_table(taps, dev), n_units=len(units), wt=tc["wt"],
                              kc_max=tc["kc_max"], n_slots=tc["n_slots"], min_ksteps=tc["min_ksteps"])
        return cache[key]


def cbn_fold(bn):
    """Per-channel Z (C,2,2) and b' (C,2) of ComplexBatchNormal(train=False)
    (model/complex_progress.py:L161-209).  ``bn`` maps leaf names to tensors."""
    d = {k: v.detach().double().cpu().reshape(-1) for k, v in bn.items()}
    vrr, vri, vii = d["Vrr"], d["Vri"], d["Vii"]
    delta = torch.clamp(vrr * vii - vri ** 2 + EPS_CBN, min=1e-8)
    s = torch.sqrt(delta)
    t = torch.sqrt(vrr + vii + 2 * s + EPS_CBN)
    ist = 1.0 / (s * t + EPS_CBN)
    wrr, wii, wri = (vii + s) * ist, (vrr + s) * ist, -vri * ist
    grr, gri, gii = d["gamma_rr"], d["gamma_ri"], d["gamma_ii"]
    Z = torch.stack((torch.stack((grr * wrr + gri * wri, grr * wri + gri * wii), -1),
                     torch.stack((gri * wrr + gii * wri, gri * wri + gii * wii), -1)), -2)   # (C,2,2)
    mu = torch.stack((d["running_mean_real"], d["running_mean_imag"]), -1)                    # (C,2)
    beta = torch.stack((d["beta_r"], d["beta_i"]), -1)
    bprime = beta - torch.einsum("cij,cj->ci", Z, mu)
    return Z, bprime


def _identity_fold(C, device="cpu"):
    """(Z, b') of "no ComplexBatchNormal": None selects _block_weights' copy-only path."""
    return None, None


def _block_weights(m_re, m_im, b_re, b_im, Z, bprime, ch_in, ch_out):
    """m_re/m_im: (taps, Cin, Cout) real matrices of the re/im sub-layers (input-major).
    Returns W (taps, 2*ch_in, 2*ch_out) and bias (2*ch_out) with the 2x2 fold applied.
    Raw block: y_re = m_re x_re - m_im x_im, y_im = m_re x_im + m_im x_re (complex_progress.py:L17-19)."""
    taps, cin, cout = m_re.shape
    if Z is None:
        # no fold (the training path keeps ComplexBatchNormal separate): the four blocks are copies of +-m_re / +-m_im -
        # a handful of ops in the weights' own precision instead of the fp64 fold arithmetic (re-done every step)
        W = torch.zeros(taps, 2 * ch_in, 2 * ch_out, dtype=torch.float64, device=m_re.device)
        W[:, :cin, :cout] = m_re
        W[:, ch_in:ch_in + cin, :cout] = -m_im
        W[:, :cin, ch_out:ch_out + cout] = m_im
        W[:, ch_in:ch_in + cin, ch_out:ch_out + cout] = m_re
        bias = torch.zeros(2 * ch_out, dtype=torch.float64, device=m_re.device)
        bias[:cout] = b_re - b_im
        bias[ch_out:ch_out + cout] = b_re + b_im
        return W, bias
    m_re, m_im = m_re.double(), m_im.double()
    zrr, zri, zir, zii = (Z[:, 0, 0], Z[:, 0, 1], Z[:, 1, 0], Z[:, 1, 1])
    W = torch.zeros(taps, 2 * ch_in, 2 * ch_out, dtype=torch.float64, device=m_re.device)
    # rows: [re-in | im-in], cols: [re-out | im-out]
    W[:, :cin, :cout] = zrr * m_re + zri * m_im
    W[:, ch_in:ch_in + cin, :cout] = -zrr * m_im + zri * m_re
    W[:, :cin, ch_out:ch_out + cout] = zir * m_re + zii * m_im
    W[:, ch_in:ch_in + cin, ch_out:ch_out + cout] = -zir * m_im + zii * m_re
    yb_re = (b_re - b_im).double()
    yb_im = (b_re + b_im).double()
    bias = torch.zeros(2 * ch_out, dtype=torch.float64, device=m_re.device)
    bias[:cout] = zrr * yb_re + zri * yb_im + bprime[:, 0]
    bias[ch_out:ch_out + cout] = zir * yb_re + zii * yb_im + bprime[:, 1]
    return W, bias


# Packs are computed on the CPU by default (inference: once per weight version).  A training step changes every
# weight, so its packs are recomputed every step: `with on_device():` keeps the (O(parameters)) packing arithmetic
# on the device the weights live on instead of round-tripping them through the host.
PACK_ON_DEVICE = [False]


class on_device:
    def __enter__(self):
        self.old = PACK_ON_DEVICE[0]
        PACK_ON_DEVICE[0] = True

    def __exit__(self, *a):
        PACK_ON_DEVICE[0] = self.old


def _cpu(t):
    t = t.detach()
    return t if PACK_ON_DEVICE[0] else t.cpu()


def _tc_paired(self, w, dev):
    """Tensor-core tables of a NARROW (transposed) conv with G = ``plane_group`` (default 2; measured: 4 planes of the
    N = 64 transposed conv are no faster than 2) consecutive output planes per unit (N_tc = G N,
    columns [g N, (g+1) N) = plane G q + g; the kernel wraps them into the planes because N_tc > out_ld): neighbouring
    output planes read overlapping input planes (a stride-2 conv: planes 2fo-2 .. 2fo+2; the transposed conv: fo/2 - 1 ..
    fo/2 + 1), so one tile loads each activation box once for all of them and issues N = 256 MMAs - the narrow tiles were
    bound by operand ingest and by their epilogue, not by the tensor pipe.  Taps of the group are merged by (source,
    plane, dt, K range); a plane that lacks a tap gets a zero weight block."""
    import collections
    N = self.N
    G = getattr(self, "plane_group", 2)
    NG = G * N
    kc_max = max(t[4] for t in self._taps_l)
    units, taps, slots = [], [], {}
    for q in range((len(self._units_l) + G - 1) // G):
        grp = self._units_l[G * q:G * q + G]
        assert grp[0][2] == G * q and all(u[3] == 0 and u[4] == 0 for u in grp)
        groups = collections.OrderedDict()
        for g, u in enumerate(grp):
            assert u[2] == G * q + g
            for t in self._taps_l[u[0]:u[0] + u[1]]:
                groups.setdefault((t[0], t[1], t[2], t[3], t[4]), [None] * G)[g] = t[5]
        begin = len(taps)
        for key, offs in groups.items():
            skey = tuple(offs) + (key[4],)
            if skey not in slots:
                slots[skey] = len(slots)
            taps.append([key[0], key[1], key[2], key[3], key[4], slots[skey]])
        units.append([begin, len(groups), G * q, 0, 0, sum(k[4] // 64 for k in groups)])
    wt = torch.zeros(len(slots), NG, kc_max, dtype=torch.float32, device=w.device)
    for skey, si in slots.items():
        kc = skey[-1]
        for g, off in enumerate(skey[:-1]):
            if off is not None:
                wt[si, g * N:(g + 1) * N, :kc] = w[off:off + kc * N].view(kc, N).t()
    hi = wt.to(torch.bfloat16)
    lo = (wt - hi.to(torch.float32)).to(torch.bfloat16)
    return dict(wt=torch.stack((hi, lo)).contiguous().to(dev), kc_max=kc_max, n_slots=len(slots),
                taps=dev_table(taps, dev),
                units=dev_table(units, dev),
                N=NG, n_units=len(units), bias=torch.cat([self.bias] * G).contiguous())


TapGemmPack._tc_paired = _tc_paired


def pack_conv(conv_re_w, conv_re_b, conv_im_w, conv_im_b, bn, slope, f_in, stride_f, pad_f, device, pad_t=1):
    """Complex conv (kernel (kh,kw), stride (stride_f,1), pad (pad_f,pad_t)) [+ CBN + PReLU].
    weights: (Cout, Cin, kh, kw).  Time tap kt reads x[t-pad_t+kt]: the causal layer (pad_t = 1, last column
    dropped) reads x[t-1], x[t]; the non-causal one (pad_t = 0, model/net_config.py) x[t], x[t+1]
    and has kw-1 fewer valid frames."""
    wr, wi = _cpu(conv_re_w), _cpu(conv_im_w)
    cout, cin, kh, kw = wr.shape
    if kw not in (1, 2) or pad_t not in (0, 1):
        raise NotImplementedError("complex conv is built for 1- or 2-tap time kernels with time padding 0 or 1 "
                                  "(got kernel %d, padding %d)" % (kw, pad_t))
    ch_in, ch_out = round8(cin), round8(cout)
    Z, bp = cbn_fold(bn) if bn is not None else _identity_fold(cout, wr.device)
    m_re = wr.permute(2, 3, 1, 0).reshape(kh * kw, cin, cout)
    m_im = wi.permute(2, 3, 1, 0).reshape(kh * kw, cin, cout)
    W, bias = _block_weights(m_re, m_im, _cpu(conv_re_b), _cpu(conv_im_b), Z, bp, ch_in, ch_out)
    N, kc = 2 * ch_out, 2 * ch_in
    f_out = (f_in + 2 * pad_f - kh) // stride_f + 1
    units, taps = [], []
    for fo in range(f_out):
        begin = len(taps)
        for kf in range(kh):
            fi = stride_f * fo + kf - pad_f
            if fi < 0 or fi >= f_in:
                continue
            for kt in range(kw):
                taps.append([0, fi, pad_t - kt, 0, kc, (kf * kw + kt) * kc * N])
        units.append([begin, len(taps) - begin, fo, 0, 0, 0])
    p = TapGemmPack(W.reshape(-1), bias, units, taps, N, f_out, N, slope is not None, slope or 0.0, device)
    p.f_out, p.c_out = f_out, cout
    # (measured on B200: grouping the output planes of the strided conv 32 -> 64, N = 128, into N = 256 units makes the
    #  tile MMA-bound at 97 % tensor-pipe active but 7 % slower - 14 instead of 10 K chunks per plane pair with zero
    #  weight blocks; the layer stays one plane per unit)
    return p


def pack_conv_transpose(t_re_w, t_re_b, t_im_w, t_im_b, bn, slope, f_in, c_p, c_skip, device,
                        stride_f=2, pad_f=2):
    """Complex transposed conv (time padding 0) in gather form over two sources: the running
    activation p (c_p complex channels) and the skip tensor (c_skip, 0 = none / zero skip).
    weights: (Cin_total, Cout, kh, kw); input channel order [p | skip] like torch.cat
    (model/pvae_module.py:L2098).  Time tap kt reads x[t-kt]; the causal layer drops the last of the T+kw-1
    output frames, the non-causal one keeps it (the caller sets the valid frame count)."""
    wr, wi = _cpu(t_re_w), _cpu(t_im_w)
    cin_tot, cout, kh, kw = wr.shape
    if kw not in (1, 2):
        raise NotImplementedError("complex transposed conv is built for 1- or 2-tap time kernels")
    ch_out = round8(cout)
    Z, bp = cbn_fold(bn) if bn is not None else _identity_fold(cout, wr.device)
    f_out = (f_in - 1) * stride_f - 2 * pad_f + kh
    N = 2 * ch_out
    srcs = [(0, 0, c_p)]
    if c_skip:
        srcs.append((1, c_p, c_skip))
    Ws, w_offs, kcs, off = [], [], [], 0
    bias = None
    for (src, c0, cn) in srcs:
        ch_in = round8(cn)
        m_re = wr[c0:c0 + cn].permute(2, 3, 0, 1).reshape(kh * kw, cn, cout)
        m_im = wi[c0:c0 + cn].permute(2, 3, 0, 1).reshape(kh * kw, cn, cout)
        W, b = _block_weights(m_re, m_im, _cpu(t_re_b), _cpu(t_im_b), Z, bp, ch_in, ch_out)
        if bias is None:
            bias = b
        Ws.append(W.reshape(-1))
        w_offs.append(off)
        kcs.append(2 * ch_in)
        off += W.numel()
    units, taps = [], []
    for fo in range(f_out):
        begin = len(taps)
        for kf in range(kh):
            num = fo + pad_f - kf
            if num % stride_f:
                continue
            fi = num // stride_f
            if fi < 0 or fi >= f_in:
                continue
            for kt in range(kw):
                for si, (src, _, _) in enumerate(srcs):
                    taps.append([src, fi, kt, 0, kcs[si], w_offs[si] + (kf * kw + kt) * kcs[si] * N])
        units.append([begin, len(taps) - begin, fo, 0, 0, 0])
    p = TapGemmPack(torch.cat(Ws), bias, units, taps, N, f_out, N, slope is not None, slope or 0.0, device)
    p.f_out, p.c_out = f_out, cout
    p.pair_planes = PAIR_PLANES[0] and N in (64, 128) and stride_f == 2     # two output planes per tensor-core tile (_tc_paired)
    return p


def pack_dense_conv_transpose(w_read, b_read, w_imag, b_imag, t_re_w, t_re_b, t_im_w, t_im_b, bn, slope, f_in, c_p, c_skip,
                              device, stride_f=2, pad_f=2):
    """ComplexDense followed by the first CAUSAL complex transposed conv, composed in float64 into ONE tap-GEMM on the z
    planes (model/pvae_module.py:L2085-2099: dense -> reshape (B, T, C, F, 2) -> permute -> decoders[0]):
        out[fo][t] = sum_kt C[fo,kt] z[t-kt] + (sum over the kt whose frame t-kt exists of e[fo,kt]) + b
        C[fo,kt] = sum over (kf, fi) feeding fo of D_fi W[kf,kt],   e[fo,kt] = sum d_fi W[kf,kt]
    with D_fi / d_fi the dense layer's matrix / bias for plane fi.  K per time tap is 2*zdim instead of 2*C per (input
    plane, time tap): at C = 256, F = 5 the layer runs 1.5 instead of 8.4 GMAC per 4-s utterance and the dense launch,
    its 5 output planes and their re-read disappear.  The frame t = 0 only sees the kt = 0 constant (x[-1] is the zero
    pad row, not d): ``bias_first``.  The skip source (c_skip > 0) keeps its taps.  Sources: 0 = z planes, 1 = skip."""
    wr, wi = _cpu(t_re_w), _cpu(t_im_w)
    cin_tot, cout, kh, kw = wr.shape
    if kw != 2:
        raise NotImplementedError("the dense + transposed-conv composition is built for the 2-tap causal time kernel")
    dev = wr.device
    zdim = w_read.shape[1]
    ch_z, ch_p, ch_out = round8(zdim), round8(c_p), round8(cout)
    Z, bp = cbn_fold(bn) if bn is not None else _identity_fold(cout, dev)
    f_out = (f_in - 1) * stride_f - 2 * pad_f + kh
    N = 2 * ch_out
    m_re = wr[:c_p].permute(2, 3, 0, 1).reshape(kh * kw, c_p, cout)
    m_im = wi[:c_p].permute(2, 3, 0, 1).reshape(kh * kw, c_p, cout)
    Wp, bias = _block_weights(m_re, m_im, _cpu(t_re_b), _cpu(t_im_b), Z, bp, ch_p, ch_out)       # (taps, 2 ch_p, N), (N)
    # dense: p[f][r] = z[r] Dm[f] + d[f]   (re and im parts are independent real linears: complex_progress.py:L83-89)
    Dm = torch.zeros(f_in, 2 * ch_z, 2 * ch_p, dtype=torch.float64, device=dev)
    d = torch.zeros(f_in, 2 * ch_p, dtype=torch.float64, device=dev)
    for part, (w, b) in enumerate(((w_read, b_read), (w_imag, b_imag))):
        w = _cpu(w).double().reshape(c_p, f_in, zdim)                  # [c][f][k]
        Dm[:, part * ch_z:part * ch_z + zdim, part * ch_p:part * ch_p + c_p] = w.permute(1, 2, 0)
        d[:, part * ch_p:part * ch_p + c_p] = _cpu(b).double().reshape(c_p, f_in).t()
    kz = 2 * ch_z
    Ws, units, taps = [], [], []
    bias_full = torch.zeros(f_out, N, dtype=torch.float64, device=dev)
    bias_first = torch.zeros(f_out, N, dtype=torch.float64, device=dev)
    off = 0
    # skip source weights (unchanged taps of pack_conv_transpose)
    if c_skip:
        ch_s = round8(c_skip)
        s_re = wr[c_p:c_p + c_skip].permute(2, 3, 0, 1).reshape(kh * kw, c_skip, cout)
        s_im = wi[c_p:c_p + c_skip].permute(2, 3, 0, 1).reshape(kh * kw, c_skip, cout)
        Wsk, _ = _block_weights(s_re, s_im, _cpu(t_re_b), _cpu(t_im_b), Z, bp, ch_s, ch_out)
        Ws.append(Wsk.reshape(-1))
        skip_off, ks = 0, 2 * ch_s
        off = Wsk.numel()
    for fo in range(f_out):
        begin = len(taps)
        feeds = []
        for kf in range(kh):
            num = fo + pad_f - kf
            if num % stride_f:
                continue
            fi = num // stride_f
            if 0 <= fi < f_in:
                feeds.append((kf, fi))
        e = []
        for kt in range(kw):
            Cm = torch.zeros(kz, N, dtype=torch.float64, device=dev)
            ev = torch.zeros(N, dtype=torch.float64, device=dev)
            for kf, fi in feeds:
                Cm += Dm[fi] @ Wp[kf * kw + kt]
                ev += d[fi] @ Wp[kf * kw + kt]
            e.append(ev)
            Ws.append(Cm.reshape(-1))
            taps.append([0, 0, kt, 0, kz, off])
            off += kz * N
        if c_skip:
            for kf, fi in feeds:
                for kt in range(kw):
                    taps.append([1, fi, kt, 0, ks, skip_off + (kf * kw + kt) * ks * N])
        bias_full[fo] = bias + e[0] + e[1]
        bias_first[fo] = bias + e[0]
        units.append([begin, len(taps) - begin, fo, 0, fo * N, 0])
    p = TapGemmPack(torch.cat(Ws), bias_full.reshape(-1), units, taps, N, f_out, N, slope is not None, slope or 0.0, device)
    p.bias_first = bias_first.reshape(-1).to(torch.float32).contiguous().to(device)
    p.f_out, p.c_out = f_out, cout
    p.pair_planes = False
    return p


def pack_enc0(conv_re_w, conv_re_b, conv_im_w, conv_im_b, bn, slope, device):
    """First encoder layer (Cin = 1): w [10][2][2*Cout], bias [2*Cout] for idv_enc0_fwd."""
    wr, wi = _cpu(conv_re_w), _cpu(conv_im_w)
    cout, cin, kh, kw = wr.shape
    assert cin == 1 and kh == 5 and kw == 2 and cout % 32 == 0
    Z, bp = cbn_fold(bn) if bn is not None else _identity_fold(cout, wr.device)
    m_re = wr.permute(2, 3, 1, 0).reshape(10, 1, cout)
    m_im = wi.permute(2, 3, 1, 0).reshape(10, 1, cout)
    W = torch.zeros(10, 2, 2 * cout, dtype=torch.float64, device=wr.device)
    Wb, bias = _block_weights(m_re, m_im, _cpu(conv_re_b), _cpu(conv_im_b), Z, bp, 8, cout)
    W[:, 0] = Wb[:, 0]           # re-in row
    W[:, 1] = Wb[:, 8]           # im-in row (ch_in = 8 padding of one channel)
    return (W.to(torch.float32).contiguous().to(device), bias.to(torch.float32).to(device), cout,
            float(slope if slope is not None else 1.0))


ENC0_ROWS_LD = 576        # columns of the STFT activation rows: 8 (zero bins below bin 0) + 2*257 + zero padding
ENC0_COL0 = 8             # column of (bin 0, re): a multiple of 8 so the STFT epilogue writes 16-byte vectors


def pack_enc0_tc(conv_re_w, conv_re_b, conv_im_w, conv_im_b, bn, slope, f_in, device):
    """First encoder layer (Cin = 1, kernel (5, 2), stride (2, 1), causal) as a tap-GEMM on the STFT's activation rows
    [1 plane][R][ENC0_ROWS_LD] (column ENC0_COL0 + 2*bin + part; written by the STFT GEMM's epilogue): output plane fo
    reads the bins 2fo-2 .. 2fo+2 = the 10 columns from 4*fo + 4 on.  Four consecutive output planes 4q .. 4q+3 read 22
    neighbouring columns, so they share ONE 64-column box starting at 16q (a TMA box starts at a multiple of 8
    columns) and run as one N = 256 unit (TapGemmPack._tc_paired): plane g of the group has its 10 live K rows at
    4 + 4g.  2 time taps, K = 64 per tap of which 22 are live - the layer was a 0.85 ms issue-bound SIMT kernel; the dead MMA columns cost nothing next to its 1.3 GB
    of output."""
    w, bias, cout, sl = pack_enc0(conv_re_w, conv_re_b, conv_im_w, conv_im_b, bn, slope, "cpu")    # [10][2][2 cout]
    N = 2 * cout
    f_out = (f_in + 4 - 5) // 2 + 1
    if ENC0_COL0 != 8 or ENC0_COL0 + 2 * f_in > ENC0_ROWS_LD or 16 * ((f_out - 1) // 4) + 64 > ENC0_ROWS_LD:
        raise RuntimeError("STFT rows of %d columns do not hold %d bins" % (ENC0_ROWS_LD, f_in))
    mats = []
    for g in range(4):                                    # position of the plane in its group of 4
        for kt in range(2):
            m = torch.zeros(64, N, dtype=torch.float32)
            for kf in range(5):
                for part in range(2):
                    m[4 + 4 * g + 2 * kf + part] = w[kf * 2 + kt, part]
            mats.append(m.reshape(-1))
    units, taps = [], []
    for fo in range(f_out):
        q, g = fo // 4, fo % 4
        begin = len(taps)
        for kt in range(2):                               # time tap kt reads x[t - 1 + kt]
            taps.append([0, 0, 1 - kt, 16 * q, 64, (g * 2 + kt) * 64 * N])
        units.append([begin, 2, fo, 0, 0, 0])
    p = TapGemmPack(torch.cat(mats), bias.cpu(), units, taps, N, f_out, N, slope is not None, sl if slope is not None else 0.0,
                    device)
    p.f_out, p.c_out = f_out, cout
    p.pair_planes = PAIR_PLANES[0] and N == 64
    p.plane_group = 4
    return p


def pack_dec5(t_re_w, t_re_b, t_im_w, t_im_b, bn, slope, c_p, c_skip, device):
    """Last decoder layer (Cout = 1): w [10][2*ch_p + 2*ch_skip][2], bias [2] for idv_dec5_head_fwd."""
    wr, wi = _cpu(t_re_w), _cpu(t_im_w)
    cin_tot, cout, kh, kw = wr.shape
    assert cout == 1 and kh == 5 and kw == 2 and cin_tot >= c_p + c_skip
    Z, bp = cbn_fold(bn) if bn is not None else _identity_fold(1, wr.device)
    parts, bias = [], None
    for (c0, cn) in ((0, c_p), (c_p, c_skip)):
        if cn == 0:
            continue
        ch_in = round8(cn)
        m_re = wr[c0:c0 + cn].permute(2, 3, 0, 1).reshape(10, cn, 1)
        m_im = wi[c0:c0 + cn].permute(2, 3, 0, 1).reshape(10, cn, 1)
        W, b = _block_weights(m_re, m_im, _cpu(t_re_b), _cpu(t_im_b), Z, bp, ch_in, 8)
        bias = b if bias is None else bias
        parts.append(torch.stack((W[:, :, 0], W[:, :, 8]), -1))      # (10, 2*ch_in, 2): re-out, im-out
    w = torch.cat(parts, 1).to(torch.float32).contiguous().to(device)
    b2 = torch.stack((bias[0], bias[8])).to(torch.float32).to(device)
    return w, b2, float(slope if slope is not None else 1.0)


def unfold_dec5_wgrad(dW, c_p, c_skip, cin_tot):
    """dW (10, 2 ch_p + 2 ch_skip, 2): gradient of pack_dec5's w10 (no fold) -> gradients of tconv_re / tconv_im.weight
    (cin_tot, 1, 5, 2): w10[tap][ci][0] = m_re, [ch+ci][0] = -m_im, [ci][1] = m_im, [ch+ci][1] = m_re."""
    d_re = torch.zeros(cin_tot, 1, 5, 2, dtype=dW.dtype, device=dW.device)
    d_im = torch.zeros_like(d_re)
    k0 = 0
    for (c0, cn) in ((0, c_p), (c_p, c_skip)):
        if cn == 0:
            continue
        ch = round8(cn)
        blk = dW[:, k0:k0 + 2 * ch]
        g_re = blk[:, :cn, 0] + blk[:, ch:ch + cn, 1]                 # (10, cn)
        g_im = blk[:, :cn, 1] - blk[:, ch:ch + cn, 0]
        d_re[c0:c0 + cn, 0] = g_re.t().reshape(cn, 5, 2)
        d_im[c0:c0 + cn, 0] = g_im.t().reshape(cn, 5, 2)
        k0 += 2 * ch
    return d_re, d_im


HEAD_BINS = [int(os.environ.get("IDV_HEAD_BINS", "16"))]           # output bins per unit of the fused last-layer + head tap-GEMM (1..16)


def pack_dec5_tc(w10, bias2, slope, f_in, kcs, device, bins=None):
    """Last decoder layer on the tensor-core kernel with the head fused (idv_tapgemm_tc_head).
    w10: [10 (kf*2+kt)][sum(kcs)][2] folded weights of pack_dec5 (sources concatenated along k), bias2: [2].
    A unit holds ``bins`` consecutive output bins fo0 .. fo0+bins-1, bin e in columns (2e, 2e+1) of the N = 32 tile.
    Output bin fo reads input plane fi with kernel row kf = fo + 2 - 2 fi (kernel 5, stride 2, pad 2): the bins of a
    unit read the input planes fo0/2 - 1 .. (fo0 + bins - 1)/2 + 1, one weight slot per (source, plane offset, time tap)
    with the rows of the bins that plane does not feed left zero.  More bins per unit = fewer activation reads per bin
    (the layer is bound by them): 6 planes for 8 bins against 3 planes for 2."""
    bins = HEAD_BINS[0] if bins is None else bins
    if not 1 <= bins <= 16 or bins % 2:
        raise ValueError("bins per head unit must be even and <= 16")
    w10 = _cpu(w10).to(torch.float32)
    wdev = w10.device
    N = 32
    ksum = sum(kcs)
    assert w10.shape == (10, ksum, 2) and all(k % 64 == 0 for k in kcs)
    k_off = [0]
    for k in kcs:
        k_off.append(k_off[-1] + k)
    kc_max = max(kcs)
    f_out = 2 * f_in - 1
    offs = list(range(-1, bins // 2 + 1))                  # input plane = fo0/2 + d
    slot_id = {}
    for si in range(len(kcs)):
        for d in offs:
            for kt in range(2):
                slot_id[(si, d, kt)] = len(slot_id)
    # wt[slot][2e + part][k] = w10[kf*2 + kt][k_off[si] + k][part] with kf = e + 2 - 2d: ONE gather through an index
    # table that depends on the geometry only (kept on the device: a training step re-packs this layer every step)
    gkey = ("dec5_tc", bins, tuple(kcs), str(wdev))
    if gkey not in _TABLES:
        idx = torch.full((len(slot_id), N, kc_max), 10 * ksum * 2, dtype=torch.int64)       # -> the appended zero
        for (si, d, kt), sl in slot_id.items():
            k = torch.arange(kcs[si])
            for e in range(bins):
                kf = e + 2 - 2 * d
                if 0 <= kf < 5:
                    for part in range(2):
                        idx[sl, 2 * e + part, :kcs[si]] = ((kf * 2 + kt) * ksum + k_off[si] + k) * 2 + part
        _TABLES[gkey] = idx.to(wdev)
    wt = torch.cat((w10.reshape(-1), w10.new_zeros(1)))[_TABLES[gkey]]
    hi = wt.to(torch.bfloat16)
    lo = (wt - hi.to(torch.float32)).to(torch.bfloat16)
    n_units = (f_out + bins - 1) // bins
    units, taps = [], []
    for q in range(n_units):
        begin = len(taps)
        fo0 = q * bins
        nb = min(bins, f_out - fo0)
        for d in offs:
            fi = fo0 // 2 + d
            if fi < 0 or fi >= f_in:
                continue
            if not any(0 <= e + 2 - 2 * d < 5 for e in range(nb)):      # the plane feeds none of the unit's live bins
                continue
            for kt in range(2):
                for si in range(len(kcs)):
                    taps.append([si, fi, kt, 0, kcs[si], slot_id[(si, d, kt)]])
        ks = sum(t[4] // 64 for t in taps[begin:])
        units.append([begin, len(taps) - begin, fo0, nb, 0, ks])
    bias = torch.zeros(N, device=wdev)
    bias[0:2] = _cpu(bias2).to(torch.float32)
    return dict(wt=torch.stack((hi, lo)).contiguous().to(device), kc_max=kc_max, n_slots=len(slot_id),
                taps=dev_table(taps, device),
                units=dev_table(units, device), n_units=n_units,
                bias=bias.to(device), slope=float(slope), N=N)


def pack_lstm_inproj0(lstm_re, lstm_im, hidden, c_in, f_in, device):
    """Layer-0 input projection of both nn.LSTM modules as one tap-GEMM: feature d = c*f_in + f
    (the reshape at model/pvae_module.py:L2241).  N = 8H: [module re gates | module im gates];
    unit p (input part x_re / x_im) writes plane p of G0[2][R][8H]."""
    H, ch = hidden, round8(c_in)
    N = 8 * H
    dev = _cpu(lstm_re["weight_ih_l0"]).device
    W = torch.zeros(f_in, ch, N, dtype=torch.float64, device=dev)
    bias = torch.zeros(N, dtype=torch.float64, device=dev)
    for m, mod in enumerate((lstm_re, lstm_im)):
        wih = _cpu(mod["weight_ih_l0"]).double()                     # (4H, c_in*f_in)
        W[:, :c_in, m * 4 * H:(m + 1) * 4 * H] = wih.reshape(4 * H, c_in, f_in).permute(2, 1, 0)
        bias[m * 4 * H:(m + 1) * 4 * H] = _cpu(mod["bias_ih_l0"]).double() + _cpu(mod["bias_hh_l0"]).double()
    units, taps = [], []
    for p in range(2):
        begin = len(taps)
        for f in range(f_in):
            taps.append([0, f, 0, p * ch, ch, f * ch * N])
        units.append([begin, f_in, p, 0, 0, 0])
    return TapGemmPack(W.reshape(-1), bias, units, taps, N, 2, N, False, 0.0, device)


def pack_lstm_inproj1(lstm_re, lstm_im, hidden, device, layer=1):
    """Layer>=1 input projection: A = hseq of the previous layer [4 streams][R][H]; unit (m,p) uses
    module m's weight_ih_l{layer} and writes plane m*2+p of G1[4][R][4H]."""
    H = hidden
    N = 4 * H
    dev = _cpu(lstm_re["weight_ih_l%d" % layer]).device
    W = torch.zeros(2, H, N, dtype=torch.float64, device=dev)
    bias = torch.zeros(2 * N, dtype=torch.float64, device=dev)
    for m, mod in enumerate((lstm_re, lstm_im)):
        W[m] = _cpu(mod["weight_ih_l%d" % layer]).double().t()
        bias[m * N:(m + 1) * N] = _cpu(mod["bias_ih_l%d" % layer]).double() + _cpu(mod["bias_hh_l%d" % layer]).double()
    units, taps = [], []
    for m in range(2):
        for p in range(2):
            taps.append([0, m * 2 + p, 0, 0, H, m * H * N])
            units.append([len(taps) - 1, 1, m * 2 + p, 0, m * N, 0])
    return TapGemmPack(W.reshape(-1), bias, units, taps, N, 4, N, False, 0.0, device)


def pack_lstm_step(lstm_re, lstm_im, hidden, layer, device):
    """One streaming time step of layer `layer` as a tap-GEMM on the carried h planes [4 streams][NB][H]
    (frame streaming: nn.LSTM's carried state): unit s = (module m, part p) computes h_s W_hh^m^T; for layer >= 1
    the same launch adds the input projection x W_ih^m^T of the layer below's new h (source 0) and both biases.
    Output G [4][NB][4H] fp32."""
    H = hidden
    N = 4 * H
    mats, bias = [], torch.zeros(2 * N, dtype=torch.float64)
    for m, mod in enumerate((lstm_re, lstm_im)):
        if layer > 0:
            mats.append(_cpu(mod["weight_ih_l%d" % layer]).double().t())
            bias[m * N:(m + 1) * N] = _cpu(mod["bias_ih_l%d" % layer]).double() + _cpu(mod["bias_hh_l%d" % layer]).double()
        mats.append(_cpu(mod["weight_hh_l%d" % layer]).double().t())
    W = torch.stack(mats)                                   # (2 or 4, H, 4H), K-major rows
    per = 2 if layer > 0 else 1
    units, taps = [], []
    for m in range(2):
        for p in range(2):
            s = m * 2 + p
            begin = len(taps)
            if layer > 0:
                taps.append([0, s, 0, 0, H, (m * per) * H * N])          # input projection: h of the layer below
                taps.append([1, s, 0, 0, H, (m * per + 1) * H * N])      # recurrent: this layer's carried h
            else:
                taps.append([0, s, 0, 0, H, m * H * N])
            units.append([begin, len(taps) - begin, s, 0, m * N if layer > 0 else 0, 0])
    if layer == 0:
        bias = torch.zeros(N, dtype=torch.float64)          # layer-0 biases are inside the input projection
    return TapGemmPack(W.reshape(-1), bias, units, taps, N, 4, N, False, 0.0, device)


def pack_lstm_h1(lstm, num_layers, c_in, f_in, device):
    """nn.LSTM with ONE hidden unit over the features (c, f, part) flattened as (c*f_in + f)*2 + part (the reshape at
    model/pvae_module.py:L2336-2339): layer-0 input projection as a tap-GEMM (N = 32, columns 0-3 = gates i, f, g, o
    with b_ih_l0 + b_hh_l0 folded in) and the scalar recurrent parameters fp32 [num_layers][12] =
    (w_ih[4] (layers >= 1), w_hh[4], b_ih + b_hh [4] (layers >= 1)) for idv_lstm_h1_fwd."""
    ch, N = round8(c_in), 32
    wih = _cpu(lstm["weight_ih_l0"]).double()                          # (4, c_in*f_in*2)
    dev = wih.device
    w = wih.reshape(4, c_in, f_in, 2)
    W = torch.zeros(f_in, 2 * ch, N, dtype=torch.float64, device=dev)
    W[:, :c_in, :4] = w[..., 0].permute(2, 1, 0)
    W[:, ch:ch + c_in, :4] = w[..., 1].permute(2, 1, 0)
    bias = torch.zeros(N, dtype=torch.float64, device=dev)
    bias[:4] = _cpu(lstm["bias_ih_l0"]).double() + _cpu(lstm["bias_hh_l0"]).double()
    taps = [[0, f, 0, 0, 2 * ch, f * 2 * ch * N] for f in range(f_in)]
    units = [[0, f_in, 0, 0, 0, 0]]
    inproj = TapGemmPack(W.reshape(-1), bias, units, taps, N, 1, N, False, 0.0, device)
    rec = torch.zeros(num_layers, 12, dtype=torch.float64, device=dev)
    for l in range(num_layers):
        rec[l, 4:8] = _cpu(lstm["weight_hh_l%d" % l]).double().reshape(4)
        if l > 0:
            rec[l, 0:4] = _cpu(lstm["weight_ih_l%d" % l]).double().reshape(4)
            rec[l, 8:12] = _cpu(lstm["bias_ih_l%d" % l]).double() + _cpu(lstm["bias_hh_l%d" % l]).double()
    return inproj, rec.to(torch.float32).contiguous().to(device)


def pack_lstm_whh(lstm_re, lstm_im, layer, device):
    return torch.stack((_cpu(lstm_re["weight_hh_l%d" % layer]), _cpu(lstm_im["weight_hh_l%d" % layer]))) \
        .to(torch.float32).contiguous().to(device)


def pack_lstm_bias_tc(lstm_re, lstm_im, layer, n_cols, n_ctas, device):
    """b_ih + b_hh of one layer in the CTA-major order of pack_lstm_whh_tc: fp32 [2][n_ctas][n_cols]."""
    hs = n_cols // 4
    out = []
    for mod in (lstm_re, lstm_im):
        b = (_cpu(mod["bias_ih_l%d" % layer]).double() + _cpu(mod["bias_hh_l%d" % layer]).double())
        out.append(b.view(4, n_ctas, hs).permute(1, 0, 2).reshape(n_ctas, n_cols))
    return torch.stack(out).to(torch.float32).contiguous().to(device)


def pack_lstm_whh_tc(lstm_re, lstm_im, layer, n_cols, n_ctas, device, kind="hh"):
    """Recurrent (kind='hh') or layer>=1 input (kind='ih', square H x H) weights for the tensor-core LSTM
    kernels: bf16 [2 (hi,lo)][2 (module)][n_ctas][n_cols][H], CTA c holds the rows W[gate*H + c*Hs + j] at
    (gate*Hs + j), Hs = n_cols / 4."""
    hs = n_cols // 4
    mats = []
    for mod in (lstm_re, lstm_im):
        w = _cpu(mod["weight_%s_l%d" % (kind, layer)]).to(torch.float32)          # (4H, H)
        H = w.shape[1]
        assert hs * n_ctas == H
        mats.append(w.view(4, n_ctas, hs, H).permute(1, 0, 2, 3).reshape(n_ctas, n_cols, H))
    w = torch.stack(mats)                                                 # (2, n_ctas, n_cols, H)
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    return torch.stack((hi, lo)).contiguous().to(device)


def pack_lstm_cluster_tc(lstm_re, lstm_im, layer, upc, cs, device, kind="hh"):
    """Weights for the small-batch cluster recurrence (csrc/lstm_cluster_tc.cu): bf16 [2 (hi,lo)][2 (module)][cs][4*upc][H],
    CTA c holds row W[gate*H + c*upc + j] at 4*j + gate (the four gates of a hidden unit in neighbouring TMEM lanes)."""
    mats = []
    for mod in (lstm_re, lstm_im):
        w = _cpu(mod["weight_%s_l%d" % (kind, layer)]).to(torch.float32)          # (4H, H)
        H = w.shape[1]
        assert upc * cs == H
        mats.append(w.view(4, cs, upc, H).permute(1, 2, 0, 3).reshape(cs, 4 * upc, H))
    w = torch.stack(mats)
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    return torch.stack((hi, lo)).contiguous().to(device)


def pack_lstm_cluster_bias(lstm_re, lstm_im, layer, upc, cs, device):
    """b_ih + b_hh of one layer in the lane order of pack_lstm_cluster_tc: fp32 [2][cs][128]."""
    out = []
    for mod in (lstm_re, lstm_im):
        b = (_cpu(mod["bias_ih_l%d" % layer]).double() + _cpu(mod["bias_hh_l%d" % layer]).double())
        o = torch.zeros(cs, 128, dtype=torch.float64, device=b.device)
        o[:, :4 * upc] = b.view(4, cs, upc).permute(1, 2, 0).reshape(cs, 4 * upc)
        out.append(o)
    return torch.stack(out).to(torch.float32).contiguous().to(device)


def pack_dense(w_read, b_read, w_imag, b_imag, c_out, f_out, device):
    """ComplexDense (no cross terms, model/complex_progress.py:L83-89) followed by the reshape/permute
    to (B, C, F, T) (model/pvae_module.py:L2085-2088): output feature n = c*f_out + f goes to plane f,
    channel part*ch_c + c.  Input: z planes [1][R][2*ch_z]."""
    zdim = w_read.shape[1]
    ch_z, ch_c = round8(zdim), round8(c_out)
    N = ch_c
    dev = _cpu(w_read).device
    W = torch.zeros(2, f_out, ch_z, N, dtype=torch.float64, device=dev)
    bias = torch.zeros(2, f_out, N, dtype=torch.float64, device=dev)
    for part, (w, b) in enumerate(((w_read, b_read), (w_imag, b_imag))):
        w = _cpu(w).double().reshape(c_out, f_out, zdim)             # [c][f][k]
        W[part, :, :zdim, :c_out] = w.permute(1, 2, 0)
        bias[part, :, :c_out] = _cpu(b).double().reshape(c_out, f_out).t()
    units, taps = [], []
    for part in range(2):
        for f in range(f_out):
            taps.append([0, 0, 0, part * ch_z, ch_z, (part * f_out + f) * ch_z * N])
            units.append([len(taps) - 1, 1, f, part * ch_c, (part * f_out + f) * N, 0])
    return TapGemmPack(W.reshape(-1), bias.reshape(-1), units, taps, N, f_out, 2 * ch_c, False, 0.0, device)


def hann_periodic(win):
    n = torch.arange(win, dtype=torch.float64)
    return 0.5 * (1.0 - torch.cos(2 * math.pi * n / win))


def pack_stft_basis(n_fft, win, device):
    """[win][ncol_pad] analysis basis: column 2k = cos(2 pi (j+off) k / n_fft) w[j], 2k+1 = -sin(..) w[j]."""
    nb = n_fft // 2 + 1
    ncol = 2 * nb
    ncol_pad = (ncol + 127) // 128 * 128
    off = (n_fft - win) // 2
    j = torch.arange(win, dtype=torch.float64) + off
    k = torch.arange(nb, dtype=torch.float64)
    ang = 2 * math.pi * torch.outer(j, k) / n_fft
    w = hann_periodic(win)[:, None]
    basis = torch.zeros(win, ncol_pad, dtype=torch.float64)
    basis[:, 0:ncol:2] = torch.cos(ang) * w
    basis[:, 1:ncol:2] = -torch.sin(ang) * w
    return basis.to(torch.float32).contiguous().to(device)


def pack_istft_basis(n_fft, win, device):
    """[kpad][ncol_pad] synthesis basis (rows 2k / 2k+1 = re / im of bin k) and w^2 [win]."""
    nb = n_fft // 2 + 1
    kpad = (2 * nb + 15) // 16 * 16
    ncol_pad = (win + 127) // 128 * 128
    off = (n_fft - win) // 2
    j = torch.arange(win, dtype=torch.float64) + off
    k = torch.arange(nb, dtype=torch.float64)
    ck = torch.full((nb,), 2.0, dtype=torch.float64)
    ck[0] = 1.0
    ck[-1] = 1.0
    ang = 2 * math.pi * torch.outer(k, j) / n_fft
    w = hann_periodic(win)[None, :]
    basis = torch.zeros(kpad, ncol_pad, dtype=torch.float64)
    basis[0:2 * nb:2, :win] = ck[:, None] / n_fft * torch.cos(ang) * w
    basis[1:2 * nb:2, :win] = -ck[:, None] / n_fft * torch.sin(ang) * w
    wsq = (hann_periodic(win) ** 2)
    return basis.to(torch.float32).contiguous().to(device), wsq.to(torch.float32).to(device)


def _split_rows(w):
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    return torch.stack((hi, lo)).contiguous()


def pack_stft_tc(n_fft, win, device):
    """Analysis basis for the tensor-core STFT: one tap, K = win padded to 64, N = 2*nbins padded to 256.
    wt: bf16 [2][1][N][kpad] (row n = 2k+part of pack_stft_basis)."""
    nb = n_fft // 2 + 1
    kpad = (win + 63) // 64 * 64
    N = (2 * nb + 255) // 256 * 256
    basis = pack_stft_basis(n_fft, win, "cpu")[:, :2 * nb]            # [win][2nb]
    w = torch.zeros(N, kpad)
    w[:2 * nb, :win] = basis.t()
    return dict(wt=_split_rows(w.unsqueeze(0)).to(device), kc_max=kpad, n_slots=1, N=N, kpad=kpad, nbins=nb,
                bias=torch.zeros(N, device=device),
                taps=torch.tensor([[0, 0, 0, 0, kpad, 0]], dtype=torch.int32, device=device),
                units=torch.tensor([[0, 1, 0, 0, 0, kpad // 64]], dtype=torch.int32, device=device))


def pack_istft_tc(n_fft, win, device):
    """Synthesis basis for the tensor-core iSTFT: K = 2*nbins padded to 64, N = win padded to 256."""
    nb = n_fft // 2 + 1
    kpad = (2 * nb + 63) // 64 * 64
    N = (win + 255) // 256 * 256
    basis, wsq = pack_istft_basis(n_fft, win, "cpu")                  # [>=2nb][>=win]
    w = torch.zeros(N, kpad)
    w[:win, :2 * nb] = basis[:2 * nb, :win].t()
    return dict(wt=_split_rows(w.unsqueeze(0)).to(device), kc_max=kpad, n_slots=1, N=N, kpad=kpad, nbins=nb,
                bias=torch.zeros(N, device=device), wsq=wsq.to(device),
                taps=torch.tensor([[0, 0, 0, 0, kpad, 0]], dtype=torch.int32, device=device),
                units=torch.tensor([[0, 1, 0, 0, 0, kpad // 64]], dtype=torch.int32, device=device))


def pack_istft_adjoint_tc(n_fft, win, device):
    """Adjoint of pack_istft_tc's DFT GEMM (backward of the iSTFT): gradient of the frames [R][kpad >= win] times the
    synthesis basis transposed -> gradient rows [R][N >= 2 nbins] in idv_spec_rows_split's order (2k + part)."""
    nb = n_fft // 2 + 1
    kpad = (win + 63) // 64 * 64
    N = (2 * nb + 255) // 256 * 256
    basis, wsq = pack_istft_basis(n_fft, win, "cpu")                  # [>= 2nb][>= win]
    w = torch.zeros(N, kpad)
    w[:2 * nb, :win] = basis[:2 * nb, :win]
    return dict(wt=_split_rows(w.unsqueeze(0)).to(device), kc_max=kpad, n_slots=1, N=N, kpad=kpad, nbins=nb,
                bias=torch.zeros(N, device=device), wsq=wsq.to(device),
                taps=torch.tensor([[0, 0, 0, 0, kpad, 0]], dtype=torch.int32, device=device),
                units=torch.tensor([[0, 1, 0, 0, 0, kpad // 64]], dtype=torch.int32, device=device))


# ------------------------------------------------------------------------------------------------------------------
# backward (phase-1 training step): data-gradient packs and weight-gradient tap tables
# ------------------------------------------------------------------------------------------------------------------
def raw_block_weights(conv_re_w, conv_im_w, transposed=False):
    """Block-real weights W (taps, 2*ch_in, 2*ch_out) of a complex (transposed) conv WITHOUT the CBN fold (the
    training path keeps ComplexBatchNormal separate).  Returns (W, kh, kw, cin, cout)."""
    wr, wi = _cpu(conv_re_w), _cpu(conv_im_w)
    if transposed:
        cin, cout, kh, kw = wr.shape
        m_re, m_im = wr.permute(2, 3, 0, 1).reshape(kh * kw, cin, cout), wi.permute(2, 3, 0, 1).reshape(kh * kw, cin, cout)
    else:
        cout, cin, kh, kw = wr.shape
        m_re, m_im = wr.permute(2, 3, 1, 0).reshape(kh * kw, cin, cout), wi.permute(2, 3, 1, 0).reshape(kh * kw, cin, cout)
    Z, bp = _identity_fold(cout, wr.device)
    zero = torch.zeros(cout, device=wr.device)
    W, _ = _block_weights(m_re, m_im, zero, zero, Z, bp, round8(cin), round8(cout))
    return W, kh, kw, cin, cout


def pack_conv_dgrad(conv_re_w, conv_im_w, f_in, stride_f, pad_f, pad_t, device):
    """Data gradient of the complex conv packed by pack_conv: dx[fi][r] = sum over (fo, kf, kt) with
    fi = stride_f*fo + kf - pad_f of dy[fo][r + (pad_t - kt)] W_tap^T - a tap-GEMM over the planes of dy with the
    mirrored time shifts (the transposed conv of model/complex_progress.py:L16-22's conv)."""
    W, kh, kw, cin, cout = raw_block_weights(conv_re_w, conv_im_w)
    f_out = (f_in + 2 * pad_f - kh) // stride_f + 1
    K, N = 2 * round8(cout), 2 * round8(cin)
    Wt = W.transpose(1, 2).contiguous()                              # (taps, 2ch_out, 2ch_in)
    units, taps = [], []
    for fi in range(f_in):
        begin = len(taps)
        for kf in range(kh):
            num = fi + pad_f - kf
            if num % stride_f:
                continue
            fo = num // stride_f
            if fo < 0 or fo >= f_out:
                continue
            for kt in range(kw):
                taps.append([0, fo, -(pad_t - kt), 0, K, (kf * kw + kt) * K * N])
        units.append([begin, len(taps) - begin, fi, 0, 0, 0])
    p = TapGemmPack(Wt.reshape(-1), dev_zeros(N, device), units, taps, N, f_in, N, False, 0.0, device)
    p.f_out, p.c_out = f_in, cin
    return p


def wgrad_conv_tables(f_in, f_out, kh, kw, stride_f, pad_f, pad_t, rpad, groups, device, transposed=False):
    """Tap tables of the weight-gradient GEMM of a complex conv: unit (kf, kt, g) sums, over the output planes fo of
    group g, dyT[fo] (rows = output channels, K = rows of the activation, shifted by the tap's time offset: source
    0 = shift 0, source 1 = shift 1) against xT[fi] (weight-operand slot fi): out[unit] = dW_tap^T (2ch_out, 2ch_in).
    transposed: the complex transposed conv of pack_conv_transpose (fi = (fo + pad_f - kf) / stride_f, time tap kt
    pairs x[t - kt] with dy[t])."""
    def plane_in(fo, kf):
        if not transposed:
            return stride_f * fo + kf - pad_f
        num = fo + pad_f - kf
        return num // stride_f if num % stride_f == 0 else -1

    units, taps = [], []
    # enough consecutive output planes per group that every group has an in-range input plane for every kf
    min_per = 4 if transposed else 3
    per = max(min_per, (f_out + groups - 1) // groups)
    bounds = list(range(0, f_out, per)) + [f_out]
    if len(bounds) > 2 and bounds[-1] - bounds[-2] < min_per:
        del bounds[-2]                                                # merge a short last group into its neighbour
    groups = len(bounds) - 1
    for kf in range(kh):
        for kt in range(kw):
            dt = kt if transposed else pad_t - kt
            if dt not in (0, 1):
                raise NotImplementedError("weight gradients are built for time shifts 0 / 1 (causal taps)")
            for g in range(groups):
                begin = len(taps)
                for fo in range(bounds[g], bounds[g + 1]):
                    fi = plane_in(fo, kf)
                    if 0 <= fi < f_in:
                        taps.append([dt, fo, 0, 0, rpad, fi])
                if len(taps) == begin:
                    raise RuntimeError("empty weight-gradient unit (kf %d, group %d)" % (kf, g))
                units.append([begin, len(taps) - begin, len(units), 0, 0, (len(taps) - begin) * (rpad // 64)])
    return (dev_table(units, device),
            dev_table(taps, device), len(units), groups)


def unfold_conv_wgrad(dwt, kh, kw, cin, cout, transposed=False):
    """dwt: (kh*kw, 2ch_out, 2ch_in) = dW_tap^T of the block-real weights -> gradients of conv_re.weight and
    conv_im.weight (Cout, Cin, kh, kw) - transposed: tconv_re / tconv_im.weight (Cin, Cout, kh, kw):
    W[:cin,:cout] = m_re, W[ch_in+ci, co] = -m_im, W[ci, ch_out+co] = m_im, W[ch_in+ci, ch_out+co] = m_re
    (model/complex_progress.py:L17-19, L245-247)."""
    ch_in, ch_out = round8(cin), round8(cout)
    d = dwt.transpose(1, 2)                                           # (taps, 2ch_in, 2ch_out)
    d_re = d[:, :cin, :cout] + d[:, ch_in:ch_in + cin, ch_out:ch_out + cout]
    d_im = d[:, :cin, ch_out:ch_out + cout] - d[:, ch_in:ch_in + cin, :cout]
    perm = (2, 3, 0, 1) if transposed else (3, 2, 0, 1)
    f = lambda m: m.reshape(kh, kw, cin, cout).permute(*perm).contiguous()
    return f(d_re), f(d_im)


def pack_convT_dgrad(t_re_w, t_im_w, c0, cn, f_in, stride_f, pad_f, device):
    """Data gradient w.r.t. input channels [c0, c0 + cn) (the running activation or the skip tensor) of the complex
    transposed conv packed by pack_conv_transpose: dx[fi][r] = sum over (kf, kt) of dy[stride_f*fi - pad_f + kf][r + kt]
    W_tap^T - the strided conv that model/complex_progress.py:L244-250's transposed conv is the adjoint of."""
    W, kh, kw, cin, cout = raw_block_weights(t_re_w[c0:c0 + cn], t_im_w[c0:c0 + cn], transposed=True)
    f_out = (f_in - 1) * stride_f - 2 * pad_f + kh
    K, N = 2 * round8(cout), 2 * round8(cn)
    Wt = W.transpose(1, 2).contiguous()                              # (taps, 2ch_out, 2ch_in)
    units, taps = [], []
    for fi in range(f_in):
        begin = len(taps)
        for kf in range(kh):
            fo = stride_f * fi - pad_f + kf
            if fo < 0 or fo >= f_out:
                continue
            for kt in range(kw):
                taps.append([0, fo, -kt, 0, K, (kf * kw + kt) * K * N])
        units.append([begin, len(taps) - begin, fi, 0, 0, 0])
    p = TapGemmPack(Wt.reshape(-1), dev_zeros(N, device), units, taps, N, f_in, N, False, 0.0, device)
    p.f_out, p.c_out = f_in, cn
    return p


def pack_dense_dgrad(w_read, w_imag, c_out, f_out, device):
    """Gradient of z from the gradient of the ComplexDense output planes [f_out][R][2 ch_c] (pack_dense's layout):
    dz_part[r][k] = sum over (f, c) of g[f][r][part*ch_c + c] W_part[c*f_out + f][k].  Output: z planes [1][R][2 ch_z]."""
    zdim = w_read.shape[1]
    ch_z, ch_c = round8(zdim), round8(c_out)
    dev = _cpu(w_read).device
    W = torch.zeros(2, f_out, ch_c, ch_z, dtype=torch.float64, device=dev)
    for part, w in enumerate((w_read, w_imag)):
        W[part, :, :c_out, :zdim] = _cpu(w).double().reshape(c_out, f_out, zdim).permute(1, 0, 2)
    units, taps = [], []
    for part in range(2):
        begin = len(taps)
        for f in range(f_out):
            taps.append([0, f, 0, part * ch_c, ch_c, (part * f_out + f) * ch_c * ch_z])
        units.append([begin, f_out, 0, part * ch_z, 0, 0])
    return TapGemmPack(W.reshape(-1), dev_zeros(ch_z, device), units, taps, ch_z, 1, 2 * ch_z, False, 0.0, device)


def wgrad_dense_tables(f_out, rpad, device):
    """Weight-gradient GEMM of ComplexDense: unit (f, part) = gT[f] (rows = the 2 ch_c output channels of plane f,
    K = rows of the sequence) against slot `part` of zT ([2 ch_z][rpad] viewed as 2 slots of ch_z rows)."""
    units, taps = [], []
    ks = rpad // 64
    for f in range(f_out):
        for part in range(2):
            taps.append([0, f, 0, 0, rpad, part])
            units.append([len(taps) - 1, 1, len(units), 0, 0, ks])
    return (dev_table(units, device),
            dev_table(taps, device), len(units))


def pack_lstm_gates(lstm_re, lstm_im, hidden, layer, c_in, f_in, device):
    """Gate pre-activations of one nn.LSTM layer for ALL time steps as one two-source tap-GEMM (backward pass: the
    forward kernels keep them on chip).  Unit s = (module m, part p), N = 4H:
      layer 0: source 0 = the encoder output planes (f_in planes, channel half p), source 1 = this layer's h
               planes [4][R][H] read one row back (h(t-1); the pad row is h(-1) = 0);
      layer l: source 0 = h planes of the layer below, source 1 = this layer's h planes one row back."""
    H = hidden
    N = 4 * H
    dev = _cpu(lstm_re["weight_hh_l%d" % layer]).device
    mats, bias = [], torch.zeros(2 * N, dtype=torch.float64, device=dev)
    units, taps = [], []
    off = 0
    offs = {}
    for m, mod in enumerate((lstm_re, lstm_im)):
        wih = _cpu(mod["weight_ih_l%d" % layer]).double()
        if layer == 0:
            ch = round8(c_in)
            w = torch.zeros(f_in, ch, N, dtype=torch.float64, device=dev)
            w[:, :c_in] = wih.reshape(N, c_in, f_in).permute(2, 1, 0)
            for f in range(f_in):
                offs[("ih", m, f)] = off
                mats.append(w[f].reshape(-1))
                off += ch * N
        else:
            offs[("ih", m)] = off
            mats.append(wih.t().contiguous().reshape(-1))
            off += H * N
        offs[("hh", m)] = off
        mats.append(_cpu(mod["weight_hh_l%d" % layer]).double().t().contiguous().reshape(-1))
        off += H * N
        bias[m * N:(m + 1) * N] = _cpu(mod["bias_ih_l%d" % layer]).double() + _cpu(mod["bias_hh_l%d" % layer]).double()
    for m in range(2):
        for p in range(2):
            s = m * 2 + p
            begin = len(taps)
            if layer == 0:
                ch = round8(c_in)
                for f in range(f_in):
                    taps.append([0, f, 0, p * ch, ch, offs[("ih", m, f)]])
            else:
                taps.append([0, s, 0, 0, H, offs[("ih", m)]])
            taps.append([1, s, 1, 0, H, offs[("hh", m)]])
            units.append([begin, len(taps) - begin, s, 0, m * N, 0])
    return TapGemmPack(torch.cat(mats), bias, units, taps, N, 4, N, False, 0.0, device)


def bptt_split_k(hidden, sms=148):
    """Split-K factor of the per-step dh = dP W_hh GEMM (2 modules x H/64 column tiles x split CTAs <= SMs)."""
    ksteps = 4 * hidden // 64
    best = 1
    for s in range(1, ksteps + 1):
        if ksteps % s == 0 and 2 * max(1, hidden // 64) * s <= sms:
            best = s
    return best


def pack_lstm_dgrad(lstm_re, lstm_im, hidden, layer, kind, device, c_in=0, f_in=0, split_k=0):
    """Gate gradients dP [4][rows][4H] times a weight matrix (K = 4H):
      kind 'hh'          -> dh(t-1) = dP W_hh^m, N = H, out [4][rows][H];
      kind 'hh', split_k -> the one-time-step form of the BPTT loop: the two streams of a module are ONE plane of
                            2*rows rows (they share W_hh^m), K is cut into split_k slices so that 2 x H/64 x split_k
                            CTAs stream the weights instead of 4 x H/64: out [split_k][4][rows][H] partial sums;
      kind 'ih', layer>0 -> gradient of the layer below's h, N = H, out [4][rows][H];
      kind 'ih', layer 0 -> gradient of the encoder output planes: unit (f, p) sums the two modules,
                            N = ch, out planes [f_in][rows][2ch] at channel offset p*ch."""
    H = hidden
    K = 4 * H
    if kind == "ih" and layer == 0:
        ch = round8(c_in)
        mats, units, taps = [], [], []
        for m, mod in enumerate((lstm_re, lstm_im)):
            w = _cpu(mod["weight_ih_l0"]).double().reshape(K, c_in, f_in)
            for f in range(f_in):
                wf = torch.zeros(K, ch, dtype=torch.float64, device=w.device)
                wf[:, :c_in] = w[:, :, f]
                mats.append(wf.reshape(-1))
        for f in range(f_in):
            for p in range(2):
                begin = len(taps)
                for m in range(2):
                    taps.append([0, m * 2 + p, 0, 0, K, (m * f_in + f) * K * ch])
                units.append([begin, 2, f, p * ch, 0, 0])
        return TapGemmPack(torch.cat(mats), dev_zeros(ch, device), units, taps, ch, f_in, 2 * ch, False, 0.0, device)
    mats, units, taps = [], [], []
    for m, mod in enumerate((lstm_re, lstm_im)):
        mats.append(_cpu(mod["weight_%s_l%d" % (kind, layer)]).double().reshape(-1))       # (4H, H) = [K][N]
    if split_k:
        if kind != "hh" or K % (64 * split_k):
            raise ValueError("split_k must divide the 64-wide K steps of the recurrent matrix")
        kc = K // split_k
        for j in range(split_k):
            for m in range(2):
                taps.append([0, m, 0, j * kc, kc, m * K * H + j * kc * H])
                units.append([len(taps) - 1, 1, j * 2 + m, 0, 0, 0])
        return TapGemmPack(torch.cat(mats), dev_zeros(H, device), units, taps, H, 2 * split_k, H, False, 0.0, device)
    for m in range(2):
        for p in range(2):
            taps.append([0, m * 2 + p, 0, 0, K, m * K * H])
            units.append([len(taps) - 1, 1, m * 2 + p, 0, 0, 0])
    return TapGemmPack(torch.cat(mats), dev_zeros(H, device), units, taps, H, 4, H, False, 0.0, device)


def wgrad_lstm_tables(rpad, device, f_in=0):
    """Weight-gradient GEMM of an LSTM matrix: unit m sums the two parts p: dPT[(m,p)] (rows = 4H gate columns,
    K = rows of the sequence) against the transposed input slot.  f_in == 0: slots are the 4 streams (h planes) ->
    out [2][4H][H]; f_in > 0 (layer-0 W_ih): the encoder planes transposed and viewed as slots f*2 + p of ch
    channels -> units (m, f), out [2*f_in][4H][ch]."""
    units, taps = [], []
    ks = rpad // 64
    if f_in == 0:
        for m in range(2):
            for p in range(2):
                taps.append([0, m * 2 + p, 0, 0, rpad, m * 2 + p])
            units.append([2 * m, 2, m, 0, 0, 2 * ks])
    else:
        for m in range(2):
            for f in range(f_in):
                begin = len(taps)
                for p in range(2):
                    taps.append([0, m * 2 + p, 0, 0, rpad, f * 2 + p])
                units.append([begin, 2, m * f_in + f, 0, 0, 2 * ks])
    return (dev_table(units, device),
            dev_table(taps, device), len(units))
