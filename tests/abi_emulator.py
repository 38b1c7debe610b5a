"""TEST DOUBLE of the C ABI (include/idv.h) in plain torch on the CPU.

Each function restates the *contract* written in the header (not the kernels) so that the host side
of the package — weight packing, CBN folding, tap tables, plane layout, per-sample decoder passes,
module wiring, state_dict handling — can be checked against the reference goldens in the CPU-only
test tier.  It is injected by the ``emulated_abi`` fixture (tests/conftest.py) via monkeypatching
``lib.call``; the product never imports this file and has no CPU path of its own.
"""
import math

import torch

D = torch.float64


def _tv(t_valid, T):
    """Valid frames per utterance: 0 or >= T means all T (include/idv.h, "Valid frames")."""
    return t_valid if 0 < t_valid < T else T


def _mask_rows(acc, R, Tp, t_valid):
    """Pad rows and rows of frames beyond the valid length are written as zero."""
    Tp = abs(Tp)
    if Tp > 0:
        tt = torch.arange(R) % Tp
        acc[(tt == 0) | (tt > _tv(t_valid, Tp - 1))] = 0
    return acc


def _written_rows(R, Tp):
    """Tp < 0 (streaming): pad rows are not written at all (they carry state)."""
    if Tp < 0:
        return (torch.arange(R) % (-Tp)) != 0
    return torch.ones(R, dtype=torch.bool)


def _flat(t):
    return t.view(-1)


def _rd(t, split, n=None):
    """Logical fp64 values of an activation tensor: fp32 [n] or split bf16 [2][n] (hi + lo)."""
    f = _flat(t)
    if not split:
        return f.to(D) if n is None else f[:n].to(D)
    n = f.numel() // 2 if n is None else n
    half = f.numel() // 2
    return f[:n].to(D) + f[half:half + n].to(D)


def _wr(t, split, vals):
    """Store logical values (flat, fp64/fp32) into an activation tensor in its format."""
    f = _flat(t)
    v = vals.reshape(-1).to(torch.float32)
    if not split:
        f[:v.numel()] = v
        return
    half = f.numel() // 2
    hi = v.to(torch.bfloat16)
    lo = (v - hi.to(torch.float32)).to(torch.bfloat16)
    f[:v.numel()] = hi
    f[half:half + v.numel()] = lo


def idv_tapgemm_tc_head(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, N, units, taps,
                        n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, slope, head, head_fout,
                        head_bmul, head_boff, stft_x, predict, t_valid=0):
    if head == 3:                                       # STFT epilogue: column pairs -> (B, nbins, Tp, 2)
        tmp = torch.zeros(R * N)
        idv_tapgemm_tc(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, 0, wt, kc_max, n_slots, bias, N, units, taps,
                       n_units, tmp, N, R * N, 0, 0, 0, 0.0)
        T = Tp
        B = R // T
        vals = tmp.view(B, T, N)[:, :, :2 * head_fout]
        predict.view(B, head_fout, T, 2).copy_(vals.reshape(B, T, head_fout, 2).permute(0, 2, 1, 3))
        if out is not None:                               # split-bf16 activation rows (one plane, causal row layout)
            rows = _flat(out).view(2, B, T + 1, out_ld)
            hi = vals.to(torch.bfloat16)
            lo = (vals - hi.to(torch.float32)).to(torch.bfloat16)
            rows[0][:, 1:, head_boff:head_boff + 2 * head_fout] = hi        # (the kernel may also write the exact
            rows[1][:, 1:, head_boff:head_boff + 2 * head_fout] = lo        #  zeros of the padding columns)
        return
    assert head in (1, 2) and N == 32
    tmp = torch.zeros(n_units * R * 32)
    un = units.clone()
    un[:, 2] = torch.arange(n_units)          # plane per unit in the scratch output
    un[:, 3] = 0
    idv_tapgemm_tc._head_bias = True          # the head applies bias (re, im) to every bin's column pair
    try:
        idv_tapgemm_tc(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, 0, wt, kc_max, n_slots, bias, N, un, taps,
                       n_units, tmp, 32, R * 32, 0, 0, 1, slope)
    finally:
        idv_tapgemm_tc._head_bias = False
    T = Tp - 1
    NB = R // Tp
    acc = tmp.view(n_units, NB, Tp, 32)[:, :, 1:].to(D)             # (units, NB, T, 32), bias + PReLU applied
    pv = predict.view(-1, head_fout, T, 2)
    for ui, (tb, nt, fo0, nbins, _, _) in enumerate(units.tolist()):
        assert 1 <= nbins <= 16
        for e in range(nbins):
            fo = fo0 + e
            if fo >= head_fout:
                continue
            yr, yi = acc[ui, :, :, 2 * e], acc[ui, :, :, 2 * e + 1]
            if head == 2:
                mag = torch.tanh(torch.sqrt(yr ** 2 + yi ** 2))
                ph = torch.atan2(yi / (mag + 1e-8), yr / (mag + 1e-8))
                X = stft_x.view(-1, head_fout, T, 2)[:NB, fo].to(D)
                in_mag = torch.sqrt(X[..., 0] ** 2 + X[..., 1] ** 2)
                in_ph = torch.atan2(X[..., 1], X[..., 0])
                yr, yi = in_mag * mag * torch.cos(in_ph + ph), in_mag * mag * torch.sin(in_ph + ph)
            val = torch.stack((yr, yi), -1).to(torch.float32)
            pv[head_boff::head_bmul][:NB, fo] = val
            if out is not None:                           # split-bf16 K-major spectrum rows for the iSTFT GEMM
                rows = _flat(out).view(2, -1, T, out_ld)
                hi = val.to(torch.bfloat16)
                lo = (val - hi.to(torch.float32)).to(torch.bfloat16)
                rows[0][head_boff::head_bmul][:NB, :, 2 * fo:2 * fo + 2] = hi
                rows[1][head_boff::head_bmul][:NB, :, 2 * fo:2 * fo + 2] = lo


def idv_tapgemm_tc_b2(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, bias_first, N, units,
                      taps, n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, slope, t_valid=0):
    """idv_tapgemm_tc whose rows r % Tp == 1 (first frame of every utterance) use ``bias_first`` instead of ``bias``."""
    idv_tapgemm_tc._bias_first = bias_first
    try:
        idv_tapgemm_tc(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, N, units, taps,
                       n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, slope, t_valid)
    finally:
        idv_tapgemm_tc._bias_first = None


def idv_tapgemm_tc_splitk(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, bias_first, N,
                          units, taps, n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, slope, t_valid=0,
                          min_ksteps=0):
    """Contract = idv_tapgemm_tc / idv_tapgemm_tc_b2 (how the K range is divided over CTAs is not part of it)."""
    if bias_first is not None:
        return idv_tapgemm_tc_b2(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, bias_first,
                                 N, units, taps, n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, slope,
                                 t_valid)
    return idv_tapgemm_tc(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, N, units, taps,
                          n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, slope, t_valid)


def idv_tapgemm_tc(a0, a0_cp, a0_planes, a1, a1_cp, a1_planes, R, Tp, wt, kc_max, n_slots, bias, N, units, taps,
                   n_units, out, out_ld, out_plane, out_hl, out_split, apply_prelu, slope, t_valid=0):
    """Contract of the tensor-core tap-GEMM incl. its arithmetic: a_hi*w_hi + a_hi*w_lo + a_lo*w_hi."""
    taps_l, units_l = taps.tolist(), units.tolist()
    W = wt.view(2, n_slots, N, kc_max).to(D)

    def planes(a, cp, npl):
        v = _flat(a).view(2, npl, R, cp).to(D)
        return v[0], v[1]
    A0 = planes(a0, a0_cp, a0_planes)
    A1 = planes(a1, a1_cp, a1_planes) if a1 is not None else None
    res = torch.zeros(out.numel() // (2 if out_split else 1), dtype=D)
    res[:] = float("nan")
    written = torch.zeros_like(res, dtype=torch.bool)
    for (tap_begin, n_taps, out_f, out_ch_off, bias_off, ksteps) in units_l:
        acc = torch.zeros(R, N, dtype=D)
        assert ksteps == sum(t[4] // 64 for t in taps_l[tap_begin:tap_begin + n_taps])
        for (src, f_in, dt, ch_off, kc, slot) in taps_l[tap_begin:tap_begin + n_taps]:
            assert kc % 64 == 0
            hi, lo = (A0 if src == 0 else A1)

            def shifted(x):
                x = x[f_in]
                lo_c, hi_c = max(ch_off, 0), min(ch_off + kc, x.shape[1])      # channels outside the tensor read as zero
                x = torch.nn.functional.pad(x[:, lo_c:hi_c], (lo_c - ch_off, ch_off + kc - hi_c))
                if dt > 0:
                    x = torch.cat((torch.zeros(dt, kc, dtype=D), x[:R - dt]), 0)
                elif dt < 0:
                    x = torch.cat((x[-dt:], torch.zeros(-dt, kc, dtype=D)), 0)
                return x
            ah, al = shifted(hi), shifted(lo)
            wh, wl = W[0, slot, :, :kc].t(), W[1, slot, :, :kc].t()
            acc += ah @ wh + ah @ wl + al @ wh
        bvec = _flat(bias)[bias_off:bias_off + N].to(D).clone()
        if getattr(idv_tapgemm_tc, "_head_bias", False):
            bvec = bvec[0:2].repeat(N // 2)
        acc += bvec
        b1 = getattr(idv_tapgemm_tc, "_bias_first", None)
        if b1 is not None:
            first = (torch.arange(R) % abs(Tp)) == 1
            acc[first] += (_flat(b1)[bias_off:bias_off + N].to(D) - bvec)
        if apply_prelu:
            acc = torch.where(acc > 0, acc, slope * acc)
        _mask_rows(acc, R, Tp, t_valid)
        if N > out_ld:                                      # columns wrap into consecutive output planes
            assert out_ld % 32 == 0 and N % out_ld == 0 and out_ch_off == 0
            for j in range(N // out_ld):
                o0 = (out_f + j) * out_plane
                if out_hl > 0 and o0 + out_plane > out_hl:
                    continue
                res[o0:o0 + R * out_ld].view(R, out_ld)[:] = acc[:, j * out_ld:(j + 1) * out_ld]
                written[o0:o0 + R * out_ld].view(R, out_ld)[:] = _written_rows(R, Tp)[:, None]
            continue
        view = res[out_f * out_plane:out_f * out_plane + R * out_ld].view(R, out_ld)
        view[:, out_ch_off:out_ch_off + N] = acc
        written[out_f * out_plane:out_f * out_plane + R * out_ld].view(R, out_ld)[:, out_ch_off:out_ch_off + N] = \
            _written_rows(R, Tp)[:, None]
    # only touch what the kernel writes
    f = _flat(out)
    vals = res[written].to(torch.float32)
    if out_split:
        half = f.numel() // 2
        assert half == out_hl
        hi = vals.to(torch.bfloat16)
        lo = (vals - hi.to(torch.float32)).to(torch.bfloat16)
        f[:half][written] = hi
        f[half:][written] = lo
    else:
        f[written] = vals


def idv_tapgemm_f32(a0, a0_ld, a0_plane, a1, a1_ld, a1_plane, R, Tp, w, bias, N, units, taps, n_units, out,
                    out_ld, out_plane, apply_prelu, slope, t_valid=0):
    taps_l, units_l = taps.tolist(), units.tolist()
    assert len(units_l) == n_units
    for (tap_begin, n_taps, out_f, out_ch_off, bias_off, _) in units_l:
        acc = torch.zeros(R, N, dtype=D)
        for (src, f_in, dt, ch_off, kc, w_off) in taps_l[tap_begin:tap_begin + n_taps]:
            a, ld, plane = (a0, a0_ld, a0_plane) if src == 0 else (a1, a1_ld, a1_plane)
            A = _flat(a)[f_in * plane:f_in * plane + R * ld].view(R, ld)[:, ch_off:ch_off + kc].to(D)
            if dt > 0:
                A = torch.cat((torch.zeros(dt, kc, dtype=D), A[:R - dt]), 0)
            elif dt < 0:
                A = torch.cat((A[-dt:], torch.zeros(-dt, kc, dtype=D)), 0)
            W = _flat(w)[w_off:w_off + kc * N].view(kc, N).to(D)
            acc += A @ W
        acc += _flat(bias)[bias_off:bias_off + N].to(D)
        if apply_prelu:
            acc = torch.where(acc > 0, acc, slope * acc)
        _mask_rows(acc, R, Tp, t_valid)
        ov = _flat(out)[out_f * out_plane:out_f * out_plane + R * out_ld].view(R, out_ld)
        wr = _written_rows(R, Tp)
        ov[wr, out_ch_off:out_ch_off + N] = acc.to(torch.float32)[wr]


def idv_stft_fwd(x, B, L, basis, n_fft, hop, win, out):
    T = L // hop + 1
    nb = n_fft // 2 + 1
    off = (n_fft - win) // 2
    xp = torch.nn.functional.pad(x.view(B, 1, L).to(D), (n_fft // 2, n_fft // 2), mode="reflect")[:, 0]
    frames = xp.unfold(1, n_fft, hop)[:, :T, off:off + win]                      # (B, T, win)
    y = frames @ basis[:, :2 * nb].to(D)                                          # (B, T, 2nb)
    out.copy_(y.view(B, T, nb, 2).permute(0, 2, 1, 3).to(torch.float32))


def idv_istft_fwd(spec, B, T, basis, wsq, n_fft, hop, win, frames, out):
    nb = n_fft // 2 + 1
    off = (n_fft - win) // 2
    A = spec.view(B, nb, T, 2).permute(0, 2, 1, 3).reshape(B * T, 2 * nb).to(D)
    fr = A @ basis[:2 * nb, :win].to(D)
    frames.view(B * T, win).copy_(fr.to(torch.float32))
    total = n_fft + hop * (T - 1)
    y = torch.zeros(B, total, dtype=D)
    env = torch.zeros(total, dtype=D)
    fr = fr.view(B, T, win)
    for t in range(T):
        y[:, t * hop + off:t * hop + off + win] += fr[:, t]
        env[t * hop + off:t * hop + off + win] += wsq.to(D)
    h = n_fft // 2
    out.copy_((y[:, h:total - h] / env[h:total - h]).to(torch.float32))


def idv_stft_frames_split(x, B, L, n_fft, hop, win, kpad, lengths, out):
    T = L // hop + 1
    off = (n_fft - win) // 2
    fr = torch.zeros(B, T, kpad, dtype=D)
    for b in range(B):
        Lb = L if lengths is None else min(int(lengths[b]), L)
        Tb = Lb // hop + 1
        xp = torch.nn.functional.pad(x.view(B, 1, L)[b:b + 1, :, :Lb].to(D), (n_fft // 2, n_fft // 2), mode="reflect")[0, 0]
        fr[b, :Tb, :win] = xp.unfold(0, n_fft, hop)[:Tb, off:off + win]
    _wr(out, 1, fr)


def idv_spec_rows_split(spec, B, nbins, T, kpad, out):
    rows = torch.zeros(B, T, kpad, dtype=D)
    rows[:, :, :2 * nbins] = spec.view(B, nbins, T, 2).permute(0, 2, 1, 3).reshape(B, T, 2 * nbins).to(D)
    _wr(out, 1, rows)


def idv_ola_fwd(frames, frame_ld, wsq, B, T, n_fft, hop, win, lengths, out):
    off = (n_fft - win) // 2
    fr = frames.view(B, T, frame_ld)[:, :, :win].to(D)
    h = n_fft // 2
    res = torch.zeros(B, hop * (T - 1), dtype=D)
    for b in range(B):
        Tb = T if lengths is None else min(int(lengths[b]) // hop + 1, T)
        total = n_fft + hop * (Tb - 1)
        y = torch.zeros(total, dtype=D)
        env = torch.zeros(total, dtype=D)
        for t in range(Tb):
            y[t * hop + off:t * hop + off + win] += fr[b, t]
            env[t * hop + off:t * hop + off + win] += wsq.to(D)
        res[b, :hop * (Tb - 1)] = y[h:total - h] / env[h:total - h]
    out.copy_(res.to(torch.float32))


def idv_enc0_fwd(stft, B, Fin, T, w, bias, Cout, slope, out, out_split=0, causal=1, t_valid=0, prev=None, keep_pad=0):
    N = 2 * Cout
    Fout = (Fin + 4 - 5) // 2 + 1
    Tp = T + 1
    x = stft.view(B, Fin, T, 2).to(D)
    xpad = torch.zeros(B, Fin + 4, T + 1, 2, dtype=D)        # freq pad 2/2; time: one zero left (causal) / right
    if causal:
        xpad[:, 2:2 + Fin, 1:] = x
        if prev is not None:
            xpad[:, 2:2 + Fin, 0] = prev.view(B, Fin, 2).to(D)
    else:
        xpad[:, 2:2 + Fin, :T] = x
    W = w.view(10, 2, N).to(D)
    res = torch.zeros(Fout, B, Tp, N, dtype=D)
    for kf in range(5):
        for kt in range(2):
            sl = xpad[:, kf:kf + 2 * Fout - 1:2, kt:kt + T]                       # (B, Fout, T, 2)
            res[:, :, 1:] += torch.einsum("bftp,pn->fbtn", sl, W[kf * 2 + kt])
    res[:, :, 1:] += bias.to(D)
    res = torch.where(res > 0, res, slope * res)
    res[:, :, 0] = 0
    res[:, :, 1 + _tv(t_valid, T):] = 0
    if keep_pad:
        old = _rd(out, out_split, res.numel()).view(res.shape)
        res[:, :, 0] = old[:, :, 0]
    _wr(out, out_split, res)


def idv_dec5_head_fwd(p, p_cp, skip, s_cp, in_split, NB, Fin, T, w, bias, slope, mask, stft_x, predict, out_bmul,
                      out_boff):
    Tp = T + 1
    R = NB * Tp
    Fout = 2 * Fin - 1
    if skip is None:
        s_cp = 0
    A = _rd(p, in_split, Fin * R * p_cp).view(Fin, R, p_cp)
    if s_cp:
        A = torch.cat((A, _rd(skip, in_split, Fin * R * s_cp).view(Fin, R, s_cp)), 2)
    W = w.view(10, p_cp + s_cp, 2).to(D)
    y = torch.zeros(Fout, R, 2, dtype=D)
    for fo in range(Fout):
        for kf in range(5):
            if (fo + 2 - kf) % 2:
                continue
            fi = (fo + 2 - kf) // 2
            if fi < 0 or fi >= Fin:
                continue
            for kt in range(2):
                Ash = A[fi] if kt == 0 else torch.cat((torch.zeros(1, A.shape[2], dtype=D), A[fi, :R - 1]), 0)
                y[fo] += Ash @ W[kf * 2 + kt]
    y += bias.to(D)
    y = torch.where(y > 0, y, slope * y)
    y = y.view(Fout, NB, Tp, 2)[:, :, 1:].permute(1, 0, 2, 3)                     # (NB, Fout, T, 2)
    if mask:
        yr, yi = y[..., 0], y[..., 1]
        mag = torch.tanh(torch.sqrt(yr ** 2 + yi ** 2))
        ph = torch.atan2(yi / (mag + 1e-8), yr / (mag + 1e-8))
        X = stft_x.view(-1, Fout, T, 2)[:NB].to(D)
        in_mag = torch.sqrt(X[..., 0] ** 2 + X[..., 1] ** 2)
        in_ph = torch.atan2(X[..., 1], X[..., 0])
        y = torch.stack((in_mag * mag * torch.cos(in_ph + ph), in_mag * mag * torch.sin(in_ph + ph)), -1)
    pv = predict.view(-1, Fout, T, 2)
    pv[out_boff::out_bmul][:NB] = y.to(torch.float32)


def idv_lstm_recurrent_fwd(g, g_m_off, g_p_off, g_ld, whh, NB, T, H, hseq, hsplit, sync, t_valid=0):
    Tp = T + 1
    R = NB * Tp
    hs = hseq.view(4, R, H)
    W = whh.view(2, 4 * H, H).to(D)
    gf = _flat(g)
    rows0 = torch.arange(NB) * Tp
    for m in range(2):
        for p in range(2):
            base = m * g_m_off + p * g_p_off
            h = torch.zeros(NB, H, dtype=D)
            c = torch.zeros(NB, H, dtype=D)
            hs[m * 2 + p, rows0] = 0
            for t in range(_tv(t_valid, T)):
                rows = rows0 + 1 + t
                idx = base + rows[:, None] * g_ld + torch.arange(4 * H)[None, :]
                a = gf[idx].to(D) + h @ W[m].t()
                i, f, gg, o = a[:, :H], a[:, H:2 * H], a[:, 2 * H:3 * H], a[:, 3 * H:]
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                hs[m * 2 + p, rows] = h.to(torch.float32)
    if hsplit is not None:
        _wr(hsplit, 1, hs)


def _bf16_split(x):
    hi = x.to(torch.float32).to(torch.bfloat16)
    lo = (x.to(torch.float32) - hi.to(torch.float32)).to(torch.bfloat16)
    return hi.to(D), lo.to(D)


def idv_lstm_recurrent_tc(g, g_m_off, g_p_off, g_ld, wpack, NB, T, H, hseq, hsplit, hx, sync, t_valid=0):
    """Contract of the tensor-core recurrence: the recurrent product uses the split value of h(t-1)."""
    _lstm_rec_tc(_lstm_tc_config(H), g, g_m_off, g_p_off, g_ld, wpack, NB, T, H, hseq, hsplit, t_valid)


def idv_lstm_layer_pair_tc(g, g_m_off, g_p_off, g_ld, wpack, NB, T, H, hseq, hsplit, work, sync, t_valid=0):
    """Same contract as idv_lstm_recurrent_tc with the weight pack in the CTA-pair kernel's (n_cols, n_ctas) order."""
    from idccrn_b200 import lib
    _lstm_rec_tc(lib.lstm_layer_pair_config(H)[:2], g, g_m_off, g_p_off, g_ld, wpack, NB, T, H, hseq, hsplit, t_valid)


def _lstm_rec_tc(cfg, g, g_m_off, g_p_off, g_ld, wpack, NB, T, H, hseq, hsplit, t_valid):
    n_cols, n_ctas = cfg
    hs = n_cols // 4
    Tp = T + 1
    R = NB * Tp
    wp = wpack.view(2, 2, n_ctas, 4, hs, H).to(D)                     # [hl][m][c][gate][j][k]
    W = wp.permute(0, 1, 3, 2, 4, 5).reshape(2, 2, 4 * H, H)          # [hl][m][gate*H + c*hs + j][k]
    gf = _flat(g)
    rows0 = torch.arange(NB) * Tp
    out = torch.zeros(4, R, H, dtype=D)
    for m in range(2):
        for p in range(2):
            base = m * g_m_off + p * g_p_off
            h = torch.zeros(NB, H, dtype=D)
            c = torch.zeros(NB, H, dtype=D)
            for t in range(_tv(t_valid, T)):
                rows = rows0 + 1 + t
                idx = base + rows[:, None] * g_ld + torch.arange(4 * H)[None, :]
                hh, hl = _bf16_split(h)
                a = gf[idx].to(D) + hh @ W[0, m].t() + hh @ W[1, m].t() + hl @ W[0, m].t()
                i, f, gg, o = a[:, :H], a[:, H:2 * H], a[:, 2 * H:3 * H], a[:, 3 * H:]
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                out[m * 2 + p, rows] = h
    valid = torch.zeros(R, dtype=torch.bool)
    valid[(rows0[:, None] + 1 + torch.arange(_tv(t_valid, T))[None, :]).reshape(-1)] = True
    if hseq is not None:
        hseq.view(4, R, H)[:, valid] = out[:, valid].to(torch.float32)
    if hsplit is not None:
        hi, lo = _bf16_split(out)
        hv = hsplit.view(2, 4, R, H)
        hv[0][:, valid] = hi[:, valid].to(torch.bfloat16)
        hv[1][:, valid] = lo[:, valid].to(torch.bfloat16)


def idv_lstm2_wave_tc(g0, g_m_off, g_p_off, g_ld, w_hh0, w_ih1, w_hh1, bias1, NB, T, H, hseq1, work, sync,
                      t_valid=0):
    """Contract: layer 0 recurrence (split h), layer-1 input projection on the split h0 (3 products) + bias,
    layer-1 recurrence (split h1)."""
    n_cols, n_ctas = 64, H // 16
    hs = n_cols // 4

    def unpack(wp):
        w = wp.view(2, 2, n_ctas, 4, hs, H).to(D)
        return w.permute(0, 1, 3, 2, 4, 5).reshape(2, 2, 4 * H, H)
    b1 = bias1.view(2, n_ctas, 4, hs).to(D).permute(0, 2, 1, 3).reshape(2, 4 * H)
    _lstm2_tc(g0, g_m_off, g_p_off, g_ld, unpack(w_hh0), unpack(w_ih1), unpack(w_hh1), b1, NB, T, H, hseq1, t_valid)


def idv_lstm2_cluster_tc(g0, g_m_off, g_p_off, g_ld, w_hh0, w_ih1, w_hh1, bias1, NB, T, H, hseq1, work, sync,
                         t_valid=0):
    """Same contract as idv_lstm2_wave_tc for NB <= 8, with the packs in the cluster kernel's order: CTA c holds row
    W[gate*H + c*upc + j] at 4*j + gate, bias fp32 [2][cs][128]."""
    from idccrn_b200 import lib
    upc, cs, _ = lib.lstm2_cluster_config(H, NB, T)

    def unpack(wp):
        w = wp.view(2, 2, cs, upc, 4, H).to(D)                         # [hl][m][c][j][gate][k]
        return w.permute(0, 1, 4, 2, 3, 5).reshape(2, 2, 4 * H, H)
    b1 = bias1.view(2, cs, 128)[:, :, :4 * upc].reshape(2, cs, upc, 4).to(D).permute(0, 3, 1, 2).reshape(2, 4 * H)
    _lstm2_tc(g0, g_m_off, g_p_off, g_ld, unpack(w_hh0), unpack(w_ih1), unpack(w_hh1), b1, NB, T, H, hseq1, t_valid)


def _lstm2_tc(g0, g_m_off, g_p_off, g_ld, W0, Wi, W1, b1, NB, T, H, hseq1, t_valid):
    Tp = T + 1
    R = NB * Tp
    gf = _flat(g0)
    rows0 = torch.arange(NB) * Tp
    out = hseq1.view(4, R, H)

    def mm3(x, W, m):
        xh, xl = _bf16_split(x)
        return xh @ W[0, m].t() + xh @ W[1, m].t() + xl @ W[0, m].t()

    def cell(a, c):
        i, f, gg, o = a[:, :H], a[:, H:2 * H], a[:, 2 * H:3 * H], a[:, 3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        return torch.sigmoid(o) * torch.tanh(c), c
    for m in range(2):
        for p in range(2):
            base = m * g_m_off + p * g_p_off
            h0 = torch.zeros(NB, H, dtype=D)
            c0 = torch.zeros(NB, H, dtype=D)
            h1 = torch.zeros(NB, H, dtype=D)
            c1 = torch.zeros(NB, H, dtype=D)
            for t in range(_tv(t_valid, T)):
                rows = rows0 + 1 + t
                idx = base + rows[:, None] * g_ld + torch.arange(4 * H)[None, :]
                h0, c0 = cell(gf[idx].to(D) + mm3(h0, W0, m), c0)
                g1 = (mm3(h0, Wi, m) + b1[m]).to(torch.float32).to(D)        # G1 is stored as fp32
                h1, c1 = cell(g1 + mm3(h1, W1, m), c1)
                out[m * 2 + p, rows] = h1.to(torch.float32)


def _lstm_tc_config(H):
    assert H % 64 == 0
    from idccrn_b200 import lib
    return lib.lstm_tc_config(H)          # pure host function of the real library (no GPU needed)


def idv_lstm_h1_fwd(g, g_ld, wrec, num_layers, NB, T, t_valid, out):
    Tp, Tv = T + 1, _tv(t_valid, T)
    gv = _flat(g).view(NB, Tp, g_ld)[:, 1:1 + Tv, :4].to(D)
    w = wrec.view(num_layers, 12).to(D)
    h = torch.zeros(num_layers, NB, dtype=D)
    c = torch.zeros(num_layers, NB, dtype=D)
    res = torch.zeros(NB, Tv, dtype=D)
    for t in range(Tv):
        x = None
        for l in range(num_layers):
            a = gv[:, t] if l == 0 else x[:, None] * w[l, 0:4][None] + w[l, 8:12][None]
            a = a + h[l][:, None] * w[l, 4:8][None]
            c[l] = torch.sigmoid(a[:, 1]) * c[l] + torch.sigmoid(a[:, 0]) * torch.tanh(a[:, 2])
            h[l] = torch.sigmoid(a[:, 3]) * torch.tanh(c[l])
            x = h[l]
        res[:, t] = x
    out.copy_(res.view(NB, Tv, 1).to(torch.float32))


def idv_lstm_combine_fwd(hseq, NB, T, H, latent, t_valid=0):
    Tp = T + 1
    hs = hseq.view(4, NB, Tp, H)[:, :, 1:1 + _tv(t_valid, T)]
    rr, ir, ri, ii = hs[0], hs[1], hs[2], hs[3]
    latent.copy_(torch.stack((rr - ii, ir + ri), -1))


def idv_reparam_fwd(latent, NB, T, Htot, ch0, zdim, S, eps_r, eps_i, seed, offset, offset_dev, variant, z):
    assert eps_r is not None, "the emulator only supports supplied eps"
    e = 1e-6
    lat = latent.view(NB, T, Htot, 2)
    mu, ls, dl = lat[:, :, ch0:ch0 + zdim], lat[:, :, ch0 + zdim:ch0 + 2 * zdim], lat[:, :, ch0 + 2 * zdim:ch0 + 3 * zdim]
    sig = torch.exp(torch.clamp(ls[..., 0], -13, 13) if variant else ls[..., 0])
    dr, di = dl[..., 0], dl[..., 1]
    ad = torch.sqrt(dr * dr + di * di + e)
    tmp = sig * 0.99 / (ad + e)
    cl = ad >= sig - 1e-3
    dr, di = torch.where(cl, dr * tmp, dr), torch.where(cl, di * tmp, di)
    ad = torch.sqrt(dr * dr + di * di + e)
    er, ei = eps_r.view(NB, S, T, zdim), eps_i.view(NB, S, T, zdim)
    if variant:
        den = torch.sqrt(torch.clamp(2 * (sig + dr), min=e))
        zr = mu[..., 0][:, None] + ((sig + dr) / den)[:, None] * er
        zi = mu[..., 1][:, None] + (di / den)[:, None] * er + (torch.sqrt(torch.clamp(sig * sig - ad * ad, min=e)) / den)[:, None] * ei
    else:
        den = torch.sqrt(2 * (sig + dr) + e)
        zr = mu[..., 0][:, None] + ((sig + dr) / (den + e))[:, None] * er
        zi = mu[..., 1][:, None] + (di / (den + e))[:, None] * er + \
            (torch.sqrt(sig * sig - ad * ad + e) / (den + e))[:, None] * ei
    z.copy_(torch.stack((zr, zi), -1).view(NB * S, T, zdim, 2))


def idv_latent_fwd(hseq, NB, T, H, t_valid, zdim, latent_num, S, eps_r0, eps_i0, eps_r1, eps_i1, seed, offset, offset_dev,
                   latent, z0, z1, zplanes, out_split, keep_pad=0):
    """Contract = idv_lstm_combine_fwd, then idv_reparam_fwd per latent, then idv_z_to_planes per sample; keep_pad: the
    pad rows of zplanes keep their content."""
    Tv = _tv(t_valid, T)
    old = _flat(zplanes).clone() if keep_pad else None
    idv_lstm_combine_fwd(hseq, NB, T, H, latent, t_valid)
    for k, (er, ei, z) in enumerate(((eps_r0, eps_i0, z0), (eps_r1, eps_i1, z1))[:latent_num]):
        idv_reparam_fwd(latent, NB, Tv, H, 3 * zdim * k, zdim, S, er, ei, seed, offset, offset_dev, 0, z)
    per = NB * (T + 1) * 2 * _r8(zdim) * (2 if out_split else 1)
    zp = _flat(zplanes)
    for s in range(S):
        idv_z_to_planes(z0, NB, S, s, T, zdim, zp[s * per:(s + 1) * per], out_split, t_valid)
    if keep_pad:
        # planes [S][hl][NB][T + 1][Cp]: row 0 of every utterance back to what it was
        cp = 2 * _r8(zdim)
        new, prev = zp.view(-1, NB, T + 1, cp), old.view(-1, NB, T + 1, cp)
        new[:, :, 0] = prev[:, :, 0]


def idv_lstm_combine_planes(hseq, NB, T, H, t_valid, latent, planes, out_split):
    idv_lstm_combine_fwd(hseq, NB, T, H, latent, t_valid)
    idv_z_to_planes(latent, NB, 1, 0, T, H, planes, out_split, t_valid)


def idv_bin_affine(x, B, F, T, scale, shift, zero_edge_imag, out):
    v = x.view(B, F, T, 2).to(D) * scale.view(1, F, 1, 2).to(D) + shift.view(1, F, 1, 2).to(D)
    if zero_edge_imag:
        v[:, 0, :, 1] = 0
        v[:, -1, :, 1] = 0
    out.view(B, F, T, 2).copy_(v.to(torch.float32))


def _r8(c):
    return (c + 7) // 8 * 8


def idv_planes_to_user(planes, in_split, NB, C, F, T, user, t_valid=0):
    Ch, Tp = _r8(C), T + 1
    p = _rd(planes, in_split, F * NB * Tp * 2 * Ch).view(F, NB, Tp, 2, Ch)[:, :, 1:1 + _tv(t_valid, T), :, :C]
    user.copy_(p.permute(1, 4, 0, 2, 3).to(torch.float32))


def idv_user_to_planes(user, NB, C, F, T, planes, out_split=0, t_valid=0):
    Ch, Tp, Tv = _r8(C), T + 1, _tv(t_valid, T)
    p = torch.zeros(F, NB, Tp, 2, Ch, dtype=D)
    p[:, :, 1:1 + Tv, :, :C] = user.view(NB, C, F, Tv, 2).permute(2, 0, 3, 4, 1).to(D)
    _wr(planes, out_split, p)


def idv_z_to_planes(z, NB, S, s, T, zdim, planes, out_split=0, t_valid=0):
    Ch, Tp, Tv = _r8(zdim), T + 1, _tv(t_valid, T)
    p = torch.zeros(NB, Tp, 2, Ch, dtype=D)
    p[:, 1:1 + Tv, :, :zdim] = z.view(NB, S, Tv, zdim, 2)[:, s].permute(0, 1, 3, 2).to(D)
    _wr(planes, out_split, p)


def idv_cbn_eval_user(x, outer, C, inner, zb, out):
    v = x.view(outer, C, inner, 2)
    k = zb.view(C, 6)[None, :, None, :]
    o = out.view(outer, C, inner, 2)
    o_r = k[..., 0] * v[..., 0] + k[..., 1] * v[..., 1] + k[..., 4]          # x may alias out: compute both first
    o_i = k[..., 2] * v[..., 0] + k[..., 3] * v[..., 1] + k[..., 5]
    o[..., 0] = o_r
    o[..., 1] = o_i


def idv_cbn_stats_planes(planes, split, NB, C, F, T, acc, t_valid=0):
    Ch, Tp = _r8(C), T + 1
    p = _rd(planes, split, F * NB * Tp * 2 * Ch).view(F, NB, Tp, 2, Ch)[:, :, 1:1 + _tv(t_valid, T), :, :C]
    re, im = p[:, :, :, 0], p[:, :, :, 1]
    acc.view(C, 5).copy_(torch.stack((re.sum((0, 1, 2)), im.sum((0, 1, 2)), (re * re).sum((0, 1, 2)),
                                      (im * im).sum((0, 1, 2)), (re * im).sum((0, 1, 2))), 1))


def idv_cbn_stats_user(x, outer, C, inner, acc):
    v = x.view(outer, C, inner, 2).to(D)
    re, im = v[..., 0], v[..., 1]
    acc.view(C, 5).copy_(torch.stack((re.sum((0, 2)), im.sum((0, 2)), (re * re).sum((0, 2)), (im * im).sum((0, 2)),
                                      (re * im).sum((0, 2))), 1))


def idv_cbn_train_finalize(acc, count, C, g_rr, g_ri, g_ii, beta_r, beta_i, run_mr, run_mi, run_vrr, run_vri, run_vii,
                           momentum, first, zb, stats=None):
    a = acc.view(C, 5).to(D)
    eps = 1e-5
    mr, mi = a[:, 0] / count, a[:, 1] / count
    f32 = lambda t: t.to(torch.float32)
    mu_r, mu_i = f32(mr), f32(mi)
    vrr = f32(a[:, 2] / count - mr * mr) + eps
    vii = f32(a[:, 3] / count - mi * mi) + eps
    vri = f32(a[:, 4] / count - mr * mi)
    if stats is not None:
        stats.view(C, 5).copy_(torch.stack((mu_r, mu_i, vrr, vri, vii), 1))
    for buf, val in ((run_mr, mu_r), (run_mi, mu_i), (run_vrr, vrr), (run_vri, vri), (run_vii, vii)):
        b = buf.view(-1)
        b.copy_(val if first else momentum * b + (1 - momentum) * val)
    delta = torch.clamp(vrr * vii - vri * vri + eps, min=1e-8)
    s = torch.sqrt(delta)
    t = torch.sqrt(vrr + vii + 2 * s + eps)
    ist = 1.0 / (s * t + eps)
    wrr, wii, wri = (vii + s) * ist, (vrr + s) * ist, -vri * ist
    grr, gri, gii = g_rr.view(-1), g_ri.view(-1), g_ii.view(-1)
    zrr, zri = grr * wrr + gri * wri, grr * wri + gri * wii
    zir, zii = gri * wrr + gii * wri, gri * wri + gii * wii
    zb.view(C, 6).copy_(torch.stack((zrr, zri, zir, zii, beta_r.view(-1) - (zrr * mu_r + zri * mu_i),
                                     beta_i.view(-1) - (zir * mu_r + zii * mu_i)), 1))


def idv_cbn_apply_planes(planes, split, NB, C, F, T, zb, apply_prelu, slope, t_valid, out, out_split):
    Ch, Tp = _r8(C), T + 1
    te = 1 + _tv(t_valid, T)
    n = F * NB * Tp * 2 * Ch
    p = _rd(planes, split, n).view(F, NB, Tp, 2, Ch).clone()
    k = zb.view(C, 6).to(D)
    re, im = p[:, :, 1:te, 0, :C].clone(), p[:, :, 1:te, 1, :C].clone()
    o_r = k[:, 0] * re + k[:, 1] * im + k[:, 4]
    o_i = k[:, 2] * re + k[:, 3] * im + k[:, 5]
    if apply_prelu:
        o_r = torch.where(o_r > 0, o_r, slope * o_r)
        o_i = torch.where(o_i > 0, o_i, slope * o_i)
    if out is not None and out is not planes:
        p = torch.zeros_like(p)
    p[:, :, 1:te, 0, :C] = o_r
    p[:, :, 1:te, 1, :C] = o_i
    if out is not None and out is not planes:
        _wr(out, out_split, p)
    else:
        _wr(planes, split, p)


def idv_head_user(y, n_per_utt, n_utt, slope, mask, stft_x, s_rep):
    v = y.view(n_utt, n_per_utt, 2).to(D)
    yr = torch.where(v[..., 0] > 0, v[..., 0], slope * v[..., 0])
    yi = torch.where(v[..., 1] > 0, v[..., 1], slope * v[..., 1])
    if mask:
        mag = torch.tanh(torch.sqrt(yr ** 2 + yi ** 2))
        ph = torch.atan2(yi / (mag + 1e-8), yr / (mag + 1e-8))
        X = stft_x.view(-1, n_per_utt, 2).to(D)[torch.arange(n_utt) // s_rep]
        in_mag = torch.sqrt(X[..., 0] ** 2 + X[..., 1] ** 2)
        in_ph = torch.atan2(X[..., 1], X[..., 0])
        yr, yi = in_mag * mag * torch.cos(in_ph + ph), in_mag * mag * torch.sin(in_ph + ph)
    y.view(n_utt, n_per_utt, 2).copy_(torch.stack((yr, yi), -1).to(torch.float32))


# ---- backward of the encoder (csrc/backward.cu) -----------------------------------------------------------------------
def idv_planes_transpose_split(planes, in_split, F, R, Cp, Rpad, shift, out):
    x = _rd(planes, in_split, F * R * Cp).view(F, R, Cp)
    o = torch.zeros(F, Cp, Rpad, dtype=D)
    for k in range(Rpad):
        r = k + shift
        if 0 <= r < R:
            o[:, :, k] = x[:, r, :]
    _wr(out, 1, o)


def idv_f32_to_split(x, n, out):
    _wr(out, 1, _flat(x)[:n].to(D))


def _cbn_bwd_common(y, y_split, g, g_split, NB, C, F, T, stats, zb, slope, t_valid):
    Ch, Tp, Tv = _r8(C), T + 1, _tv(t_valid, T)
    n = F * NB * Tp * 2 * Ch
    yv = _rd(y, y_split, n).view(F, NB, Tp, 2, Ch)[:, :, 1:1 + Tv, :, :C]
    gv = _rd(g, g_split, n).view(F, NB, Tp, 2, Ch)[:, :, 1:1 + Tv, :, :C]
    k = zb.view(C, 6).to(D)
    st = stats.view(C, 5).to(D)
    yr, yi, gr, gi = yv[..., 0, :], yv[..., 1, :], gv[..., 0, :], gv[..., 1, :]
    pr = k[:, 0] * yr + k[:, 1] * yi + k[:, 4]
    pi = k[:, 2] * yr + k[:, 3] * yi + k[:, 5]
    gpr = torch.where(pr > 0, gr, slope * gr)
    gpi = torch.where(pi > 0, gi, slope * gi)
    return pr, pi, gr, gi, gpr, gpi, yr - st[:, 0], yi - st[:, 1]


def idv_cbn_bwd_reduce(y, y_split, g, g_split, NB, C, F, T, stats, zb, slope, acc, t_valid=0):
    pr, pi, gr, gi, gpr, gpi, xr, xi = _cbn_bwd_common(y, y_split, g, g_split, NB, C, F, T, stats, zb, slope, t_valid)
    sm = lambda v: v.sum((0, 1, 2))
    sl = sm(torch.where(pr > 0, torch.zeros_like(pr), gr * pr)) + sm(torch.where(pi > 0, torch.zeros_like(pi), gi * pi))
    acc.view(C, 8).copy_(torch.stack((sm(gpr), sm(gpi), sm(gpr * xr), sm(gpr * xi), sm(gpi * xr), sm(gpi * xi), sl,
                                      torch.zeros(C, dtype=D)), 1))


def idv_cbn_bwd_finalize(acc, count, C, stats, g_rr, g_ri, g_ii, coef, d_grr, d_gri, d_gii, d_br, d_bi, d_slope):
    """Independent derivation: autograd through cbn()'s per-channel algebra (model/complex_progress.py:L168-205)."""
    s = acc.view(C, 8).to(D)
    st = stats.view(C, 5).to(D)
    eps = 1e-5
    with torch.enable_grad():                     # called from inside an autograd backward (grad mode is off there)
        V = [st[:, 2].clone().requires_grad_(True), st[:, 3].clone().requires_grad_(True), st[:, 4].clone().requires_grad_(True)]
        G = [g_rr.view(-1).to(D).clone().requires_grad_(True), g_ri.view(-1).to(D).clone().requires_grad_(True),
             g_ii.view(-1).to(D).clone().requires_grad_(True)]
        a, cri, b = V
        delta = torch.clamp(a * b - cri * cri + eps, min=1e-8)
        sq = torch.sqrt(delta)
        t = torch.sqrt(a + b + 2 * sq + eps)
        ist = 1.0 / (sq * t + eps)
        wrr, wii, wri = (b + sq) * ist, (a + sq) * ist, -cri * ist
        zrr, zri = G[0] * wrr + G[1] * wri, G[0] * wri + G[1] * wii
        zir, zii = G[1] * wrr + G[2] * wri, G[1] * wri + G[2] * wii
        # L(Z) = <dZ, Z> with dZ_ab = sum gp_a xc_b
        obj = (s[:, 2] * zrr + s[:, 3] * zri + s[:, 4] * zir + s[:, 5] * zii).sum()
        ga, gc, gb, dgrr, dgri, dgii = torch.autograd.grad(obj, V + G)
    for dst, val in ((d_grr, dgrr), (d_gri, dgri), (d_gii, dgii), (d_br, s[:, 0]), (d_bi, s[:, 1])):
        dst.view(-1).add_(val.to(torch.float32))
    n = count
    zrr, zri, zir, zii = zrr.detach(), zri.detach(), zir.detach(), zii.detach()
    coef.view(C, 10).copy_(torch.stack((zrr, zir, zri, zii, 2 * ga / n, gc / n, 2 * gb / n,
                                        (zrr * s[:, 0] + zir * s[:, 1]) / n, (zri * s[:, 0] + zii * s[:, 1]) / n,
                                        torch.zeros(C, dtype=D)), 1).to(torch.float32))
    if d_slope is not None:
        d_slope += s[:, 6].sum()


def idv_cbn_bwd_apply(y, y_split, g, g_split, NB, C, F, T, stats, zb, coef, slope, dy, dy_split, t_valid=0):
    pr, pi, gr, gi, gpr, gpi, xr, xi = _cbn_bwd_common(y, y_split, g, g_split, NB, C, F, T, stats, zb, slope, t_valid)
    k = coef.view(C, 10).to(D)
    Ch, Tp, Tv = _r8(C), T + 1, _tv(t_valid, T)
    o = torch.zeros(F, NB, Tp, 2, Ch, dtype=D)
    o[:, :, 1:1 + Tv, 0, :C] = k[:, 0] * gpr + k[:, 1] * gpi + k[:, 4] * xr + k[:, 5] * xi - k[:, 7]
    o[:, :, 1:1 + Tv, 1, :C] = k[:, 2] * gpr + k[:, 3] * gpi + k[:, 5] * xr + k[:, 6] * xi - k[:, 8]
    _wr(dy, dy_split, o)


def idv_lstm_combine_bwd(dlatent, NB, T, H, dH, t_valid=0):
    Tp, Tv = T + 1, _tv(t_valid, T)
    d = dlatent.view(NB, Tv, H, 2)
    o = torch.zeros(4, NB, Tp, H)
    o[0, :, 1:1 + Tv], o[1, :, 1:1 + Tv], o[2, :, 1:1 + Tv], o[3, :, 1:1 + Tv] = d[..., 0], d[..., 1], d[..., 1], -d[..., 0]
    dH.view(4, NB, Tp, H).copy_(o)


def idv_lstm_scan_c(P, NB, T, H, cst, t_valid=0):
    Tp, Tv = T + 1, _tv(t_valid, T)
    p = P.view(4, NB, Tp, 4 * H).to(D)
    co = cst.view(4, NB, Tp, H)
    c = torch.zeros(4, NB, H, dtype=D)
    co[:, :, 0] = 0
    for t in range(Tv):
        a = p[:, :, 1 + t]
        c = torch.sigmoid(a[..., H:2 * H]) * c + torch.sigmoid(a[..., :H]) * torch.tanh(a[..., 2 * H:3 * H])
        co[:, :, 1 + t] = c.to(torch.float32)


def idv_lstm_cell_bwd_step(P, cst, dH, dh_rec, dc, NB, T, H, t, last, dh_parts, dP, dP_step):
    Tp = T + 1
    a = P.view(4, NB, Tp, 4 * H)[:, :, 1 + t].to(D)
    ig, fg, gg, og = torch.sigmoid(a[..., :H]), torch.sigmoid(a[..., H:2 * H]), torch.tanh(a[..., 2 * H:3 * H]), \
        torch.sigmoid(a[..., 3 * H:])
    cs = cst.view(4, NB, Tp, H).to(D)
    c, cprev = cs[:, :, 1 + t], cs[:, :, t]
    tc = torch.tanh(c)
    dh = dH.view(4, NB, Tp, H)[:, :, 1 + t].to(D)
    dcv = dh * og * (1 - tc * tc)
    if not last:
        dh = dh + _flat(dh_rec)[:dh_parts * 4 * NB * H].view(dh_parts, 4, NB, H).to(D).sum(0)
        dcv = dc.view(4, NB, H).to(D) + dh * og * (1 - tc * tc)
    d = torch.cat((dcv * gg * ig * (1 - ig), dcv * cprev * fg * (1 - fg), dcv * ig * (1 - gg * gg),
                   dh * tc * og * (1 - og)), -1)
    dc.view(4, NB, H).copy_((dcv * fg).to(torch.float32))
    dP.view(4, NB, Tp, 4 * H)[:, :, 1 + t] = d.to(torch.float32)
    _wr(dP_step, 1, d)


def idv_colsum_add(x, rows, cols, ld, out):
    _flat(out)[:cols].add_(_flat(x)[:rows * ld].view(rows, ld)[:, :cols].to(D).sum(0).to(torch.float32))


def idv_enc0_wgrad(stft, dY, B, Fin, T, Cout, causal, dW):
    N = 2 * Cout
    Fout = (Fin + 4 - 5) // 2 + 1
    x = stft.view(B, Fin, T, 2).to(D)
    xpad = torch.zeros(B, Fin + 4, T + 1, 2, dtype=D)
    if causal:
        xpad[:, 2:2 + Fin, 1:] = x
    else:
        xpad[:, 2:2 + Fin, :T] = x
    g = dY.view(Fout, B, T + 1, N)[:, :, 1:].to(D)
    o = torch.zeros(10, 2, N, dtype=D)
    for kf in range(5):
        for kt in range(2):
            sl = xpad[:, kf:kf + 2 * Fout - 1:2, kt:kt + T]                       # (B, Fout, T, 2)
            o[kf * 2 + kt] = torch.einsum("bftp,fbtn->pn", sl, g)
    dW.view(10, 2, N).copy_(o.to(torch.float32))


def idv_kl_fwd_bwd(lat1, H1, ch1, lat2, H2, ch2, n_bt, zdim, scale, mean_scale, dlat1, acc):
    """Contract = the oracle's cal_kl restatement differentiated by autograd (test double only)."""
    from oracle import ref_port as P
    with torch.enable_grad():
        a = lat1.view(1, n_bt, H1, 2).detach().clone().requires_grad_(True)
        b = lat2.view(1, n_bt, H2, 2)
        sl = lambda t, c, k: t[:, :, c + k * zdim:c + (k + 1) * zdim]
        kl = P.cal_kl(sl(a, ch1, 0), sl(b, ch2, 0), sl(a, ch1, 1), sl(b, ch2, 1), sl(a, ch1, 2), sl(b, ch2, 2), zdim).sum()
        (ga,) = torch.autograd.grad(kl, a)
    acc[0] += float(kl) * scale
    acc[1] += float(kl) * mean_scale
    if dlat1 is not None:
        dlat1.view(-1).add_((scale * ga).reshape(-1))


def idv_adam_step(p, g, m, v, n, lr, b1, b2, eps, wd, step):
    pv, gv, mv, vv = (_flat(t)[:n] for t in (p, g, m, v))
    gg = gv + wd * pv
    mv.copy_(b1 * mv + (1 - b1) * gg)
    vv.copy_(b2 * vv + (1 - b2) * gg * gg)
    bc1, bc2 = 1 - b1 ** step, math.sqrt(1 - b2 ** step)
    pv.sub_((lr / bc1) * mv / (vv.sqrt() / bc2 + eps))


# ---- frame streaming --------------------------------------------------------------------------------------------
def idv_stream_frames_split(hist, x_new, NB, k, base, hop, win, kpad, frames):
    hl = win - hop
    w = torch.cat((hist.view(NB, hl), x_new.view(NB, hop * k)), 1).to(D)
    wlen = w.shape[1]
    fr = torch.zeros(NB, k, kpad, dtype=D)
    for f in range(k):
        for j in range(win):
            wi = hop * f + j
            if base + wi < 0:
                wi = -(base + wi) - base
            if 0 <= wi < wlen:
                fr[:, f, j] = w[:, wi]
    _wr(frames, 1, fr)


def idv_stream_hist_shift(hist, x_new, NB, k, hop, win):
    hl = win - hop
    w = torch.cat((hist.view(NB, hl), x_new.view(NB, hop * k)), 1)
    hist.view(NB, hl).copy_(w[:, hop * k:hop * k + hl].clone())


def idv_lstm_cell_step(g_in, g_m_off, g_p_off, g_ld, g_rec, NB, H, T, frame, c, h_split, hseq):
    Tp = T + 1
    a = g_rec.view(4, NB, 4 * H).to(D).clone()
    if g_in is not None:
        gf = _flat(g_in)
        rows = torch.arange(NB) * Tp + 1 + frame
        for s in range(4):
            idx = (s >> 1) * g_m_off + (s & 1) * g_p_off + rows[:, None] * g_ld + torch.arange(4 * H)[None, :]
            a[s] += gf[idx].to(D)
    i, f, gg, o = a[..., :H], a[..., H:2 * H], a[..., 2 * H:3 * H], a[..., 3 * H:]
    cn = torch.sigmoid(f) * c.view(4, NB, H).to(D) + torch.sigmoid(i) * torch.tanh(gg)
    hn = torch.sigmoid(o) * torch.tanh(cn)
    c.view(4, NB, H).copy_(cn.to(torch.float32))
    _wr(h_split, 1, hn)
    if hseq is not None:
        hseq.view(4, NB, Tp, H)[:, :, 1 + frame] = hn.to(torch.float32)


def idv_carry_rows(table, n_entries, counter):
    """table: list of (tensor, n_planes, NB, Tp, src_row) in the emulator (the product passes idv_carry_t records)."""
    for (t, n_planes, NB, Tp, src_row) in table._entries[:n_entries]:
        v = _flat(t).view(n_planes, NB, Tp, -1)
        v[:, :, 0] = v[:, :, src_row].clone()
    if counter is not None:
        counter += 1


def idv_stream_last_frame(stft, NB, F, k, prev):
    prev.view(NB, F, 2).copy_(stft.view(NB, F, k, 2)[:, :, k - 1])


def idv_stream_ola(frames, frame_ld, wsq, acc, NB, k, t0, hop, win, out):
    tl = win - hop
    fr = frames.view(NB, k, frame_ld)[:, :, :win].to(D)
    buf = torch.zeros(NB, hop * k + tl + win, dtype=D)
    buf[:, :tl] = acc.view(NB, tl).to(D)
    for f in range(k):
        buf[:, hop * f:hop * f + win] += fr[:, f]
    env = torch.zeros(hop * k, dtype=D)
    for i in range(hop * k):
        q = hop * t0 + i
        for t in range(max(0, (q - win + hop) // hop if q - win + 1 > 0 else 0), q // hop + 1):
            j = q - hop * t
            if 0 <= j < win:
                env[i] += wsq[j].to(D)
    o = torch.where(env > 0, buf[:, :hop * k] / torch.where(env > 0, env, torch.ones_like(env)), torch.zeros_like(env))
    out.view(NB, hop * k).copy_(o.to(torch.float32))
    acc.view(NB, tl).copy_(buf[:, hop * k:hop * k + tl].to(torch.float32))


def idv_stream_tail(table, n_entries, counter, hist, x_new, stft, F, prev, frames, frame_ld, wsq, acc, t0, out, NB, k,
                    hop, win):
    """Contract = the four state entry points above in one launch."""
    idv_carry_rows(table, n_entries, counter)
    if hist is not None:
        idv_stream_hist_shift(hist, x_new, NB, k, hop, win)
    if prev is not None:
        idv_stream_last_frame(stft, NB, F, k, prev)
    idv_stream_ola(frames, frame_ld, wsq, acc, NB, k, t0, hop, win, out)


# ---- decoder side of the training step (csrc/backward.cu) -------------------------------------------------------------
def idv_sisnr_fwd_bwd(src, est, B, L, scale, d_est, sums, loss):
    """Contract = si_snr of model/nsvae_loss.py:L877-889 (restated) differentiated by autograd."""
    with torch.enable_grad():
        s = src.view(B, L).to(D)
        e = est.view(B, L).to(D).detach().clone().requires_grad_(True)
        eps = 1e-8
        alpha = (e * s).sum(1, keepdim=True) / ((s * s).sum(1, keepdim=True) + eps)
        tgt = alpha * s
        noise = e - tgt
        snr = 10 * torch.log10((tgt ** 2).sum(1) / ((noise ** 2).sum(1) + eps) + eps)
        lo = -snr.mean()
        (g,) = torch.autograd.grad(lo, e)
    loss[0] += float(lo)
    sums.view(B, 3).copy_(torch.stack([(e * s).sum(1), (s * s).sum(1), (e * e).sum(1)], 1).detach())
    if d_est is not None:
        d_est.view(B, L).add_((scale * g).to(torch.float32))


def idv_spec_loss_fwd_bwd(pred, ori, n_bins, w_cpx, w_mag, inv_bt, d_pred, acc):
    """Contract = the two spectral terms of multi_recon_loss (model/nsvae_loss.py:L891-906, restated, the ori-magnitude
    quirk of L899 included) differentiated by autograd."""
    with torch.enable_grad():
        p = pred.reshape(n_bins, 2).to(D).detach().clone().requires_grad_(True)
        o = ori.reshape(n_bins, 2).to(D)
        cpx = (((p[:, 0] - o[:, 0]) ** 2).sum() + ((p[:, 1] - o[:, 1]) ** 2).sum()) * inv_bt
        pm = torch.sqrt(p[:, 0] ** 2 + p[:, 1] ** 2 + 1e-6)
        om = torch.sqrt(o[:, 0] ** 2 + o[:, 0] ** 2 + 1e-6)
        mag = ((pm - om) ** 2).sum() * inv_bt
        (g,) = torch.autograd.grad(w_cpx * cpx + w_mag * mag, p)
    acc[0] += float(cpx)
    acc[1] += float(mag)
    if d_pred is not None:
        d_pred.view(n_bins, 2).add_(g.to(torch.float32))


def idv_ola_bwd(dsig, wsq, B, T, n_fft, hop, win, frame_ld, dframes):
    """Adjoint of idv_ola_fwd (autograd of its restatement)."""
    off = (n_fft - win) // 2
    total = n_fft + hop * (T - 1)
    h = n_fft // 2
    env = torch.zeros(total, dtype=D)
    for t in range(T):
        env[t * hop + off:t * hop + off + win] += wsq.to(D)
    gy = torch.zeros(B, total, dtype=D)
    gy[:, h:total - h] = dsig.view(B, hop * (T - 1)).to(D) / env[h:total - h]
    o = torch.zeros(B, T, frame_ld, dtype=D)
    for t in range(T):
        o[:, t, :win] = gy[:, t * hop + off:t * hop + off + win]
    dframes.view(B, T, frame_ld).copy_(o.to(torch.float32))


def idv_head_bwd(raw, zb, slope, mask, stft_x, drows, drows_ld, dpred, NB, F, T, y_planes, g_planes):
    """Contract: autograd of CBN-affine -> PReLU -> (mask head) on the raw last-layer output."""
    k = zb.view(6).to(D)
    y = raw.view(NB, F, T, 2).to(D)
    gS = drows.view(NB, T, drows_ld)[:, :, :2 * F].reshape(NB, T, F, 2).permute(0, 2, 1, 3).to(D)
    if dpred is not None:
        gS = gS + dpred.view(NB, F, T, 2).to(D)
    pr = k[0] * y[..., 0] + k[1] * y[..., 1] + k[4]
    pi = k[2] * y[..., 0] + k[3] * y[..., 1] + k[5]
    with torch.enable_grad():
        m = torch.stack((torch.where(pr > 0, pr, slope * pr), torch.where(pi > 0, pi, slope * pi)), -1)
        m = m.detach().clone().requires_grad_(True)
        if mask:
            mr, mi = m[..., 0], m[..., 1]
            mag = torch.tanh(torch.sqrt(mr ** 2 + mi ** 2))
            ph = torch.atan2(mi / (mag + 1e-8), mr / (mag + 1e-8))
            X = stft_x.view(NB, F, T, 2).to(D)
            in_mag = torch.sqrt(X[..., 0] ** 2 + X[..., 1] ** 2)
            in_ph = torch.atan2(X[..., 1], X[..., 0])
            S = torch.stack((in_mag * mag * torch.cos(in_ph + ph), in_mag * mag * torch.sin(in_ph + ph)), -1)
            (gm,) = torch.autograd.grad((S * gS).sum(), m)
        else:
            gm = gS
    Tp = T + 1
    yp = torch.zeros(F, NB, Tp, 16)
    gp = torch.zeros(F, NB, Tp, 16)
    yp[:, :, 1:, 0], yp[:, :, 1:, 8] = y[..., 0].permute(1, 0, 2), y[..., 1].permute(1, 0, 2)
    gp[:, :, 1:, 0], gp[:, :, 1:, 8] = gm[..., 0].permute(1, 0, 2), gm[..., 1].permute(1, 0, 2)
    y_planes.view(F, NB, Tp, 16).copy_(yp)
    g_planes.view(F, NB, Tp, 16).copy_(gp)


def _dec5_taps(Fin):
    Fout = 2 * Fin - 1
    for fi in range(Fin):
        for kf in range(5):
            fo = 2 * fi - 2 + kf
            if 0 <= fo < Fout:
                yield fi, kf, fo


def idv_dec5_dgrad(dy, w10, Ktot, k_off, Cp, Fin, NB, T, dx):
    Tp = T + 1
    Fout = 2 * Fin - 1
    g = dy.view(Fout, NB, Tp, 16)[..., [0, 8]].to(D)                      # (Fout, NB, Tp, 2)
    W = w10.view(10, Ktot, 2)[:, k_off:k_off + Cp].to(D)
    o = torch.zeros(Fin, NB, Tp, Cp, dtype=D)
    for fi, kf, fo in _dec5_taps(Fin):
        o[fi, :, 1:] += g[fo, :, 1:] @ W[kf * 2].t()                       # out[t] uses x[t] through kt = 0
        o[fi, :, 1:T] += g[fo, :, 2:] @ W[kf * 2 + 1].t()                  # and x[t-1] through kt = 1
    dx.view(Fin, NB, Tp, Cp).copy_(o.to(torch.float32))


def idv_dec5_wgrad(x, x_split, dy, Ktot, k_off, Cp, Fin, NB, T, dW):
    Tp = T + 1
    Fout = 2 * Fin - 1
    g = dy.view(Fout, NB, Tp, 16)[..., [0, 8]].to(D)
    xv = _rd(x, x_split, Fin * NB * Tp * Cp).view(Fin, NB, Tp, Cp)
    o = torch.zeros(10, Cp, 2, dtype=D)
    for fi, kf, fo in _dec5_taps(Fin):
        o[kf * 2] += torch.einsum("btc,btp->cp", xv[fi, :, 1:], g[fo, :, 1:])
        o[kf * 2 + 1] += torch.einsum("btc,btp->cp", xv[fi, :, 1:T], g[fo, :, 2:])
    dW.view(10, Ktot, 2)[:, k_off:k_off + Cp].add_(o.to(torch.float32))


def idv_axpy(y, x, a, n):
    _flat(y)[:n].add_(a * _flat(x)[:n])


def idv_reparam_bwd(latent, NB, T, Htot, ch0, zdim, eps_r, eps_i, dz, dlatent):
    """Contract: autograd of the idv_reparam_fwd restatement (S = 1)."""
    with torch.enable_grad():
        lat = latent.view(NB, T, Htot, 2).to(D).detach().clone().requires_grad_(True)
        z = torch.zeros(NB, T, zdim, 2, dtype=D)
        e = 1e-6
        mu, ls, dl = lat[:, :, ch0:ch0 + zdim], lat[:, :, ch0 + zdim:ch0 + 2 * zdim], lat[:, :, ch0 + 2 * zdim:ch0 + 3 * zdim]
        sig = torch.exp(ls[..., 0])
        dr, di = dl[..., 0], dl[..., 1]
        ad = torch.sqrt(dr * dr + di * di + e)
        tmp = sig * 0.99 / (ad + e)
        cl = ad >= sig - 1e-3
        dr, di = torch.where(cl, dr * tmp, dr), torch.where(cl, di * tmp, di)
        ad = torch.sqrt(dr * dr + di * di + e)
        den = torch.sqrt(2 * (sig + dr) + e)
        er, ei = eps_r.view(NB, T, zdim).to(D), eps_i.view(NB, T, zdim).to(D)
        zr = mu[..., 0] + ((sig + dr) / (den + e)) * er
        zi = mu[..., 1] + (di / (den + e)) * er + (torch.sqrt(sig * sig - ad * ad + e) / (den + e)) * ei
        z = torch.stack((zr, zi), -1)
        (g,) = torch.autograd.grad((z * dz.view(NB, T, zdim, 2).to(D)).sum(), lat)
    dlatent.view(NB, T, Htot, 2).add_(g.to(torch.float32))



TABLE = {k: v for k, v in globals().items() if k.startswith("idv_")}


def call(name, *args, soft_resource=False):
    TABLE[name](*args)
    return True
