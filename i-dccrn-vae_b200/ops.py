"""Thin Python front of the C ABI: shape bookkeeping + output allocation (torch is used only for device
memory and streams).  Every function ends in a ``lib.call`` – there is no other compute path."""
import torch

from . import lib
from .pack import round8


class Planes:
    """Internal activation: fp32 [F][R][Cp], R = NB*(T+1) (row b*(T+1) is the causal zero row),
    Cp = 2*round8(C) with the real parts in [0, Ch) and the imaginary parts in [Ch, 2Ch)."""
    __slots__ = ("data", "NB", "C", "F", "T", "_cp")

    def __init__(self, data, NB, C, F, T, cp=None):
        self.data, self.NB, self.C, self.F, self.T, self._cp = data, NB, C, F, T, cp

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


def tapgemm(pack, a0, a1, NB, T, zero_pad_rows=True):
    """Run one packed tap-GEMM.  a0/a1: Planes (a1 may be None).  Returns the flat output tensor
    [pack.out_planes][R][pack.out_ld]."""
    R = NB * (T + 1)
    out = _empty(pack.out_planes * R * pack.out_ld, a0.data.device)
    lib.call("idv_tapgemm_f32",
             a0.data, a0.Cp, a0.plane_stride,
             a1.data if a1 is not None else None, a1.Cp if a1 is not None else 0,
             a1.plane_stride if a1 is not None else 0,
             R, (T + 1) if zero_pad_rows else 0,
             pack.w, pack.bias, pack.N, pack.units, pack.taps, pack.n_units,
             out, pack.out_ld, R * pack.out_ld, 1 if pack.prelu else 0, pack.slope)
    return out


def enc0(stft_x, w, bias, cout, slope):
    B, Fin, T, _ = stft_x.shape
    Fout = (Fin + 4 - 5) // 2 + 1
    out = _empty(Fout * B * (T + 1) * 2 * cout, stft_x.device)
    lib.call("idv_enc0_fwd", stft_x, B, Fin, T, w, bias, cout, slope, out)
    return Planes(out, B, cout, Fout, T)


def dec5_head(p, skip, w, bias, slope, mask, stft_x, predict, out_bmul, out_boff):
    lib.call("idv_dec5_head_fwd", p.data, p.Cp, skip.data if skip is not None else None,
             skip.Cp if skip is not None else 0, p.NB, p.F, p.T, w, bias, slope, 1 if mask else 0,
             stft_x if mask else None, predict, out_bmul, out_boff)


def lstm_recurrent(g, g_m_off, g_p_off, g_ld, whh, NB, T, H):
    hseq = _empty(4 * NB * (T + 1) * H, g.device)
    sync = torch.empty(2, dtype=torch.int32, device=g.device)
    lib.call("idv_lstm_recurrent_fwd", g, g_m_off, g_p_off, g_ld, whh, NB, T, H, hseq, sync)
    return hseq


def lstm_combine(hseq, NB, T, H):
    latent = torch.empty((NB, T, H, 2), dtype=torch.float32, device=hseq.device)
    lib.call("idv_lstm_combine_fwd", hseq, NB, T, H, latent)
    return latent


def reparam(latent, ch0, zdim, S, eps_r, eps_i, seed, offset):
    NB, T, Htot, _ = latent.shape
    z = torch.empty((NB * S, T, zdim, 2), dtype=torch.float32, device=latent.device)
    if eps_r is not None:
        eps_r = lib.require_f32_cuda(eps_r, "eps_r")
        eps_i = lib.require_f32_cuda(eps_i, "eps_i")
        if tuple(eps_r.shape) != (NB, S, T, zdim) or tuple(eps_i.shape) != (NB, S, T, zdim):
            raise RuntimeError("eps must have shape (B, S, T, zdim) = %s" % ((NB, S, T, zdim),))
    lib.call("idv_reparam_fwd", latent, NB, T, Htot, ch0, zdim, S, eps_r, eps_i, int(seed), int(offset), z)
    return z


def planes_to_user(p):
    out = torch.empty((p.NB, p.C, p.F, p.T, 2), dtype=torch.float32, device=p.data.device)
    lib.call("idv_planes_to_user", p.data, p.NB, p.C, p.F, p.T, out)
    return out


def user_to_planes(x):
    x = lib.require_f32_cuda(x, "activation")
    if x.dim() != 5 or x.shape[-1] != 2:
        raise RuntimeError("activation must be (B, C, F, T, 2), got %s" % (tuple(x.shape),))
    NB, C, F, T, _ = x.shape
    data = _empty(F * NB * (T + 1) * 2 * round8(C), x.device)
    lib.call("idv_user_to_planes", x, NB, C, F, T, data)
    return Planes(data, NB, C, F, T)


def z_to_planes(z, NB, S, s):
    z = lib.require_f32_cuda(z, "z")
    _, T, zdim, _ = z.shape
    data = _empty(NB * (T + 1) * 2 * round8(zdim), z.device)
    lib.call("idv_z_to_planes", z, NB, S, s, T, zdim, data)
    return Planes(data, NB, zdim, 1, T)


def cbn_eval_user(x, zb):
    x = lib.require_f32_cuda(x, "activation")
    B, C = x.shape[0], x.shape[1]
    inner = x[0, 0].numel() // 2
    out = torch.empty_like(x)
    lib.call("idv_cbn_eval_user", x, B, C, inner, zb, out)
    return out
