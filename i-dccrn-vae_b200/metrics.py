"""Scoring on the far side of the path: per-utterance SI-SDR of a whole batch on the GPU.

utils/eval_metrics.py:L49-64 (compute_sisdr) is one numpy call per file; here the three dot products of every utterance
of a (zero-padded) batch come from ONE pass of ``idv_sisnr_fwd_bwd`` (fp64 accumulation) and the closed form is applied
to the (B, 3) sums.  Zero padding does not change the sums, so ragged batches score exactly like single files."""
import torch

from . import lib


def si_sdr(x_est, x_ref):
    """SI-SDR in dB per utterance.  x_est, x_ref: (B, L) float32 CUDA tensors (zero-padded to one length).  Follows
    compute_sisdr: eps = float32 machine epsilon in the scaling factor and in the ratio."""
    est = lib.require_f32_cuda(x_est, "estimate")
    ref = lib.require_f32_cuda(x_ref, "reference")
    if est.dim() != 2 or est.shape != ref.shape:
        raise RuntimeError("si_sdr expects estimate and reference of one shape (B, L)")
    B, L = est.shape
    sums = torch.empty(B * 3, dtype=torch.float64, device=est.device)
    loss = torch.zeros(1, dtype=torch.float64, device=est.device)
    lib.call("idv_sisnr_fwd_bwd", ref, est, B, L, 0.0, None, sums, loss)
    dot, rss, ee = sums.view(B, 3).unbind(1)
    eps = float(torch.finfo(torch.float32).eps)
    a = (eps + dot) / (rss + eps)
    sss = a * a * rss
    snn = ee - 2 * a * dot + sss
    return 10 * torch.log10((eps + sss) / (eps + snn))


# ---- (E)STOI on the host -------------------------------------------------------------------------------------------
# utils/eval_metrics.py:L73-122 scores with the third-party package ``pystoi`` (``stoi(x_ref, x_est, fs, extended=True)``),
# which is not in this image: what follows restates the PUBLISHED algorithm (Taal et al. 2011, Jensen & Taal 2016: 10 kHz,
# 256-sample frames with 50 % overlap, 40 dB silent-frame removal, 512-point DFT, 15 one-third-octave bands from 150 Hz,
# 30-frame segments) in numpy.  PARITY UNPINNED: it cannot be checked against pystoi here; the resampler is scipy's
# polyphase default, not pystoi's Octave-compatible filter.  Host-side scoring, not part of the GPU hot path.
def _stoi_frames(x, n_frame, hop):
    import numpy as np
    w = np.hanning(n_frame + 2)[1:-1]
    n = (len(x) - n_frame) // hop + 1
    if n <= 0:
        return np.zeros((0, n_frame)), w
    idx = np.arange(n_frame)[None, :] + hop * np.arange(n)[:, None]
    return x[idx] * w, w


def _remove_silent_frames(x, y, dyn_range=40.0, n_frame=256, hop=128):
    import numpy as np
    xf, w = _stoi_frames(x, n_frame, hop)
    yf, _ = _stoi_frames(y, n_frame, hop)
    eps = np.finfo(float).eps
    en = 20.0 * np.log10(np.linalg.norm(xf, axis=1) + eps)
    keep = (np.max(en) - dyn_range - en) < 0
    xf, yf = xf[keep], yf[keep]
    n = len(xf)
    out_x, out_y = np.zeros((n - 1) * hop + n_frame), np.zeros((n - 1) * hop + n_frame)
    for i in range(n):                                   # overlap-add of the kept (windowed) frames
        out_x[i * hop:i * hop + n_frame] += xf[i]
        out_y[i * hop:i * hop + n_frame] += yf[i]
    return out_x, out_y


def _third_octave_matrix(fs, nfft, n_bands, min_freq):
    import numpy as np
    f = np.linspace(0, fs, nfft + 1)[:nfft // 2 + 1]
    k = np.arange(n_bands, dtype=float)
    lo, hi = min_freq * 2.0 ** ((2 * k - 1) / 6), min_freq * 2.0 ** ((2 * k + 1) / 6)
    obm = np.zeros((n_bands, len(f)))
    for i in range(n_bands):
        obm[i, int(np.argmin((f - lo[i]) ** 2)):int(np.argmin((f - hi[i]) ** 2))] = 1.0
    return obm


def stoi(x_ref, x_est, fs, extended=True):
    """(Extended) short-time objective intelligibility of one utterance, 1-D float arrays (see the note above)."""
    import numpy as np
    from .wavio import resample
    FS, N_FRAME, NFFT, NUMBAND, MINFREQ, N, BETA = 10000, 256, 512, 15, 150, 30, -15.0
    x, y = np.asarray(x_ref, dtype=np.float64), np.asarray(x_est, dtype=np.float64)
    if x.shape != y.shape or x.ndim != 1:
        raise ValueError("stoi expects two 1-D signals of one length")
    if fs != FS:
        x, y = resample(x, fs, FS).astype(np.float64), resample(y, fs, FS).astype(np.float64)
    x, y = _remove_silent_frames(x, y, 40.0, N_FRAME, N_FRAME // 2)
    xs = np.fft.rfft(_stoi_frames(x, N_FRAME, N_FRAME // 2)[0], NFFT).T            # (bins, frames)
    ys = np.fft.rfft(_stoi_frames(y, N_FRAME, N_FRAME // 2)[0], NFFT).T
    if xs.shape[1] < N:
        raise ValueError("not enough non-silent frames for one %d-frame segment" % N)
    obm = _third_octave_matrix(FS, NFFT, NUMBAND, MINFREQ)
    xt, yt = np.sqrt(obm @ np.abs(xs) ** 2), np.sqrt(obm @ np.abs(ys) ** 2)
    seg = lambda t: np.stack([t[:, m - N:m] for m in range(N, t.shape[1] + 1)])   # (segments, bands, N)
    xg, yg = seg(xt), seg(yt)
    eps = np.finfo(float).eps
    if extended:
        def norm(a):
            a = a - a.mean(axis=2, keepdims=True)
            a = a / (np.linalg.norm(a, axis=2, keepdims=True) + eps)
            a = a - a.mean(axis=1, keepdims=True)
            return a / (np.linalg.norm(a, axis=1, keepdims=True) + eps)
        return float(np.sum(norm(xg) * norm(yg) / N) / xg.shape[0])
    alpha = np.linalg.norm(xg, axis=2, keepdims=True) / (np.linalg.norm(yg, axis=2, keepdims=True) + eps)
    yp = np.minimum(yg * alpha, xg * (1 + 10 ** (-BETA / 20)))
    xn, yn = xg - xg.mean(axis=2, keepdims=True), yp - yp.mean(axis=2, keepdims=True)
    xn = xn / (np.linalg.norm(xn, axis=2, keepdims=True) + eps)
    yn = yn / (np.linalg.norm(yn, axis=2, keepdims=True) + eps)
    return float(np.sum(xn * yn) / (xg.shape[0] * NUMBAND))
