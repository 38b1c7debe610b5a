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
