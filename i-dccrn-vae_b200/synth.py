"""Deterministic synthetic weights and inputs (there are no checkpoints or datasets offline).

``fill_state_dict`` assigns every entry of a reference-layout ``state_dict`` from a CPU generator
seeded by ``crc32(key) ^ seed`` – values depend only on (key, shape, seed), never on module
construction order, so the reference modules (when generating goldens), the oracle port and the
B200 modules all get bit-identical weights on any machine.

Value ranges follow the reference's default initialisers in scale (nn.Conv2d / nn.LSTM / nn.Linear
uniform(+-1/sqrt(fan)), ``gamma_ri ~ randn``: model/complex_progress.py:L96-112) but the CBN
running statistics are made non-trivial (Vri != 0, mean != 0) so a wrong fold is visible.
"""
import math
import zlib

import torch


def _gen(key, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def _uniform(shape, lo, hi, g):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def synth_tensor(key, shape, seed=0):
    g = _gen(key, seed)
    leaf = key.rsplit(".", 1)[-1]
    shape = tuple(shape)
    if ".conv_re." in key or ".conv_im." in key or ".tconv_re." in key or ".tconv_im." in key or \
            key.startswith(("conv_re.", "conv_im.", "tconv_re.", "tconv_im.")):
        # bound from the weight's fan-in; bias uses a fixed small bound (shape alone cannot tell)
        if leaf == "weight":
            fan = shape[1] * shape[2] * shape[3]
            b = 1.0 / math.sqrt(fan)
            return _uniform(shape, -b, b, g)
        return _uniform(shape, -0.05, 0.05, g)
    if "lstm_re." in key or "lstm_im." in key:
        hidden = shape[0] // 4
        b = 1.0 / math.sqrt(hidden)
        return _uniform(shape, -b, b, g)
    if "linear_read." in key or "linear_imag." in key:
        b = 1.0 / math.sqrt(128.0)
        return _uniform(shape, -b, b, g)
    if leaf in ("gamma_rr", "gamma_ii"):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf == "gamma_ri":
        return 0.5 * torch.randn(shape, generator=g)
    if leaf in ("beta_r", "beta_i"):
        return 0.1 * torch.randn(shape, generator=g)
    if leaf in ("running_mean_real", "running_mean_imag"):
        return 0.1 * torch.randn(shape, generator=g)
    if leaf in ("Vrr", "Vii"):
        return _uniform(shape, 0.5, 1.5, g)
    if leaf == "Vri":
        return _uniform(shape, -0.3, 0.3, g)
    if ".prelu." in key or key.startswith("prelu."):
        return _uniform(shape, 0.1, 0.4, g)
    if leaf in ("data_mean", "data_std"):
        return _uniform(shape, 0.5, 1.5, g)
    return 0.05 * torch.randn(shape, generator=g)


def fill_state_dict(sd, seed=0):
    """Return a new dict with the same keys/shapes as ``sd`` and synthetic fp32 values."""
    out = {}
    for k, v in sd.items():
        out[k] = synth_tensor(k, v.shape, seed).to(torch.float32)
    return out


def synth_waveform(batch, length, seed=1234, rank=0):
    """0.1*randn(B, L) fp32 – the synthetic utterances of SURVEY §8(d)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + rank)
    return 0.1 * torch.randn(batch, length, generator=g, dtype=torch.float32)


def synth_eps(shape, seed=7, n=2):
    """n standard-normal tensors (eps_real, eps_imag, ...) for a supplied-eps reparameterisation."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return [torch.randn(shape, generator=g, dtype=torch.float32) for _ in range(n)]
