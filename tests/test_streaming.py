"""Frame streaming with carried state (BASELINE config 5) must reproduce the whole-utterance causal forward in the
interior.  CPU tier: host logic over the emulated C-ABI contract; GPU tier: the CUDA kernels, eager
steps and CUDA-graph replays."""
import pytest
import torch

import common as C
from idccrn_b200.streaming import StreamingEnhancer


def _tail_cut(se):
    """Samples at the end that a stream cannot reproduce: the whole-utterance forward reflect-pads the END of the
    signal (frames reaching past the last sample), a stream sees real future samples / zeros instead."""
    return se.n_fft + se.hop


def run_stream_case(device, dec_kind, recon, latent_num, k, B=2, L=2300, seed=12, graph=False, tol=2e-5):
    enc, dec = C.build_vae(latent_num, 1, dec_kind, recon, seed, device)
    x, eps = C.vae_inputs(B, L, 1, latent_num, seed, device)
    whole = C.run_vae(enc, dec, x, eps, dec_kind)
    se = StreamingEnhancer(enc, dec, n_streams=B, frames_per_step=k, pad="sig", device=device, use_graph=graph)
    # feed the signal followed by silence so that every frame of the utterance is flushed
    xz = torch.cat((x, torch.zeros(B, se.hop * (k + 6), device=x.device)), 1)
    T_all = (xz.shape[1] - se.hop) // se.hop
    e = [torch.cat((t, torch.zeros(B, 1, T_all, C.ZDIM, device=x.device)), 2) for t in eps[:2]]
    y = se.enhance(xz, eps=e)
    n = L - _tail_cut(se)
    ref = whole["recon_sig"][:, :n]
    err = C.rel_l2(y[:, :n], ref)
    assert y.shape[1] >= n and err < tol, (err, y.shape, n)
    return se, err


@pytest.mark.parametrize("dec_kind,recon,latent_num,k", [
    ("skip_prepare", "real_imag", 1, 1),          # config 2 pairing, one frame per step
    ("twophase", "mask", 2, 3),                   # final system (real skips, mask head), 3 frames per step
])
def test_stream_equals_whole_utterance_emulated(emulated_abi, dec_kind, recon, latent_num, k):
    from idccrn_b200 import ops
    old = ops.GEMM_MODE[0]
    ops.set_gemm_mode("tc")
    try:
        run_stream_case("cpu", dec_kind, recon, latent_num, k, L=1500)
    finally:
        ops.set_gemm_mode(old)


@pytest.mark.gpu
@pytest.mark.parametrize("dec_kind,recon,latent_num,k,graph", [
    ("skip_prepare", "real_imag", 1, 1, False),
    ("skip_prepare", "real_imag", 1, 1, True),
    ("twophase", "mask", 2, 2, True),
    ("twophase", "real_imag", 1, 4, True),
])
def test_stream_equals_whole_utterance_gpu(dec_kind, recon, latent_num, k, graph):
    se, err = run_stream_case("cuda", dec_kind, recon, latent_num, k, B=3, L=6400, graph=graph, tol=5e-5)
    assert (se._graph is not None) == graph
    print("stream", dec_kind, recon, k, graph, "rel_l2 %.2e" % err)


@pytest.mark.gpu
def test_stream_reset_and_philox_noise():
    """reset() restarts the streams (same input -> same output with supplied eps); without eps the on-device Philox
    counter advances on every graph replay, so two passes over the same input differ."""
    enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", 5, "cuda")
    se = StreamingEnhancer(enc, dec, n_streams=4, frames_per_step=1)
    x = C.synth_waveform(4, 3000, seed=3).cuda()
    y1, y2 = se.enhance(x), se.enhance(x)
    assert torch.isfinite(y1).all() and C.rel_l2(y1, y2) > 1e-3
    e = [torch.zeros(4, 1, 29, C.ZDIM, device="cuda")] * 2
    z1, z2 = se.enhance(x, eps=e), se.enhance(x, eps=e)
    assert C.rel_l2(z1, z2) < 1e-6
