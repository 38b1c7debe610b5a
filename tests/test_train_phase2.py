"""Phase-2 decoder training step (train_second_phase_decoder.py:L376-433): frozen NSVAE encoder, decoder train=True with
real skips (pad='sig'), two_phase_loss.multi_recon_loss, backward through iSTFT / reconstruction head /
ComplexBatchNormal(train) + PReLU / complex transposed convs / ComplexDense.  Gradients are pinned by the fixtures that
oracle/make_golden.py wrote from the REAL reference's autograd.  CPU tier: host logic over the emulated C-ABI contract;
GPU tier: the CUDA kernels."""
import numpy as np
import pytest
import torch

import common as C
import idccrn_b200 as M
from idccrn_b200.synth import synth_eps, synth_waveform
from oracle import ref_port as P
from test_train_step import grad_probe


def run_phase2(golden, tag, device, fused_sisnr=False):
    g = golden(tag)
    B, L, latent_num, seed = int(g["B"]), int(g["L"]), int(g["latent_num"]), int(g["seed"])
    recon_type = "mask" if int(g["mask"]) else "real_imag"
    weights = [float(w) for w in g["weights"]]
    S = int(g["S"]) if "S" in g else 1            # num_samples: rows b*S + s of one batch (train_second_phase_decoder.sh:L6)
    enc, dec = C.build_vae(latent_num, S, "twophase", recon_type, seed, device)
    xs = [synth_waveform(B, L, seed=1234 + seed + j).to(device) for j in range(2)]
    T = L // C.HOP + 1
    eps = [e.to(device) for e in synth_eps((B, S, T, C.ZDIM), seed=7 + seed, n=2 * latent_num)]
    with torch.no_grad():
        r = enc(xs[0], train=False, eps=eps)
        stft_clean = enc.stft(xs[1]).repeat_interleave(S, 0)
    xs[1] = xs[1].repeat_interleave(S, 0)
    sig, pred = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
    assert sig.grad_fn is not None and pred.grad_fn is not None and sig.shape[0] == B * S
    if fused_sisnr:                     # the product's fused loss kernels (idv_spec_loss_fwd_bwd + idv_sisnr_fwd_bwd)
        from idccrn_b200 import losses
        loss, l_cpx, l_mag, l_si = losses.multi_recon_loss(pred, stft_clean, xs[1], sig, weights)
        for k, v in (("loss_cpx", l_cpx), ("loss_mag", l_mag), ("loss_sisnr", l_si)):
            assert abs(float(v) - float(g[k])) <= 2e-5 * max(1.0, abs(float(g[k]))), (k, float(v), float(g[k]))
    else:
        loss, l_cpx, l_mag, l_si = P.multi_recon_loss(pred, stft_clean, xs[1], sig, weights)   # the reference's formula
    loss.backward()
    return g, dec, loss, sig


def check_phase2(g, dec, loss, sig, tol):
    errs = {"loss": abs(float(loss) - float(g["loss"])) / max(1.0, abs(float(g["loss"]))),
            "recon_sig": C.rel_l2(sig, g["recon_sig"])}
    # a PReLU slope gradient is ONE scalar = a cancelling sum over the layer: judge it on the scale of the largest one
    slope_scale = max(float(v) for k, v in g.items() if k.startswith("norm/") and k.endswith("prelu.weight"))
    for name, p in dec.named_parameters():
        if "zero/" + name in g:
            assert p.grad is not None and float(p.grad.abs().max()) < 1e-4, name
            continue
        assert p.grad is not None, name
        gd = p.grad.detach().cpu().double()
        norm, probe = float(g["norm/" + name]), float(g["probe/" + name])
        if name.endswith("prelu.weight"):
            errs["full/" + name] = abs(float(gd) - float(g["full/" + name])) / slope_scale
            continue
        errs["norm/" + name] = abs(float(gd.norm()) - norm) / norm
        errs["probe/" + name] = abs(float((gd * grad_probe(name, p.shape)).sum()) - probe) / norm
        if "full/" + name in g:
            errs["full/" + name] = C.rel_l2(gd, g["full/" + name])

    # PReLU flips (see test_train_step.py): parameters below a PReLU seen from the loss carry ~sqrt(flip fraction)
    # relative noise; the last layer's ComplexBatchNormal / PReLU and everything under it sit below the head's PReLU
    def bound(k):
        return tol if k in ("loss", "recon_sig") else 3e-2
    bad = {k: v for k, v in errs.items() if not v < bound(k)}
    print("phase2 worst", max(errs.items(), key=lambda kv: kv[1]))
    assert not bad, bad


@pytest.mark.parametrize("tag", ["train_phase2_mask_sisnr", "train_phase2_ri_multi"])
def test_port_autograd_matches_reference_phase2(golden, tag):
    """The oracle differentiated by autograd against the reference's decoder gradients."""
    g = golden(tag)
    B, L, latent_num, seed = int(g["B"]), int(g["L"]), int(g["latent_num"]), int(g["seed"])
    recon_type = "mask" if int(g["mask"]) else "real_imag"
    enc, dec = C.build_vae(latent_num, 1, "twophase", recon_type, seed, "cpu")
    params = dict(dec.named_parameters())
    sd = {k: v.detach().clone().requires_grad_(k in params) for k, v in dec.state_dict().items()}
    xs = [synth_waveform(B, L, seed=1234 + seed + j) for j in range(2)]
    T = L // C.HOP + 1
    eps = synth_eps((B, 1, T, C.ZDIM), seed=7 + seed, n=2 * latent_num)
    with torch.no_grad():
        st = P.vae_encoder_forward(enc.state_dict(), xs[0], C.ZDIM, latent_num, 1, eps)
    dd = P.vae_decoder_forward(sd, st["stft_x"], st["z_speech"], st["skiper"], st["C"], st["F"], 1, recon_type, "sig",
                               train=True)
    loss, _, _, l_si = P.multi_recon_loss(dd["predict"], P.stft(xs[1]), xs[1], dd["recon_sig"], [float(w) for w in g["weights"]])
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * max(1.0, abs(float(g["loss"])))
    assert abs(float(l_si) - float(g["loss_sisnr"])) <= 1e-5 * max(1.0, abs(float(g["loss_sisnr"])))
    for name in sd:
        if "norm/" + name in g:
            gd = sd[name].grad.double()
            assert abs(float(gd.norm()) - float(g["norm/" + name])) / float(g["norm/" + name]) < 5e-4, name
            if "full/" + name in g:
                assert P.rel_l2(gd, g["full/" + name]) < 5e-4, name


def _tc_mode():
    from idccrn_b200 import ops
    old = ops.GEMM_MODE[0]
    ops.set_gemm_mode("tc")
    return old


@pytest.mark.parametrize("tag", ["train_phase2_mask_sisnr", "train_phase2_ri_multi", "train_phase2_mask_sisnr_s2"])
def test_phase2_gradients_emulated(emulated_abi, golden, tag):
    from idccrn_b200 import ops
    old = _tc_mode()
    try:
        check_phase2(*run_phase2(golden, tag, "cpu"), tol=2e-4)
    finally:
        ops.set_gemm_mode(old)


@pytest.mark.parametrize("tag", ["train_phase2_mask_sisnr", "train_phase2_ri_multi"])
def test_phase2_fused_loss_emulated(emulated_abi, golden, tag):
    from idccrn_b200 import ops
    old = _tc_mode()
    try:
        check_phase2(*run_phase2(golden, tag, "cpu", fused_sisnr=True), tol=2e-4)
    finally:
        ops.set_gemm_mode(old)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["train_phase2_mask_sisnr", "train_phase2_ri_multi", "train_phase2_mask_sisnr_s2"])
def test_phase2_gradients_gpu(golden, tag):
    check_phase2(*run_phase2(golden, tag, "cuda"), tol=5e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["train_phase2_mask_sisnr", "train_phase2_ri_multi", "train_phase2_mask_sisnr_s2"])
def test_phase2_fused_loss_gpu(golden, tag):
    check_phase2(*run_phase2(golden, tag, "cuda", fused_sisnr=True), tol=5e-4)
