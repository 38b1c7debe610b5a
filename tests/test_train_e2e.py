"""End-to-end training step: noisy NSVAE encoder train=True + twophase decoder train=True, loss =
closed-form KL to the frozen clean / noise posteriors + SI-SNR of the reconstruction
(nsvae_loss_with_cvae_decoder_recon.kl_loss_and_recon_loss, model/nsvae_loss.py:L598-613, recon weights (0, 0, 1)).
The decoder's loss reaches the encoder through z (reparameterisation backward) and through the skip tensors.  Gradients
of BOTH models are pinned by fixtures written from the REAL reference's autograd (oracle/make_golden.py --only-e2e)."""
import pytest
import torch

import common as C
import idccrn_b200 as M
from idccrn_b200 import losses
from idccrn_b200.synth import fill_state_dict, synth_eps, synth_waveform
from test_train_step import grad_probe


def build(latent_num, seed, device):
    net = M.get_net_params()
    noisy = M.nsvae_pvae_dccrn_encoder_twophase(net, True, device, C.ZDIM, C.NFFT, C.HOP, C.WIN, 1, latent_num)
    noisy.load_state_dict(fill_state_dict(noisy.state_dict(), seed), strict=True)
    dec = M.nsvae_pvae_dccrn_decoder_twophase(net, True, device, 1, C.ZDIM, C.NFFT, C.HOP, C.WIN, "mask", True, C.SKIPS, False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed + 3), strict=True)
    frozen = []
    for j in range(2):
        e = M.pvae_dccrn_encoder_skip_prepare(net, True, device, C.ZDIM, C.NFFT, C.HOP, C.WIN, 1)
        e.load_state_dict(fill_state_dict(e.state_dict(), seed + 1 + j), strict=True)
        frozen.append(e.to(device).eval())
    return noisy.to(device), dec.to(device), frozen


def run_e2e(golden, tag, device, tol):
    g = golden(tag)
    B, L, latent_num, seed = int(g["B"]), int(g["L"]), int(g["latent_num"]), int(g["seed"])
    noisy, dec, frozen = build(latent_num, seed, device)
    xs = [synth_waveform(B, L, seed=1234 + seed + j).to(device) for j in range(3)]
    T = L // C.HOP + 1
    dev_eps = lambda s, n: [e.to(device) for e in synth_eps((B, 1, T, C.ZDIM), seed=s, n=n)]
    with torch.no_grad():
        rc = frozen[0](xs[1], train=False, eps=dev_eps(8 + seed, 2))
        rn = frozen[1](xs[2], train=False, eps=dev_eps(9 + seed, 2))
    r = noisy(xs[0], train=True, eps=dev_eps(7 + seed, 2 * latent_num))
    sig, pred = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
    kl, _, _ = losses.nsvae_kl_loss(r, rc, rn, C.ZDIM, latent_num, 1.0)
    sisnr = losses.si_snr_loss(xs[1], sig)
    loss = kl + sisnr
    loss.backward()
    errs = {"kl": abs(float(kl) - float(g["kl"])) / max(1.0, abs(float(g["kl"]))),
            "sisnr": abs(float(sisnr) - float(g["sisnr"])) / max(1.0, abs(float(g["sisnr"]))),
            "recon_sig": C.rel_l2(sig, g["recon_sig"])}
    slope_scale = max(float(v) for k, v in g.items() if k.startswith("norm/") and k.endswith("prelu.weight"))
    for pre, mod in (("enc/", noisy), ("dec/", dec)):
        for name, p in mod.named_parameters():
            key = pre + name
            if "zero/" + key in g:
                assert p.grad is not None and float(p.grad.abs().max()) < 1e-3, key
                continue
            if "norm/" + key not in g:
                assert name.startswith("dense.") and p.grad is None, key
                continue
            assert p.grad is not None, key
            gd = p.grad.detach().cpu().double()
            norm, probe = float(g["norm/" + key]), float(g["probe/" + key])
            if name.endswith("prelu.weight"):
                errs["full/" + key] = abs(float(gd.reshape(-1)[0]) - float(g["full/" + key].reshape(-1)[0])) / slope_scale
                continue
            errs["norm/" + key] = abs(float(gd.norm()) - norm) / norm
            errs["probe/" + key] = abs(float((gd * grad_probe(name, p.shape)).sum()) - probe) / norm
            if "full/" + key in g:
                errs["full/" + key] = C.rel_l2(gd, g["full/" + key])
    # everything sits below at least the reconstruction head's PReLU (branch flips: see test_train_step.py)
    bound = lambda k: tol if k in ("kl", "sisnr", "recon_sig") else 3e-2
    bad = {k: v for k, v in errs.items() if not v < bound(k)}
    print(tag, device, "worst", sorted(errs.items(), key=lambda kv: -kv[1])[:3])
    assert not bad, bad


@pytest.mark.parametrize("tag", ["train_e2e_l2", "train_e2e_l1"])
def test_e2e_gradients_emulated(emulated_abi, golden, tag):
    from idccrn_b200 import ops
    old = ops.GEMM_MODE[0]
    ops.set_gemm_mode("tc")
    try:
        run_e2e(golden, tag, "cpu", 2e-4)
    finally:
        ops.set_gemm_mode(old)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["train_e2e_l2", "train_e2e_l1"])
def test_e2e_gradients_gpu(golden, tag):
    run_e2e(golden, tag, "cuda", 5e-4)
