"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference,
build container only) and, in the same run, assert that oracle/ref_port.py reproduces it.

    python oracle/make_golden.py            # writes tests/golden/, prints the port-vs-reference table

Weights come from ``idccrn_b200.synth.fill_state_dict`` (pure function of key/shape/seed), inputs
from ``synth_waveform``, eps from ``synth_eps`` — so the fixtures hold only activations/outputs and
every consumer regenerates weights/inputs bit-identically.  The reference has no eps hook
(model/pvae_module.py:L2219-2220) so ``torch.randn_like`` is patched while it runs.
"""
import contextlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("IDV_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import idccrn_b200  # noqa: E402
from idccrn_b200.synth import fill_state_dict, synth_eps, synth_waveform  # noqa: E402
from oracle import ref_port as P  # noqa: E402

import model.causal_netconfig as ref_causal_cfg  # noqa: E402  (reference)
import model.net_config as ref_noncausal_cfg  # noqa: E402  (reference)
import model.pvae_module as ref_mod  # noqa: E402  (reference)

OUT = os.path.join(ROOT, "tests", "golden")
NFFT, HOP, WIN, ZDIM = 512, 100, 400, 128


@contextlib.contextmanager
def supplied_eps(eps_list):
    """Make the reference's randn_like calls return the supplied tensors, in draw order."""
    it = iter(eps_list)
    orig = torch.randn_like

    def fake(t, *a, **k):
        e = next(it)
        assert e.shape == t.shape, (e.shape, t.shape)
        return e.to(t.dtype)
    torch.randn_like = fake
    try:
        yield
    finally:
        torch.randn_like = orig


def np32(t):
    t = t.detach()
    if t.is_complex():
        t = torch.view_as_real(t)
    return t.to(torch.float32).numpy()


def check(name, a, b, tol=2e-6):
    e = P.rel_l2(a, b)
    print("  port-vs-reference %-28s rel_l2 = %.2e" % (name, e))
    assert e < tol, (name, e)
    return e


def build_vae(latent_num, S, seed, causal=True):
    net = (ref_causal_cfg if causal else ref_noncausal_cfg).get_net_params()
    enc = ref_mod.nsvae_pvae_dccrn_encoder_twophase(net, causal, "cpu", ZDIM, NFFT, HOP, WIN, S, latent_num)
    enc.load_state_dict(fill_state_dict(enc.state_dict(), seed), strict=True)
    return net, enc.eval()


def case_vae(tag, B, L, latent_num, S, dec_kind, recon_type, seed, full, causal=True):
    """causal=False: model/net_config.py (T-1 frames per encoder layer, T+1 per decoder layer: SURVEY §9 V10)."""
    print("case", tag)
    torch.manual_seed(0)
    net, enc = build_vae(latent_num, S, seed, causal)
    if dec_kind == "skip_prepare":
        dec = ref_mod.pvae_dccrn_decoder_skip_prepare(net, causal, "cpu", S, ZDIM, NFFT, HOP, WIN, recon_type,
                                                      [0, 1, 2, 3, 4, 5])
        skip_mode = "zero"
    else:
        dec = ref_mod.nsvae_pvae_dccrn_decoder_twophase(net, causal, "cpu", S, ZDIM, NFFT, HOP, WIN, recon_type,
                                                        True, [0, 1, 2, 3, 4, 5], False)
        skip_mode = "sig"
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed + 1), strict=True)
    dec.eval()
    x = synth_waveform(B, L, seed=1234 + seed)
    T = L // HOP + 1 - (0 if causal else 6)        # latent frames
    eps = synth_eps((B, S, T, ZDIM), seed=7 + seed, n=2 * latent_num)
    with torch.no_grad(), supplied_eps(eps):
        r = enc(x, train=False)
        z_s, mu_s, ls_s, de_s, z_n, mu_n, ls_n, de_n, skiper, C, Fq, stft_x = r
        if dec_kind == "skip_prepare":
            sig, pred = dec(stft_x, z_s, skiper, C, Fq, train=False)
        else:
            sig, pred = dec(stft_x, z_s, skiper, C, Fq, train=False, pad="sig")
    # ---- oracle port on the same inputs
    esd, dsd = enc.state_dict(), dec.state_dict()
    with torch.no_grad():
        st = P.vae_encoder_forward(esd, x, ZDIM, latent_num, S, eps, causal=causal)
        dd = P.vae_decoder_forward(dsd, st["stft_x"], st["z_speech"], st["skiper"], st["C"], st["F"], S,
                                   recon_type, skip_mode, causal=causal)
        check("stft", st["stft_x"], stft_x)
        check("stft_dense", P.stft_dense(x), stft_x, 1e-5)
        for i in range(6):
            check("enc%d" % i, st["skiper"][i], skiper[i])
        check("miu", st["miu_speech"], mu_s)
        check("log_sigma", st["log_sigma_speech"], ls_s)
        check("delta", st["delta_speech"], de_s)
        check("z_speech", st["z_speech"], z_s)
        if latent_num == 2:
            check("z_noise", st["z_noise"], z_n)
        check("predict", dd["predict"], pred)
        check("recon_sig", dd["recon_sig"], sig)
        check("istft_dense", P.istft_dense(torch.view_as_real(pred)), sig, 1e-5)
    g = {"B": B, "L": L, "S": S, "latent_num": latent_num, "seed": seed, "causal": int(causal),
         "stft_x": np32(stft_x), "miu": np32(mu_s), "log_sigma": np32(ls_s), "delta": np32(de_s),
         "z_speech": np32(z_s), "predict": np32(pred), "recon_sig": np32(sig)}
    if latent_num == 2:
        g.update(z_noise=np32(z_n), miu_noise=np32(mu_n), log_sigma_noise=np32(ls_n), delta_noise=np32(de_n))
    if full:
        for i in range(6):
            g["enc%d" % i] = np32(skiper[i])
        if dec_kind == "skip_prepare":
            for i, o in enumerate(dec.decoder_outputs):
                g["dec%d" % i] = np32(o)
        else:
            for i, o in enumerate(dd["decoder_outputs"]):       # port already checked end-to-end above
                g["dec%d" % i] = np32(o)
    np.savez(os.path.join(OUT, tag + ".npz"), **g)


def case_vae_train(tag, B, L, seed):
    """train=True FORWARD (batch-statistics CBN, running-buffer updates): two consecutive calls so both the
    first-call copy and the EMA branch (model/complex_progress.py:L144-159) are pinned."""
    print("case", tag)
    net, enc = build_vae(1, 1, seed)
    dec = ref_mod.nsvae_pvae_dccrn_decoder_twophase(net, True, "cpu", 1, ZDIM, NFFT, HOP, WIN, "mask", True,
                                                    [0, 1, 2, 3, 4, 5], False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed + 1), strict=True)
    g = {"B": B, "L": L, "seed": seed}
    T = L // HOP + 1
    for call in range(2):
        x = synth_waveform(B, L, seed=1234 + seed + call)
        eps = synth_eps((B, 1, T, ZDIM), seed=7 + seed + call, n=2)
        with torch.no_grad(), supplied_eps(eps):
            r = enc(x, train=True)
            sig, pred = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
        g["miu_%d" % call], g["recon_sig_%d" % call], g["predict_%d" % call] = np32(r[1]), np32(sig), np32(pred)
        g["enc5_%d" % call] = np32(r[8][5])
        for name, mod in (("enc0", enc.encoders[0].bn), ("enc5", enc.encoders[5].bn), ("dec0", dec.decoders[0].bn),
                          ("dec5", dec.decoders[5].bn)):
            for buf in ("running_mean_real", "running_mean_imag", "Vrr", "Vri", "Vii"):
                g["%s_%s_%d" % (name, buf, call)] = np32(getattr(mod, buf))
    np.savez(os.path.join(OUT, tag + ".npz"), **g)


def grad_probe(name, shape, seed=4242):
    """Fixed pseudo-random direction per parameter: <grad, probe> + ||grad|| pin a gradient with two numbers."""
    import zlib
    g = torch.Generator().manual_seed(seed + zlib.crc32(name.encode()) % 100000)
    return torch.randn(shape, generator=g, dtype=torch.float64)


def case_train_step(tag, B, L, latent_num, seed):
    """Phase-1 NSVAE training step of train_nsvae.py:L472-566: frozen clean / noise CVAE encoders (train=False), noisy
    encoder train=True, closed-form KL loss (standard_nsvae_loss_true_kl, w_kl = 1), backward.  The fixture pins the
    loss and, per noisy-encoder parameter, ||grad|| and <grad, probe>; small parameters are stored in full."""
    import types
    print("case", tag)
    for m in ("matplotlib", "matplotlib.pyplot"):                    # unused import in model/nsvae_loss.py:L3
        sys.modules.setdefault(m, types.ModuleType(m))
    import model.nsvae_loss as ref_loss
    net = ref_causal_cfg.get_net_params()
    noisy = ref_mod.nsvae_pvae_dccrn_encoder_twophase(net, True, "cpu", ZDIM, NFFT, HOP, WIN, 1, latent_num)
    noisy.load_state_dict(fill_state_dict(noisy.state_dict(), seed), strict=True)
    frozen = []
    for j in range(2):
        e = ref_mod.pvae_dccrn_encoder_skip_prepare(net, True, "cpu", ZDIM, NFFT, HOP, WIN, 1)
        e.load_state_dict(fill_state_dict(e.state_dict(), seed + 1 + j), strict=True)
        frozen.append(e.eval())
    xs = [synth_waveform(B, L, seed=1234 + seed + j) for j in range(3)]          # noisy, clean, noise
    T = L // HOP + 1
    eps = synth_eps((B, 1, T, ZDIM), seed=7 + seed, n=2 * latent_num)
    with torch.no_grad():
        with supplied_eps(synth_eps((B, 1, T, ZDIM), seed=8 + seed, n=2)):
            rc = frozen[0](xs[1], train=False)
        with supplied_eps(synth_eps((B, 1, T, ZDIM), seed=9 + seed, n=2)):
            rn = frozen[1](xs[2], train=False)
    with supplied_eps(eps):
        r = noisy(xs[0], train=True)
    lossf = ref_loss.standard_nsvae_loss_true_kl(1.0, 0, 1.0, 0, ZDIM, 1, latent_num, "twophase", "False",
                                                 [0, 1, 2, 3, 4, 5], "latent")
    # final_nsvae_loss (L448-473) = w_kl * kl_loss + w_dismiu * miu_dis_loss; with the shipped w_dismiu = 0 the step
    # is kl_loss (L330-347) - called directly because miu_dis_loss cannot run with latent_num == 1 (None operands)
    out = (None,) + tuple(lossf.kl_loss(rc[1], rn[1], r[1], r[5], rc[2], rn[2], r[2], r[6], rc[3], rn[3], r[3], r[7],
                                        r[0], r[4]))
    kl_loss = out[1]
    kl_loss.backward()
    # ---- the port, differentiated by autograd on the same tensors
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and k in dict(noisy.named_parameters()))
          for k, v in noisy.state_dict().items()}
    st = P.vae_encoder_forward(sd, xs[0], ZDIM, latent_num, 1, eps, train=True, grad=True)
    with torch.no_grad():
        sc = P.vae_encoder_forward(frozen[0].state_dict(), xs[1], ZDIM, 1, 1, synth_eps((B, 1, T, ZDIM), seed=8 + seed, n=2))
        sn = P.vae_encoder_forward(frozen[1].state_dict(), xs[2], ZDIM, 1, 1, synth_eps((B, 1, T, ZDIM), seed=9 + seed, n=2))
    pl, pc, pn = P.nsvae_kl_loss(st, sc, sn, ZDIM, latent_num, 1.0)
    pl.backward()
    check("kl loss", pl.detach(), kl_loss.detach(), 1e-4)     # latent_num 1: difference of two large means
    g = {"B": B, "L": L, "latent_num": latent_num, "seed": seed, "loss": np32(kl_loss), "kl_clean": np32(out[2]),
         "kl_noise": np32(out[3]), "miu": np32(r[1])}
    worst = 0.0
    for name, p in noisy.named_parameters():
        if p.grad is None:
            assert name.startswith("dense."), name                     # unused ComplexDense: no gradient
            continue
        if ".conv.conv_" in name and name.endswith(".bias"):
            # a bias in front of a batch-statistics normalisation has an exactly zero gradient; autograd returns
            # round-off (~1e-7), which only an absolute bound can pin
            assert float(p.grad.abs().max()) < 1e-5 and float(sd[name].grad.abs().max()) < 1e-5, name
            g["zero/" + name] = np.float64(p.grad.abs().max())
            continue
        e = P.rel_l2(sd[name].grad, p.grad)
        if e > 1e-4:
            print("   ", name, "rel %.2e  |ref| %.3e |port| %.3e" % (e, float(p.grad.norm()), float(sd[name].grad.norm())))
        worst = max(worst, e)
        gd = p.grad.double()
        g["norm/" + name] = np.float64(gd.norm())
        g["probe/" + name] = np.float64((gd * grad_probe(name, p.shape)).sum())
        if p.numel() <= 1024:
            g["full/" + name] = np32(p.grad)
    print("  port-vs-reference gradients (autograd of the port): worst rel_l2 = %.2e" % worst)
    assert worst < 2e-4, worst
    np.savez(os.path.join(OUT, tag + ".npz"), **g)


def case_train_step_phase2(tag, B, L, latent_num, recon_type, weights, seed, S=1):
    """Phase-2 decoder training step of train_second_phase_decoder.py:L376-433: frozen NSVAE encoder (train=False),
    nsvae_pvae_dccrn_decoder_twophase(train=True, pad='sig'), two_phase_loss.multi_recon_loss (shipped weights '001' =
    SI-SNR only), backward.  Pins the loss terms and, per decoder parameter, ||grad|| and <grad, probe>."""
    import types
    print("case", tag)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    import model.nsvae_loss as ref_loss
    net, enc = build_vae(latent_num, S, seed)
    dec = ref_mod.nsvae_pvae_dccrn_decoder_twophase(net, True, "cpu", S, ZDIM, NFFT, HOP, WIN, recon_type, True,
                                                    [0, 1, 2, 3, 4, 5], False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed + 1), strict=True)
    xs = [synth_waveform(B, L, seed=1234 + seed + j) for j in range(2)]           # noisy, clean
    T = L // HOP + 1
    eps = synth_eps((B, S, T, ZDIM), seed=7 + seed, n=2 * latent_num)
    with torch.no_grad(), supplied_eps(eps):
        r = enc(xs[0], train=False)
    sig, pred = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
    # train_second_phase_decoder.py:L383-390: the clean targets are repeated per sample (row b*S + s <- utterance b)
    clean_rep = xs[1].unsqueeze(1).repeat(1, S, 1).view(B * S, L)
    stft_clean = enc.stft(xs[1]).unsqueeze(1).repeat(1, S, 1, 1, 1).view(B * S, NFFT // 2 + 1, T, 2)
    lossf = ref_loss.two_phase_loss(list(weights), 1.0, ZDIM, latent_num)
    final, l_cpx, l_mag, l_si = lossf.multi_recon_loss(pred, stft_clean, clean_rep, sig)
    final.backward()
    # ---- the port, differentiated by autograd
    params = dict(dec.named_parameters())
    sd = {k: v.detach().clone().requires_grad_(k in params) for k, v in fill_state_dict(dec.state_dict(), seed + 1).items()}
    with torch.no_grad():
        st = P.vae_encoder_forward(enc.state_dict(), xs[0], ZDIM, latent_num, S, eps)
    dd = P.vae_decoder_forward(sd, st["stft_x"], st["z_speech"], st["skiper"], st["C"], st["F"], S, recon_type, "sig",
                               train=True)
    pf, pc, pm, ps = P.multi_recon_loss(dd["predict"], P.stft(xs[1]).repeat_interleave(S, 0), clean_rep, dd["recon_sig"],
                                        weights)
    pf.backward()
    check("phase2 recon_sig", dd["recon_sig"].detach(), sig.detach())
    for nm, a, b in (("final", pf, final), ("cpx", pc, l_cpx), ("mag", pm, l_mag), ("sisnr", ps, l_si)):
        assert abs(float(a) - float(b)) <= 1e-5 * max(1.0, abs(float(b))), (nm, float(a), float(b))
    g = {"B": B, "L": L, "latent_num": latent_num, "seed": seed, "mask": int(recon_type == "mask"), "S": S,
         "weights": np.asarray(weights, dtype=np.float64), "loss": np32(final), "loss_cpx": np32(l_cpx),
         "loss_mag": np32(l_mag), "loss_sisnr": np32(l_si), "recon_sig": np32(sig)}
    worst = 0.0
    for name, p in dec.named_parameters():
        assert p.grad is not None, name
        if ".transconv.tconv_" in name and name.endswith(".bias"):
            assert float(p.grad.abs().max()) < 1e-4 * max(1.0, float(final.abs())), (name, float(p.grad.abs().max()))
            g["zero/" + name] = np.float64(p.grad.abs().max())
            continue
        e = P.rel_l2(sd[name].grad, p.grad)
        if e > 1e-4:
            print("   ", name, "rel %.2e  |ref| %.3e |port| %.3e" % (e, float(p.grad.norm()), float(sd[name].grad.norm())))
        worst = max(worst, e)
        gd = p.grad.double()
        g["norm/" + name] = np.float64(gd.norm())
        g["probe/" + name] = np.float64((gd * grad_probe(name, p.shape)).sum())
        if p.numel() <= 1024:
            g["full/" + name] = np32(p.grad)
    print("  port-vs-reference decoder gradients: worst rel_l2 = %.2e" % worst)
    assert worst < 5e-4, worst
    np.savez(os.path.join(OUT, tag + ".npz"), **g)


def case_train_step_e2e(tag, B, L, latent_num, seed):
    """End-to-end training step: frozen clean / noise CVAE encoders, noisy NSVAE encoder
    train=True, twophase decoder train=True (pad='sig', mask head) on z_speech, loss =
    nsvae_loss_with_cvae_decoder_recon.kl_loss_and_recon_loss (model/nsvae_loss.py:L598-613) with recon weights
    (0, 0, 1): KL + SI-SNR; backward into BOTH models (through z and through the skip tensors)."""
    import types
    print("case", tag)
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    import model.nsvae_loss as ref_loss
    net = ref_causal_cfg.get_net_params()
    noisy = ref_mod.nsvae_pvae_dccrn_encoder_twophase(net, True, "cpu", ZDIM, NFFT, HOP, WIN, 1, latent_num)
    noisy.load_state_dict(fill_state_dict(noisy.state_dict(), seed), strict=True)
    dec = ref_mod.nsvae_pvae_dccrn_decoder_twophase(net, True, "cpu", 1, ZDIM, NFFT, HOP, WIN, "mask", True,
                                                    [0, 1, 2, 3, 4, 5], False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed + 3), strict=True)
    frozen = []
    for j in range(2):
        e = ref_mod.pvae_dccrn_encoder_skip_prepare(net, True, "cpu", ZDIM, NFFT, HOP, WIN, 1)
        e.load_state_dict(fill_state_dict(e.state_dict(), seed + 1 + j), strict=True)
        frozen.append(e.eval())
    xs = [synth_waveform(B, L, seed=1234 + seed + j) for j in range(3)]          # noisy, clean, noise
    T = L // HOP + 1
    eps = synth_eps((B, 1, T, ZDIM), seed=7 + seed, n=2 * latent_num)
    with torch.no_grad():
        with supplied_eps(synth_eps((B, 1, T, ZDIM), seed=8 + seed, n=2)):
            rc = frozen[0](xs[1], train=False)
        with supplied_eps(synth_eps((B, 1, T, ZDIM), seed=9 + seed, n=2)):
            rn = frozen[1](xs[2], train=False)
    with supplied_eps(eps):
        r = noisy(xs[0], train=True)
    sig, pred = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
    lossf = ref_loss.nsvae_loss_with_cvae_decoder_recon(1.0, 1.0, 1.0, [0.0, 0.0, 1.0], latent_num, ZDIM)
    out = lossf.kl_loss_and_recon_loss(rc[1], rn[1], r[1], r[5], rc[2], rn[2], r[2], r[6], rc[3], rn[3], r[3], r[7],
                                       r[0], r[4], pred, noisy.stft(xs[1]), xs[1], sig)
    loss, kl, sisnr = out[0], out[1], out[7]
    loss.backward()
    # ---- the port
    ep, dp = dict(noisy.named_parameters()), dict(dec.named_parameters())
    esd = {k: v.detach().clone().requires_grad_(k in ep) for k, v in fill_state_dict(noisy.state_dict(), seed).items()}
    dsd = {k: v.detach().clone().requires_grad_(k in dp) for k, v in fill_state_dict(dec.state_dict(), seed + 3).items()}
    st = P.vae_encoder_forward(esd, xs[0], ZDIM, latent_num, 1, eps, train=True, grad=True)
    with torch.no_grad():
        sc = P.vae_encoder_forward(frozen[0].state_dict(), xs[1], ZDIM, 1, 1, synth_eps((B, 1, T, ZDIM), seed=8 + seed, n=2))
        sn = P.vae_encoder_forward(frozen[1].state_dict(), xs[2], ZDIM, 1, 1, synth_eps((B, 1, T, ZDIM), seed=9 + seed, n=2))
    dd = P.vae_decoder_forward(dsd, st["stft_x"], st["z_speech"], st["skiper"], st["C"], st["F"], 1, "mask", "sig", train=True)
    pkl, _, _ = P.nsvae_kl_loss(st, sc, sn, ZDIM, latent_num, 1.0)
    psi = P.si_snr(xs[1], dd["recon_sig"])
    (pkl + psi).backward()
    for nm, a, b in (("kl", pkl, kl), ("sisnr", psi, sisnr)):
        assert abs(float(a.detach()) - float(b.detach())) <= 1e-4 * max(1.0, abs(float(b.detach()))), (nm, float(a), float(b))
    g = {"B": B, "L": L, "latent_num": latent_num, "seed": seed, "loss": np32(loss), "kl": np32(kl), "sisnr": np32(sisnr),
         "recon_sig": np32(sig)}
    worst = 0.0
    for pre, mod, sd in (("enc/", noisy, esd), ("dec/", dec, dsd)):
        for name, p in mod.named_parameters():
            if p.grad is None:
                assert pre == "enc/" and name.startswith("dense."), name
                continue
            if name.endswith(".bias") and (".conv.conv_" in name or ".transconv.tconv_" in name):
                assert float(p.grad.abs().max()) < 1e-3, (name, float(p.grad.abs().max()))
                g["zero/" + pre + name] = np.float64(p.grad.abs().max())
                continue
            e = P.rel_l2(sd[name].grad, p.grad)
            if e > 1e-4:
                print("   ", pre + name, "rel %.2e  |ref| %.3e |port| %.3e" % (e, float(p.grad.norm()), float(sd[name].grad.norm())))
            worst = max(worst, e)
            gd = p.grad.double()
            g["norm/" + pre + name] = np.float64(gd.norm())
            g["probe/" + pre + name] = np.float64((gd * grad_probe(name, p.shape)).sum())
            if p.numel() <= 1024:
                g["full/" + pre + name] = np32(p.grad)
    print("  port-vs-reference end-to-end gradients: worst rel_l2 = %.2e" % worst)
    assert worst < 2e-3, worst
    np.savez(os.path.join(OUT, tag + ".npz"), **g)


def case_variant(tag):
    """One of the class variants of oracle/variants.py: the unmodified reference classes on seeded weights / inputs /
    eps (train=False); pins latents, z, the deepest skip tensor and - when the variant has a decoder - the outputs."""
    import copy
    from oracle.variants import VARIANTS, run_variant
    print("case", tag)
    v = VARIANTS[tag]
    net = copy.deepcopy(ref_causal_cfg.get_net_params())             # adapt_channel mutates its net_params
    torch.manual_seed(0)
    enc, dec = v["build"](ref_mod, net, "cpu")
    enc.load_state_dict(fill_state_dict(enc.state_dict(), v["seed"]), strict=True)
    enc.eval()
    if dec is not None:
        dec.load_state_dict(fill_state_dict(dec.state_dict(), v["seed"] + 1), strict=True)
        dec.eval()
    x = synth_waveform(v["B"], v["L"], seed=1234 + v["seed"])
    T = v["L"] // HOP + 1
    eps = synth_eps((v["B"], v["S"], T, ZDIM), seed=7 + v["seed"], n=2 * v["latent_num"])
    with torch.no_grad(), supplied_eps(eps):
        out = run_variant(v, enc, dec, x, None)
    g = {k: np32(t) for k, t in out.items()}
    g.update(B=v["B"], L=v["L"], S=v["S"], latent_num=v["latent_num"], seed=v["seed"])
    np.savez(os.path.join(OUT, tag + ".npz"), **g)
    print("   stored", sorted(k for k in g if k not in ("B", "L", "S", "latent_num", "seed")))


def case_dccrn_datanorm(tag, seed=31):
    """DCCRN_ with data_mean / data_std (model/pvae_module.py:L217-221, L236-249), both heads."""
    import copy
    print("case", tag)
    x = synth_waveform(2, 900, seed=1234 + seed)
    g = {"seed": seed}
    for recon_type in ("mask", "real_imag"):
        net = copy.deepcopy(ref_causal_cfg.get_net_params())
        m = ref_mod.DCCRN_(NFFT, HOP, net, True, "cpu", WIN, [0, 1, 2, 3, 4, 5], recon_type, False,
                           torch.zeros(1, 257, 1, 2), torch.ones(1, 257, 1, 2))
        m.load_state_dict(fill_state_dict(m.state_dict(), seed), strict=True)
        m.eval()
        with torch.no_grad():
            clean, pred = m(x, train=False)
            sd = m.state_dict()
            d = P.dccrn_forward(sd, x, recon_type=recon_type, data_norm=(sd["data_mean"], sd["data_std"]))
        check("dccrn datanorm %s predict" % recon_type, d["predict"], pred)
        check("dccrn datanorm %s clean" % recon_type, d["clean"], clean)
        g["clean_" + recon_type], g["predict_" + recon_type] = np32(clean), np32(pred)
    np.savez(os.path.join(OUT, tag + ".npz"), **g)


def variant_cases():
    from oracle.variants import VARIANTS
    for tag in VARIANTS:
        case_variant(tag)
    case_dccrn_datanorm("dccrn_datanorm")


def n2_cases():
    """pvae_dccrn_decoder_prob_skip and distinguisher of the unmodified reference (oracle/variants.py runners)."""
    from oracle import variants as V
    v = V.PROB_SKIP
    print("case var_prob_skip_dec")
    x = synth_waveform(v["B"], v["L"], seed=1234 + v["seed"])
    T = v["L"] // HOP + 1
    eps = synth_eps((v["B"], v["S"], T, ZDIM), seed=7 + v["seed"], n=2)
    with supplied_eps(eps):
        out = V.run_prob_skip(ref_mod, ref_causal_cfg.get_net_params, fill_state_dict, "cpu", x, None)
    g = {}
    for case, (sig, pred) in out.items():
        g[case + "_sig"], g[case + "_predict"] = np32(sig), np32(pred)
    assert P.rel_l2(g["train_real_sig"], g["train_zero_sig"]) > 1e-2 and P.rel_l2(g["train_zero_sig"], g["train_self_sig"]) > 1e-2
    np.savez(os.path.join(OUT, "var_prob_skip_dec.npz"), **g)
    v = V.DISTINGUISHER
    print("case distinguisher")
    x = synth_waveform(v["B"], v["L"], seed=1234 + v["seed"])
    out = V.run_distinguisher(ref_mod, ref_causal_cfg.get_net_params, fill_state_dict, "cpu", x)
    assert tuple(out["eval0"].shape) == (v["B"], v["L"] // HOP + 1, 1)
    np.savez(os.path.join(OUT, "distinguisher.npz"), **{k: np32(t) for k, t in out.items()})


def case_dccrn(tag, B, L, seed, causal=True):
    print("case", tag)
    net = (ref_causal_cfg if causal else ref_noncausal_cfg).get_net_params()
    m = ref_mod.DCCRN_(NFFT, HOP, net, causal, "cpu", WIN, [0, 1, 2, 3, 4, 5], "mask", False, None, None)
    m.load_state_dict(fill_state_dict(m.state_dict(), seed), strict=True)
    m.eval()
    x = synth_waveform(B, L, seed=1234 + seed)
    with torch.no_grad():
        clean, pred = m(x, train=False)
        d = P.dccrn_forward(m.state_dict(), x, causal=causal)
        check("dccrn.latent", d["latent"], m.std_DCCRN.latent)
        check("dccrn.predict", d["predict"], pred)
        check("dccrn.clean", d["clean"], clean)
    np.savez(os.path.join(OUT, tag + ".npz"), B=B, L=L, seed=seed, latent=np32(m.std_DCCRN.latent),
             predict=np32(pred), clean=np32(clean))


def case_primitives_noncausal(tag, seed):
    """Non-causal complex conv / transposed conv blocks (model/complex_progress.py:L24-36, L253-279) at odd shapes."""
    print("case", tag)
    g = {}
    gen = torch.Generator().manual_seed(200 + seed)
    enc = ref_mod.Encoder(3, 5, (5, 2), (2, 1), (5, 9, 1), padding=(2, 0), causal=False)
    enc.load_state_dict(fill_state_dict(enc.state_dict(), seed))
    x = torch.randn(2, 3, 11, 6, 2, generator=gen)
    with torch.no_grad():
        g["enc_in"], g["enc_out"] = np32(x), np32(enc(x, False))
        check("Encoder block (non-causal)", P.encoder_block(x, enc.state_dict(), "", causal=False), enc(x, False))
    dec = ref_mod.Decoder(4, 3, (5, 2), (2, 1), (3, 9, 1), padding=(2, 0), causal=False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed))
    x = torch.randn(2, 4, 5, 6, 2, generator=gen)
    with torch.no_grad():
        g["dec_in"], g["dec_out"] = np32(x), np32(dec(x, False))
        check("Decoder block (non-causal)", P.decoder_block(x, dec.state_dict(), "", causal=False), dec(x, False))
    assert g["enc_out"].shape[3] == 5 and g["dec_out"].shape[3] == 7
    np.savez(os.path.join(OUT, tag + ".npz"), **g)


def noncausal_cases():
    case_primitives_noncausal("primitives_noncausal", seed=8)
    case_vae("vae_nc_l1_zero_full", B=2, L=1000, latent_num=1, S=1, dec_kind="skip_prepare",
             recon_type="real_imag", seed=9, full=True, causal=False)
    case_vae("vae_nc_l2_sig_mask_s2_e2e", B=2, L=4000, latent_num=2, S=2, dec_kind="twophase",
             recon_type="mask", seed=10, full=False, causal=False)
    case_dccrn("dccrn_nc_mask_e2e", B=3, L=3000, seed=11, causal=False)


def case_primitives(tag, seed):
    """Stand-alone complex primitives at odd shapes (reference classes called directly)."""
    print("case", tag)
    import model.complex_progress as cp
    g = {}
    gen = torch.Generator().manual_seed(100 + seed)
    # causal complex conv + CBN(eval) + PReLU via the reference Encoder / Decoder blocks
    enc = ref_mod.Encoder(3, 5, (5, 2), (2, 1), (5, 9, 1), padding=(2, 1), causal=True)
    enc.load_state_dict(fill_state_dict(enc.state_dict(), seed))
    x = torch.randn(2, 3, 11, 6, 2, generator=gen)
    with torch.no_grad():
        g["enc_in"], g["enc_out"] = np32(x), np32(enc(x, False))
        check("Encoder block", P.encoder_block(x, enc.state_dict(), ""), enc(x, False))
    dec = ref_mod.Decoder(4, 3, (5, 2), (2, 1), (3, 9, 1), padding=(2, 0), causal=True)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed))
    x = torch.randn(2, 4, 5, 6, 2, generator=gen)
    with torch.no_grad():
        g["dec_in"], g["dec_out"] = np32(x), np32(dec(x, False))
        check("Decoder block", P.decoder_block(x, dec.state_dict(), ""), dec(x, False))
    lstm = cp.ComplexLSTM(20, 8, "cpu", num_layers=2)
    lstm.load_state_dict(fill_state_dict(lstm.state_dict(), seed))
    x = torch.randn(7, 3, 20, 2, generator=gen)
    with torch.no_grad():
        g["lstm_in"], g["lstm_out"] = np32(x), np32(lstm(x))
        check("ComplexLSTM", P.complex_lstm(x, lstm.state_dict(), "", 8, 2), lstm(x))
    dense = cp.ComplexDense(128, 24)
    dense.load_state_dict(fill_state_dict(dense.state_dict(), seed))
    x = torch.randn(9, 128, 2, generator=gen)
    with torch.no_grad():
        g["dense_in"], g["dense_out"] = np32(x), np32(dense(x))
        check("ComplexDense", P.complex_dense(x, dense.state_dict(), ""), dense(x))
    np.savez(os.path.join(OUT, tag + ".npz"), **g)


def e2e_cases():
    case_train_step_e2e("train_e2e_l2", B=2, L=1200, latent_num=2, seed=17)
    case_train_step_e2e("train_e2e_l1", B=3, L=2300, latent_num=1, seed=18)


def phase2_cases():
    case_train_step_phase2("train_phase2_mask_sisnr", B=2, L=1200, latent_num=2, recon_type="mask",
                           weights=(0.0, 0.0, 1.0), seed=15)
    case_train_step_phase2("train_phase2_ri_multi", B=3, L=700, latent_num=1, recon_type="real_imag",
                           weights=(0.5, 0.25, 1.0), seed=16)
    # the shipped phase-2 script runs --num_samples 2 (train_second_phase_decoder.sh:L6)
    case_train_step_phase2("train_phase2_mask_sisnr_s2", B=2, L=1300, latent_num=2, recon_type="mask",
                           weights=(0.0, 0.0, 1.0), seed=19, S=2)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if "--only-n2" in sys.argv:
        n2_cases()
        sys.exit(0)
    if "--only-variants" in sys.argv:
        variant_cases()
        sys.exit(0)
    if "--only-e2e" in sys.argv:
        e2e_cases()
        sys.exit(0)
    if "--only-phase2-s2" in sys.argv:
        case_train_step_phase2("train_phase2_mask_sisnr_s2", B=2, L=1300, latent_num=2, recon_type="mask",
                               weights=(0.0, 0.0, 1.0), seed=19, S=2)
        sys.exit(0)
    if "--only-phase2" in sys.argv:
        phase2_cases()
        sys.exit(0)
    if "--only-trainstep" in sys.argv:
        case_train_step("train_step_l2", B=2, L=1200, latent_num=2, seed=13)
        case_train_step("train_step_l1", B=3, L=700, latent_num=1, seed=14)
        sys.exit(0)
    if "--only-noncausal" in sys.argv:
        noncausal_cases()
        sys.exit(0)
    if "--only-train" in sys.argv:
        case_vae_train("vae_train_fwd", B=2, L=800, seed=6)
        sys.exit(0)
    case_vae_train("vae_train_fwd", B=2, L=800, seed=6)
    case_primitives("primitives", seed=3)
    # per-layer fixtures (tiny T, B=2 so the utterance boundary is exercised)
    case_vae("vae_l1_zero_full", B=2, L=400, latent_num=1, S=1, dec_kind="skip_prepare",
             recon_type="real_imag", seed=0, full=True)
    case_vae("vae_l2_sig_mask_full", B=2, L=400, latent_num=2, S=1, dec_kind="twophase",
             recon_type="mask", seed=1, full=True)
    # end-to-end fixtures (T spans more than one 128-row tile; S>1 for the sample replication)
    case_vae("vae_l1_zero_e2e", B=2, L=16000, latent_num=1, S=1, dec_kind="skip_prepare",
             recon_type="real_imag", seed=2, full=False)
    case_vae("vae_l2_sig_mask_s2_e2e", B=2, L=6400, latent_num=2, S=2, dec_kind="twophase",
             recon_type="mask", seed=3, full=False)
    case_vae("vae_l1_sig_ri_e2e", B=3, L=3200, latent_num=1, S=1, dec_kind="twophase",
             recon_type="real_imag", seed=4, full=False)
    case_dccrn("dccrn_mask_e2e", B=2, L=8000, seed=5)
    noncausal_cases()
    case_train_step("train_step_l2", B=2, L=1200, latent_num=2, seed=13)
    case_train_step("train_step_l1", B=3, L=700, latent_num=1, seed=14)
    phase2_cases()
    e2e_cases()
    variant_cases()
    n2_cases()
    print("golden fixtures written to", OUT)
