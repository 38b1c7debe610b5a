"""TEST INFRASTRUCTURE (like everything under oracle/): the encoder / decoder class variants of
model/pvae_module.py that the reference's scripts instantiate besides the headline pair, as one table used by BOTH
oracle/make_golden.py (with ``mod`` = the reference's model.pvae_module) and the parity tests (with ``mod`` =
idccrn_b200) - same constructor calls on both sides."""
import torch

NFFT, HOP, WIN, ZDIM = 512, 100, 400, 128
SKIPS = [0, 1, 2, 3, 4, 5]


def _norm():
    return torch.zeros(1, NFFT // 2 + 1, 1, 2), torch.ones(1, NFFT // 2 + 1, 1, 2)      # filled by fill_state_dict


# kind -> dict(B, L, S, latent_num, seed, enc(mod, net, dev), dec(mod, net, dev) or None, port options)
def _cvae_datanorm(mod, net, dev):
    m, s = _norm()
    enc = mod.pvae_dccrn_encoder(net, True, dev, ZDIM, NFFT, HOP, WIN, 2, m, s)
    m, s = _norm()
    dec = mod.pvae_dccrn_decoder(net, True, dev, 2, ZDIM, NFFT, HOP, WIN, "mask", SKIPS, False, m, s)
    return enc, dec


def _noskip_fc(mod, net, dev):
    enc = mod.pvae_dccrn_encoder_no_skip_fc_latent(net, True, dev, ZDIM, NFFT, HOP, WIN, 1, None, None)
    dec = mod.pvae_dccrn_decoder_no_skip(net, True, dev, 1, ZDIM, NFFT, HOP, WIN, "real_imag")
    return enc, dec


def _noskip_datanorm_resyn(mod, net, dev):
    m, s = _norm()
    enc = mod.pvae_dccrn_encoder_no_skip(net, True, dev, ZDIM, NFFT, HOP, WIN, 1, m, s)
    m, s = _norm()
    dec = mod.pvae_dccrn_decoder_no_skip(net, True, dev, 1, ZDIM, NFFT, HOP, WIN, "real_imag", True, m, s)
    return enc, dec


def _skip_prepare_fc(mod, net, dev):
    enc = mod.pvae_dccrn_encoder_skip_prepare_fc_latent(net, True, dev, ZDIM, NFFT, HOP, WIN, 2)
    dec = mod.pvae_dccrn_decoder_skip_prepare(net, True, dev, 2, ZDIM, NFFT, HOP, WIN, "real_imag", SKIPS)
    return enc, dec


def _twophase_fc(mod, net, dev):
    enc = mod.nsvae_pvae_dccrn_encoder_twophase_fc_latent(net, True, dev, ZDIM, NFFT, HOP, WIN, 1, 2)
    dec = mod.nsvae_pvae_dccrn_decoder_twophase(net, True, dev, 1, ZDIM, NFFT, HOP, WIN, "mask", True, SKIPS, False)
    return enc, dec


def _original_fc(mod, net, dev):
    return mod.nsvae_dccrn_encoder_original_fc_latent(net, True, dev, ZDIM, NFFT, HOP, WIN, 1, 1), None


def _original(mod, net, dev):
    return mod.nsvae_dccrn_encoder_original(net, True, dev, ZDIM, NFFT, HOP, WIN, 1, 2), None


def _double(mod, net, dev):
    return mod.nsvae_dccrn_encoder_double_channel(net, True, dev, ZDIM, NFFT, HOP, WIN, 1, 1), None


def _adapt(mod, net, dev):
    return mod.nsvae_dccrn_encoder_adapt_channel(net, True, dev, ZDIM, NFFT, HOP, WIN, 1, 2, [0, 2, 4]), None


def _prob_skip(mod, net, dev):
    return mod.pvae_dccrn_encoder_prob_skip(net, True, dev, ZDIM, NFFT, HOP, WIN, 1), None


VARIANTS = {
    "var_cvae_datanorm_s2": dict(B=2, L=1500, S=2, latent_num=1, seed=21, build=_cvae_datanorm, pad=None),
    "var_noskip_fc": dict(B=2, L=900, S=1, latent_num=1, seed=22, build=_noskip_fc, pad=None),
    "var_noskip_datanorm_resyn": dict(B=2, L=900, S=1, latent_num=1, seed=23, build=_noskip_datanorm_resyn, pad=None),
    "var_skip_prepare_fc_s2": dict(B=2, L=800, S=2, latent_num=1, seed=24, build=_skip_prepare_fc, pad=None),
    "var_twophase_fc_l2": dict(B=2, L=1100, S=1, latent_num=2, seed=25, build=_twophase_fc, pad="sig"),
    "var_original_fc_l1": dict(B=3, L=600, S=1, latent_num=1, seed=26, build=_original_fc, pad=None),
    "var_original_l2": dict(B=2, L=600, S=1, latent_num=2, seed=27, build=_original, pad=None),
    "var_double_l1": dict(B=2, L=700, S=1, latent_num=1, seed=28, build=_double, pad=None),
    "var_adapt_l2": dict(B=2, L=700, S=1, latent_num=2, seed=29, build=_adapt, pad=None),
    "var_prob_skip_enc": dict(B=2, L=500, S=1, latent_num=1, seed=30, build=_prob_skip, pad=None),
}


def run_variant(v, enc, dec, x, eps):
    """Forward of one variant (train=False) -> dict of the tensors the fixtures pin."""
    r = enc(x, train=False) if eps is None else enc(x, train=False, eps=eps)
    out = {}
    if len(r) == 12:
        names = ("z_speech", "miu", "log_sigma", "delta", "z_noise", "miu_noise", "log_sigma_noise", "delta_noise")
        skiper, C, F, stft_x = r[8:]
    else:
        names = ("z_speech", "miu", "log_sigma", "delta")
        skiper, C, F, stft_x = r[4:]
    for n, t in zip(names, r):
        if t is not None:
            out[n] = t
    out["stft_x"] = stft_x
    out["enc5"] = skiper[5]
    if dec is not None:
        if v["pad"] is not None:
            sig, pred = dec(stft_x, r[0], skiper, C, F, train=False, pad=v["pad"])
        else:
            sig, pred = dec(stft_x, r[0], skiper, C, F, train=False)
        out["recon_sig"], out["predict"] = sig, pred
    return out


# ---- N2 remainder: pvae_dccrn_decoder_prob_skip (model/pvae_module.py:L1681-1788) and the GAN distinguisher
# (L2271-2350).  One runner for the reference classes (make_golden.py --only-n2) and for idccrn_b200 (tests).
# case -> (train, torch seed whose first torch.rand(1) decides the skip drop, skip_prob)
PROB_SKIP_CASES = {
    "eval": (False, 0, 1),
    "train_real": (True, None, 1),          # seed chosen below: draw < 0.5 -> real skips
    "train_zero": (True, None, 1),          # draw >= 0.5, skip_prob 1 -> zero skips
    "train_self": (True, None, 2),          # draw >= 0.5, skip_prob 2 -> the layer's own input as its skip
}
PROB_SKIP = dict(B=2, L=700, S=2, seed=33)


def _seed_for(want_real):
    """Smallest torch seed whose first torch.rand(1) is < 0.5 (want_real) / >= 0.5."""
    for sd in range(1000):
        torch.manual_seed(sd)
        if bool(torch.rand(1)[0] < 0.5) == want_real:
            return sd
    raise RuntimeError("no seed found")


def run_prob_skip(mod, net_fn, fill, device, x, eps):
    """{case: (recon_sig, predict)}.  Every case builds a fresh decoder (train-mode forwards rewrite the CBN buffers);
    the CVAE encoder (pvae_dccrn_encoder_prob_skip, eval) is shared.  The train-mode forwards run under no_grad."""
    v = PROB_SKIP
    enc = mod.pvae_dccrn_encoder_prob_skip(net_fn(), True, device, ZDIM, NFFT, HOP, WIN, v["S"])
    enc.load_state_dict(fill(enc.state_dict(), v["seed"]), strict=True)
    enc = enc.to(device).eval()
    out = {}
    with torch.no_grad():
        r = enc(x, train=False, eps=eps) if eps is not None else enc(x, train=False)
        z, skiper, C, F, stft_x = r[0], r[4], r[5], r[6], r[7]
        for case, (train, sd, skip_prob) in PROB_SKIP_CASES.items():
            dec = mod.pvae_dccrn_decoder_prob_skip(net_fn(), True, device, v["S"], ZDIM, NFFT, HOP, WIN, "real_imag", SKIPS,
                                                   skip_prob)
            dec.load_state_dict(fill(dec.state_dict(), v["seed"] + 1), strict=True)
            dec = dec.to(device)
            torch.manual_seed(_seed_for(case == "train_real") if sd is None else sd)
            out[case] = dec(stft_x, z, skiper, C, F, train=train)
    return out


DISTINGUISHER = dict(B=3, L=900, seed=34)


def run_distinguisher(mod, net_fn, fill, device, x):
    """eval score, then two train-mode forwards (dis_cbn: the running statistics are overwritten both times), then eval
    again with the statistics of the second batch."""
    v = DISTINGUISHER
    d = mod.distinguisher(net_fn(), True, device, ZDIM, NFFT, HOP, WIN)
    d.load_state_dict(fill(d.state_dict(), v["seed"]), strict=True)
    d = d.to(device)
    with torch.no_grad():
        out = {"eval0": d(x, train=False), "train1": d(x, train=True), "train2": d(0.5 * x.flip(0), train=True),
               "eval1": d(x, train=False)}
    return out
