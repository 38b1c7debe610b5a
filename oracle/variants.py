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
