"""Config-side compatibility with the reference's drivers (host logic only, no compute).

* ``myconf``: the reference's case-preserving ConfigParser (utils/read_config.py:L15-18) — the shipped
  ``configs/*.ini`` parse unchanged; the hot path needs ``[STFT] winlen, nfft, hopfrac`` and ``[Network] z_dim``.
* ``parse_pretrain_dir`` / ``parse_nsvae_dir``: the hyper-parameters the reference encodes in checkpoint *directory
  names* and re-parses at test time (test_nsvae_se.py:L669-727).
* ``build_enhancer``: builds the (encoder, decoder) pair of test_nsvae_se.py:L735-779 (live branch) /
  test_se_cvaefinetune.py:L678-688 from those settings.
"""
import re
from configparser import ConfigParser


class myconf(ConfigParser):
    """ConfigParser that keeps option names case-sensitive (reference: utils/read_config.py:L15-18)."""

    def __init__(self, defaults=None):
        ConfigParser.__init__(self, defaults=None)

    def optionxform(self, optionstr):
        return optionstr


def read_config(path):
    cfg = myconf()
    if not cfg.read(path):
        raise FileNotFoundError(path)
    return cfg


def stft_params(cfg):
    """(n_fft, hop, win_length) from the [STFT] section (nfft, hopfrac, winlen)."""
    return cfg.getint("STFT", "nfft"), cfg.getint("STFT", "hopfrac"), cfg.getint("STFT", "winlen")


def zdim(cfg, default=128):
    return cfg.getint("Network", "z_dim", fallback=default)


def parse_pretrain_dir(name):
    """Settings encoded in a pre-trained CVAE/NVAE directory name, e.g.
    ``..._complex_CVAE_causal=True_zdim=128_..._skipc=False_skipuse=[0, 1, 2, 3, 4, 5]_spadd=True_..._recon=real_imag_...``.
    Defaults follow test_nsvae_se.py:L669-700 (missing keys -> skipuse all, causal False, spadd False, fcl False)."""
    out = {"skipuse": [0, 1, 2, 3, 4, 5] if "skipuse" not in name else None, "causal": False, "spadd": False,
           "fcl": False, "skipc": None, "recon_type": None}
    m = re.search(r"skipuse=\[([0-9, ]*)\]", name)
    if m:
        out["skipuse"] = [int(v) for v in m.group(1).split(",") if v.strip()]
    for key, field in (("causal", "causal"), ("spadd", "spadd"), ("fcl", "fcl")):
        m = re.search(r"%s=(True|False|true|false)" % key, name)
        if m:
            out[field] = m.group(1).lower() == "true"
    m = re.search(r"skipc=([A-Za-z]+)", name)
    if m:
        out["skipc"] = m.group(1)
    m = re.search(r"recon=([a-z]+(?:_imag)?)", name)
    if m:
        out["recon_type"] = "real_imag" if m.group(1) in ("real", "real_imag") else m.group(1)
    m = re.search(r"zdim=(\d+)", name)
    if m:
        out["zdim"] = int(m.group(1))
    return out


def parse_nsvae_dir(name):
    """Settings encoded in an NSVAE directory name (test_nsvae_se.py:L702-727)."""
    out = {"zdim": 0, "w_resi": 0.0, "nsvae_model": "original", "latent_num": 1, "matching": "speech"}
    for s in name.split("_"):
        if "zdim" in s:
            out["zdim"] = int(s.split("=")[-1])
        elif "wresi" in s:
            out["w_resi"] = float(s.split("=")[-1])
        elif "nsvae=" in s:
            out["nsvae_model"] = s.split("=")[-1]
        elif "latentnum" in s:
            out["latent_num"] = int(s.split("=")[-1])
        elif "match" in s:
            out["matching"] = s.split("=")[-1]
    return out


def build_enhancer(cfg, device, num_samples=1, latent_num=1, finetuned_decoder=False, recon_type=None,
                   skip_to_use=(0, 1, 2, 3, 4, 5), causal=True):
    """(noisy encoder, clean decoder) as the reference's test scripts build them: NSVAE encoder
    ``nsvae_pvae_dccrn_encoder_twophase`` + either the pre-trained CVAE decoder (zero skips, real_imag) or the
    fine-tuned two-phase decoder (real skips via pad='sig', mask head)."""
    from . import modules as M
    from .netconfig import get_net_params
    n_fft, hop, win = stft_params(cfg)
    z = zdim(cfg)
    net = get_net_params(causal)
    enc = M.nsvae_pvae_dccrn_encoder_twophase(net, causal, device, z, n_fft, hop, win, num_samples, latent_num)
    if finetuned_decoder:
        dec = M.nsvae_pvae_dccrn_decoder_twophase(net, causal, device, num_samples, z, n_fft, hop, win,
                                                  recon_type or "mask", True, list(skip_to_use), False)
    else:
        dec = M.pvae_dccrn_decoder_skip_prepare(net, causal, device, num_samples, z, n_fft, hop, win,
                                                recon_type or "real_imag", list(skip_to_use))
    return enc, dec
