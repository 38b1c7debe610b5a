"""CPU tier: .ini / directory-name compatibility helpers (host logic of the reference's drivers)."""
import idccrn_b200 as M
from idccrn_b200 import config as CFG

INI = """
[User]
model_name = complex_NSVAE
pre_clean_encoder = /x/2025-05-28-10h43_complex_CVAE_causal=True_zdim=128_numsamples=5_kl_annflag=False_skipc=False_skipuse=[0, 1, 2, 3, 4, 5]_spadd=True_reconloss=multiple_recon=real_imag_reconweight=[1.0, 1.0, 0.0]_prior=ri_inde/complex_CVAE_encoder_best_epoch.pt

[Network]
z_dim = 128

[STFT]
winlen = 400
nfft = 512
hopfrac = 100
fs = 16000

[DataFrame]
sequence_len = 481
"""


def test_ini_is_case_preserving_and_yields_stft(tmp_path):
    p = tmp_path / "config.ini"
    p.write_text(INI)
    cfg = CFG.read_config(str(p))
    assert "pre_clean_encoder" in cfg["User"] and cfg.get("User", "model_name") == "complex_NSVAE"
    assert CFG.stft_params(cfg) == (512, 100, 400) and CFG.zdim(cfg) == 128
    c2 = CFG.myconf()
    c2.read_string("[A]\nCamelKey = 1\n")
    assert list(c2["A"].keys()) == ["CamelKey"]


def test_directory_name_parsing(tmp_path):
    p = tmp_path / "config.ini"
    p.write_text(INI)
    cfg = CFG.read_config(str(p))
    pre = CFG.parse_pretrain_dir(cfg.get("User", "pre_clean_encoder").split("/")[-2])
    assert pre["causal"] is True and pre["spadd"] is True and pre["skipuse"] == [0, 1, 2, 3, 4, 5]
    assert pre["recon_type"] == "real_imag" and pre["skipc"] == "False" and pre["zdim"] == 128
    ns = CFG.parse_nsvae_dir("2025-06-01_complex_NSVAE_zdim=128_nsvae=twophase_latentnum=2_match=speech")
    assert ns == {"zdim": 128, "w_resi": 0.0, "nsvae_model": "twophase", "latent_num": 2, "matching": "speech"}
    assert CFG.parse_pretrain_dir("plain_name")["causal"] is False


def test_build_enhancer_matches_reference_pairing(tmp_path):
    p = tmp_path / "config.ini"
    p.write_text(INI)
    cfg = CFG.read_config(str(p))
    enc, dec = CFG.build_enhancer(cfg, "cpu", num_samples=1, latent_num=1)
    assert isinstance(enc, M.nsvae_pvae_dccrn_encoder_twophase) and isinstance(dec, M.pvae_dccrn_decoder_skip_prepare)
    enc2, dec2 = CFG.build_enhancer(cfg, "cpu", latent_num=2, finetuned_decoder=True)
    assert enc2.lstms[0].hidden_size == 768 and isinstance(dec2, M.nsvae_pvae_dccrn_decoder_twophase)
    assert dec2.recon_type == "mask"


def test_reference_import_shim():
    """``from model.pvae_module import *`` (what the reference's scripts do) resolves to this package's classes."""
    import sys
    from idccrn_b200 import compat
    saved = {n: sys.modules.pop(n) for n in list(sys.modules) if n == "model" or n.startswith("model.")}
    try:
        compat.install()
        ns = {}
        exec("from model.pvae_module import *\nimport model.causal_netconfig as cfgmod\n"
             "from model.complex_progress import ComplexBatchNormal as CBN", ns)
        assert ns["nsvae_pvae_dccrn_encoder_twophase"] is M.nsvae_pvae_dccrn_encoder_twophase
        assert ns["pvae_dccrn_decoder_no_skip"] is M.pvae_dccrn_decoder_no_skip and ns["DCCRN_"] is M.DCCRN_
        assert ns["CBN"] is M.ComplexBatchNormal
        net = ns["cfgmod"].get_net_params()
        assert net["encoder_paddings"][0] == (2, 1) and net["lstm_dim"][0] == 1280
        enc = ns["nsvae_pvae_dccrn_encoder_twophase"](net, True, "cpu", 128, 512, 100, 400, 1, 2)
        assert enc.lstms[0].hidden_size == 768
        import pytest
        with pytest.raises(RuntimeError):
            sys.modules["model.pvae_module"].__idccrn_b200_shim__ = False
            compat.install()
        sys.modules["model.pvae_module"].__idccrn_b200_shim__ = True
    finally:
        compat.uninstall()
        for n in list(sys.modules):
            if n == "model" or n.startswith("model."):
                del sys.modules[n]
        sys.modules.update(saved)
