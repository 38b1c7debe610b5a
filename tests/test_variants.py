"""The encoder / decoder class variants of model/pvae_module.py that the reference's scripts instantiate besides the
headline pair: *_fc_latent heads with the clamped reparameterisation, no_skip / real-skip decoders,
data_mean / data_std normalisation, channel-doubling NSVAE encoders.  Same constructor calls as the reference
(oracle/variants.py); outputs pinned by fixtures written from the unmodified reference classes."""
import copy

import pytest
import torch

import common as C
import idccrn_b200 as M
from idccrn_b200.synth import fill_state_dict, synth_eps, synth_waveform
from oracle import variants as V
from oracle.variants import VARIANTS, run_variant


def run_case(golden, tag, device, tol):
    g = golden(tag)
    v = VARIANTS[tag]
    net = copy.deepcopy(M.get_net_params())
    enc, dec = v["build"](M, net, device)
    enc.load_state_dict(fill_state_dict(enc.state_dict(), v["seed"]), strict=True)
    enc = enc.to(device).eval()
    if dec is not None:
        dec.load_state_dict(fill_state_dict(dec.state_dict(), v["seed"] + 1), strict=True)
        dec = dec.to(device).eval()
    x = synth_waveform(v["B"], v["L"], seed=1234 + v["seed"]).to(device)
    T = v["L"] // C.HOP + 1
    eps = [e.to(device) for e in synth_eps((v["B"], v["S"], T, C.ZDIM), seed=7 + v["seed"], n=2 * v["latent_num"])]
    with torch.no_grad():
        out = run_variant(v, enc, dec, x, eps)
    errs = {}
    for k, t in out.items():
        t = torch.view_as_real(t) if t.is_complex() else t
        assert tuple(t.shape) == tuple(g[k].shape), (k, tuple(t.shape), g[k].shape)
        errs[k] = C.rel_l2(t, g[k])
    print(tag, device, errs)
    bad = {k: e for k, e in errs.items() if not e < tol}
    assert not bad, bad


def test_variant_state_dict_keys_match_reference_layout():
    """Key names of the variant-specific parameters (heads, data_norm buffers) as the reference registers them."""
    net = copy.deepcopy(M.get_net_params())
    e = M.nsvae_pvae_dccrn_encoder_twophase_fc_latent(net, True, "cpu", 128, 512, 100, 400, 1, 2)
    keys = set(e.state_dict())
    assert "speech_dense_mean.linear_read.weight" in keys and "noise_dense_delta.linear_imag.bias" in keys
    assert not any(k.startswith("dense.") for k in keys)
    assert e.state_dict()["lstms.0.lstm_re.weight_hh_l1"].shape == (4 * 128, 128)
    e = M.pvae_dccrn_encoder_no_skip_fc_latent(net, True, "cpu", 128, 512, 100, 400, 1, None, None)
    assert "dense_logvar.linear_read.bias" in e.state_dict() and "data_mean" not in e.state_dict()
    m = torch.zeros(1, 257, 1, 2)
    e = M.pvae_dccrn_encoder(net, True, "cpu", 128, 512, 100, 400, 1, m, m + 1)
    assert "data_mean" in e.state_dict() and "dense.linear_read.weight" in e.state_dict()
    d = M.pvae_dccrn_decoder_no_skip(net, True, "cpu", 1, 128, 512, 100, 400, "mask")
    assert d.state_dict()["decoders.0.transconv.tconv_re.weight"].shape == (256, 256, 5, 2)
    e = M.nsvae_dccrn_encoder_double_channel(net, True, "cpu", 128, 512, 100, 400, 1, 1)
    assert e.state_dict()["encoders.5.conv.conv_re.weight"].shape == (512, 512, 5, 2)
    assert e.state_dict()["lstms.0.lstm_re.weight_ih_l0"].shape == (4 * 384, 2560)


@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_variants_emulated(emulated_abi, gemm_mode, golden, tag):
    if gemm_mode == "simt" and tag not in ("var_noskip_fc", "var_cvae_datanorm_s2"):
        pytest.skip("the fp32 SIMT path is cross-checked on two variants")
    run_case(golden, tag, "cpu", 2e-5 if gemm_mode == "simt" else 5e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_variants_gpu(golden, tag):
    run_case(golden, tag, "cuda", 1e-4)


def _dccrn_datanorm(mod, device, recon_type):
    net = copy.deepcopy(M.get_net_params())
    m = mod.DCCRN_(512, 100, net, True, device, 400, C.SKIPS, recon_type, False, torch.zeros(1, 257, 1, 2),
                   torch.ones(1, 257, 1, 2))
    m.load_state_dict(fill_state_dict(m.state_dict(), 31), strict=True)
    return m.to(device).eval()


def run_dccrn_datanorm(golden, device, tol):
    """DCCRN_ with data_mean / data_std (model/pvae_module.py:L217-221, L236-249) against the reference's outputs."""
    g = golden("dccrn_datanorm")
    x = synth_waveform(2, 900, seed=1265)
    for recon_type in ("mask", "real_imag"):
        m = _dccrn_datanorm(M, device, recon_type)
        with torch.no_grad():
            clean, pred = m(x.to(device), train=False)
        assert C.rel_l2(torch.view_as_real(pred), g["predict_" + recon_type]) < tol
        assert C.rel_l2(clean, g["clean_" + recon_type]) < tol


def test_dccrn_datanorm_emulated(emulated_abi, gemm_mode, golden):
    run_dccrn_datanorm(golden, "cpu", 5e-5)


@pytest.mark.gpu
def test_dccrn_datanorm_gpu(golden):
    run_dccrn_datanorm(golden, "cuda", 1e-4)


# ---- N2 remainder: pvae_dccrn_decoder_prob_skip (model/pvae_module.py:L1681-1788), distinguisher (L2271-2350) ---------
def run_prob_skip(golden, device, tol):
    """eval + the three train-mode skip modes (real / zero / the layer's own input), decided by the same torch.rand(1)
    draw as in the reference (the runner seeds torch's global generator identically on both sides); num_samples = 2."""
    g = golden("var_prob_skip_dec")
    v = V.PROB_SKIP
    x = synth_waveform(v["B"], v["L"], seed=1234 + v["seed"]).to(device)
    T = v["L"] // C.HOP + 1
    eps = [e.to(device) for e in synth_eps((v["B"], v["S"], T, C.ZDIM), seed=7 + v["seed"], n=2)]
    out = V.run_prob_skip(M, lambda: copy.deepcopy(M.get_net_params()), fill_state_dict, device, x, eps)
    errs = {}
    for case, (sig, pred) in out.items():
        assert tuple(sig.shape) == g[case + "_sig"].shape
        errs[case + "_sig"] = C.rel_l2(sig, g[case + "_sig"])
        errs[case + "_predict"] = C.rel_l2(torch.view_as_real(pred), g[case + "_predict"])
    print("prob_skip", device, errs)
    assert all(e < tol for e in errs.values()), errs


def run_distinguisher(golden, device, tol):
    g = golden("distinguisher")
    v = V.DISTINGUISHER
    x = synth_waveform(v["B"], v["L"], seed=1234 + v["seed"]).to(device)
    out = V.run_distinguisher(M, lambda: copy.deepcopy(M.get_net_params()), fill_state_dict, device, x)
    errs = {}
    for k, t in out.items():
        assert tuple(t.shape) == g[k].shape, (k, tuple(t.shape), g[k].shape)
        errs[k] = C.rel_l2(t, g[k])
    print("distinguisher", device, errs)
    assert all(e < tol for e in errs.values()), errs


def test_n2_state_dict_keys():
    net = copy.deepcopy(M.get_net_params())
    d = M.distinguisher(net, True, "cpu", 128, 512, 100, 400)
    sd = d.state_dict()
    assert sd["lstms.0.weight_ih_l0"].shape == (4, 2560) and sd["lstms.0.weight_hh_l1"].shape == (4, 1)
    assert "encoders.5.bn.Vri" in sd and d.encoders[0].bn.dis_cbn is True
    p = M.pvae_dccrn_decoder_prob_skip(net, True, "cpu", 2, 128, 512, 100, 400, "real_imag", C.SKIPS, 2)
    assert p.state_dict()["decoders.5.transconv.tconv_re.weight"].shape == (64, 1, 5, 2) and p.zero_flag is False
    with pytest.raises(ValueError):
        M.pvae_dccrn_decoder_prob_skip(net, True, "cpu", 1, 128, 512, 100, 400, "real_imag", C.SKIPS, 0)


def test_prob_skip_decoder_emulated(emulated_abi, gemm_mode, golden):
    run_prob_skip(golden, "cpu", 2e-5 if gemm_mode == "simt" else 5e-5)


def test_distinguisher_emulated(emulated_abi, gemm_mode, golden):
    run_distinguisher(golden, "cpu", 2e-5 if gemm_mode == "simt" else 5e-5)


@pytest.mark.gpu
def test_prob_skip_decoder_gpu(golden):
    run_prob_skip(golden, "cuda", 1e-4)


@pytest.mark.gpu
def test_distinguisher_gpu(golden):
    run_distinguisher(golden, "cuda", 1e-4)
