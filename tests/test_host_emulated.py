"""CPU tier: the package's host logic (packing, CBN fold, tap tables, plane layout, module wiring) run
over the emulated C-ABI contract must reproduce the reference goldens."""
import numpy as np
import pytest
import torch

import common as C

TOL = 2e-5   # fp32 golden vs folded weights evaluated through an fp64 emulator (tc mode: plus the bf16x3 split)


@pytest.mark.parametrize("tag,latent_num,S,dec_kind,recon,seed", [
    ("vae_l1_zero_full", 1, 1, "skip_prepare", "real_imag", 0),
    ("vae_l2_sig_mask_full", 2, 1, "twophase", "mask", 1),
    ("vae_nc_l1_zero_full", 1, 1, "skip_prepare", "real_imag", 9),          # non-causal net (model/net_config.py)
])
def test_vae_layers_match_reference(emulated_abi, gemm_mode, golden, tag, latent_num, S, dec_kind, recon, seed):
    g = golden(tag)
    B, L = int(g["B"]), int(g["L"])
    causal = bool(int(g.get("causal", 1)))
    enc, dec = C.build_vae(latent_num, S, dec_kind, recon, seed, "cpu", causal)
    dec.keep_decoder_outputs = True
    x, eps = C.vae_inputs(B, L, S, latent_num, seed, "cpu", causal)
    out = C.run_vae(enc, dec, x, eps, dec_kind)
    errs = {"stft_x": C.rel_l2(out["stft_x"], g["stft_x"])}
    for i in range(6):
        errs["enc%d" % i] = C.rel_l2(out["skiper"][i], g["enc%d" % i])
    for k in ("miu", "log_sigma", "delta", "z_speech", "predict", "recon_sig"):
        errs[k] = C.rel_l2(out[k], g[k])
    if latent_num == 2:
        errs["z_noise"] = C.rel_l2(out["z_noise"], g["z_noise"])
    for i in range(5):
        errs["dec%d" % i] = C.rel_l2(dec.decoder_outputs[i], g["dec%d" % i])
    bad = {k: v for k, v in errs.items() if not v < TOL}
    print(gemm_mode, {k: "%.1e" % v for k, v in errs.items()})
    assert not bad, errs


@pytest.mark.parametrize("tag,latent_num,S,dec_kind,recon,seed", [
    ("vae_l2_sig_mask_s2_e2e", 2, 2, "twophase", "mask", 3),
    ("vae_l1_sig_ri_e2e", 1, 1, "twophase", "real_imag", 4),
    ("vae_nc_l2_sig_mask_s2_e2e", 2, 2, "twophase", "mask", 10),            # non-causal, real skips, 2 samples
])
def test_vae_e2e_match_reference(emulated_abi, gemm_mode, golden, tag, latent_num, S, dec_kind, recon, seed):
    g = golden(tag)
    B, L = int(g["B"]), int(g["L"])
    causal = bool(int(g.get("causal", 1)))
    enc, dec = C.build_vae(latent_num, S, dec_kind, recon, seed, "cpu", causal)
    x, eps = C.vae_inputs(B, L, S, latent_num, seed, "cpu", causal)
    out = C.run_vae(enc, dec, x, eps, dec_kind)
    errs = {k: C.rel_l2(out[k], g[k]) for k in ("stft_x", "miu", "log_sigma", "delta", "z_speech", "predict", "recon_sig")}
    assert all(v < TOL for v in errs.values()), errs


@pytest.mark.parametrize("tag,causal", [("dccrn_mask_e2e", True), ("dccrn_nc_mask_e2e", False)])
def test_dccrn_matches_reference(emulated_abi, gemm_mode, golden, tag, causal):
    import idccrn_b200 as M
    from idccrn_b200.synth import fill_state_dict, synth_waveform
    g = golden(tag)
    B, L, seed = int(g["B"]), int(g["L"]), int(g["seed"])
    m = M.DCCRN_(C.NFFT, C.HOP, M.get_net_params(causal), causal, "cpu", C.WIN, C.SKIPS, "mask", False, None, None)
    m.load_state_dict(fill_state_dict(m.state_dict(), seed), strict=True)
    x = synth_waveform(B, L, seed=1234 + seed)
    with torch.no_grad():
        clean, pred = m(x, train=False)
    errs = {"latent": C.rel_l2(m.std_DCCRN.latent, g["latent"]),
            "predict": C.rel_l2(torch.view_as_real(pred), g["predict"]),
            "clean": C.rel_l2(clean, g["clean"])}
    assert all(v < TOL for v in errs.values()), errs


def test_primitives_match_reference(emulated_abi, golden):
    import idccrn_b200 as M
    from idccrn_b200.synth import fill_state_dict
    g = golden("primitives")
    seed = 3
    enc = M.Encoder(3, 5, (5, 2), (2, 1), (5, 9, 1), padding=(2, 1), causal=True)
    enc.load_state_dict(fill_state_dict(enc.state_dict(), seed))
    dec = M.Decoder(4, 3, (5, 2), (2, 1), (3, 9, 1), padding=(2, 0), causal=True)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed))
    lstm = M.ComplexLSTM(20, 8, "cpu", num_layers=2)
    lstm.load_state_dict(fill_state_dict(lstm.state_dict(), seed))
    dense = M.ComplexDense(128, 24)
    dense.load_state_dict(fill_state_dict(dense.state_dict(), seed))
    t = lambda k: torch.from_numpy(g[k])
    with torch.no_grad():
        errs = {"enc": C.rel_l2(enc(t("enc_in"), False), g["enc_out"]),
                "dec": C.rel_l2(dec(t("dec_in"), False), g["dec_out"]),
                "lstm": C.rel_l2(lstm(t("lstm_in")), g["lstm_out"]),
                "dense": C.rel_l2(dense(t("dense_in")), g["dense_out"])}
    assert all(v < TOL for v in errs.values()), errs


def test_noncausal_primitives_match_reference(emulated_abi, gemm_mode, golden):
    """Stand-alone non-causal blocks: T-1 frames out of the conv, T+1 out of the transposed conv."""
    import idccrn_b200 as M
    from idccrn_b200.synth import fill_state_dict
    g = golden("primitives_noncausal")
    seed = 8
    enc = M.Encoder(3, 5, (5, 2), (2, 1), (5, 9, 1), padding=(2, 0), causal=False)
    enc.load_state_dict(fill_state_dict(enc.state_dict(), seed))
    dec = M.Decoder(4, 3, (5, 2), (2, 1), (3, 9, 1), padding=(2, 0), causal=False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed))
    t = lambda k: torch.from_numpy(g[k])
    with torch.no_grad():
        eo, do = enc(t("enc_in"), False), dec(t("dec_in"), False)
    assert tuple(eo.shape) == g["enc_out"].shape and tuple(do.shape) == g["dec_out"].shape
    errs = {"enc": C.rel_l2(eo, g["enc_out"]), "dec": C.rel_l2(do, g["dec_out"])}
    assert all(v < TOL for v in errs.values()), errs


def run_train_case(golden, device, tol):
    """train=True forward, two consecutive calls (first-call copy, then EMA of the running statistics)."""
    import idccrn_b200 as M
    from idccrn_b200.synth import fill_state_dict, synth_eps, synth_waveform
    g = golden("vae_train_fwd")
    B, L, seed = int(g["B"]), int(g["L"]), int(g["seed"])
    enc, dec = C.build_vae(1, 1, "twophase", "mask", seed, device)
    errs = {}
    for call in range(2):
        x = synth_waveform(B, L, seed=1234 + seed + call).to(device)
        eps = [e.to(device) for e in synth_eps((B, 1, L // C.HOP + 1, C.ZDIM), seed=7 + seed + call, n=2)]
        with torch.no_grad():
            r = enc(x, train=True, eps=eps)
            sig, pred = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
        errs["miu_%d" % call] = C.rel_l2(r[1], g["miu_%d" % call])
        errs["enc5_%d" % call] = C.rel_l2(r[8][5], g["enc5_%d" % call])
        errs["predict_%d" % call] = C.rel_l2(torch.view_as_real(pred), g["predict_%d" % call])
        errs["recon_sig_%d" % call] = C.rel_l2(sig, g["recon_sig_%d" % call])
        for name, mod in (("enc0", enc.encoders[0].bn), ("enc5", enc.encoders[5].bn), ("dec0", dec.decoders[0].bn),
                          ("dec5", dec.decoders[5].bn)):
            for buf in ("running_mean_real", "running_mean_imag", "Vrr", "Vri", "Vii"):
                errs["%s_%s_%d" % (name, buf, call)] = C.rel_l2(getattr(mod, buf), g["%s_%s_%d" % (name, buf, call)])
    # running cross-covariances are small differences of larger moments: their relative error amplifies the
    # activation noise of the split-bf16 path, so the statistics get a 10x looser bound than the activations
    bad = {k: v for k, v in errs.items() if not v < (10 * tol if ("_V" in k or "running_mean" in k) else tol)}
    assert not bad, bad
    assert enc.encoders[0].bn.init_flag is False


def test_train_mode_forward_matches_reference(emulated_abi, gemm_mode, golden):
    run_train_case(golden, "cpu", 5e-5)


def test_train_mode_multi_sample_decoder_matches_oracle(emulated_abi):
    """train=True with num_samples = 2 (no autograd): the batch statistics of every ComplexBatchNormal span all B*S rows
    (model/pvae_module.py:L2550-2567), so the samples run as one batch with repeated skip tensors / noisy STFT."""
    from oracle import ref_port as P
    enc, dec = C.build_vae(1, 2, "twophase", "mask", 0, "cpu")
    x, eps = C.vae_inputs(2, 900, 2, 1, 0, "cpu")
    dsd = {k: v.clone() for k, v in dec.state_dict().items()}
    with torch.no_grad():
        r = enc(x, train=False, eps=eps)
        sig, pred = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
        st = P.vae_encoder_forward(enc.state_dict(), x, C.ZDIM, 1, 2, eps)
        dd = P.vae_decoder_forward(dsd, st["stft_x"], st["z_speech"], st["skiper"], st["C"], st["F"], 2, "mask", "sig",
                                   train=True)
    assert sig.shape[0] == 4
    assert C.rel_l2(sig, dd["recon_sig"]) < 5e-5 and C.rel_l2(torch.view_as_real(pred), torch.view_as_real(dd["predict"])) < 5e-5
