"""Shared builders for the parity tests: identical synthetic weights / inputs / eps on every machine."""
import torch

import idccrn_b200 as M
from idccrn_b200.synth import fill_state_dict, synth_eps, synth_waveform

NFFT, HOP, WIN, ZDIM = 512, 100, 400, 128
SKIPS = [0, 1, 2, 3, 4, 5]


def build_vae(latent_num, S, dec_kind, recon_type, seed, device, causal=True):
    net = M.get_net_params(causal)
    enc = M.nsvae_pvae_dccrn_encoder_twophase(net, causal, device, ZDIM, NFFT, HOP, WIN, S, latent_num)
    enc.load_state_dict(fill_state_dict(enc.state_dict(), seed), strict=True)
    if dec_kind == "skip_prepare":
        dec = M.pvae_dccrn_decoder_skip_prepare(net, causal, device, S, ZDIM, NFFT, HOP, WIN, recon_type, SKIPS)
    else:
        dec = M.nsvae_pvae_dccrn_decoder_twophase(net, causal, device, S, ZDIM, NFFT, HOP, WIN, recon_type, True,
                                                  SKIPS, False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed + 1), strict=True)
    return enc.to(device).eval(), dec.to(device).eval()


def vae_inputs(B, L, S, latent_num, seed, device, causal=True):
    x = synth_waveform(B, L, seed=1234 + seed).to(device)
    T = L // HOP + 1 - (0 if causal else 6)          # latent frames (non-causal: one fewer per encoder layer)
    eps = [e.to(device) for e in synth_eps((B, S, T, ZDIM), seed=7 + seed, n=2 * latent_num)]
    return x, eps


def run_vae(enc, dec, x, eps, dec_kind):
    with torch.no_grad():
        r = enc(x, train=False, eps=eps)
        z_s, mu_s, ls_s, de_s, z_n, mu_n, ls_n, de_n, skiper, C, F, stft_x = r
        if dec_kind == "skip_prepare":
            sig, pred = dec(stft_x, z_s, skiper, C, F, train=False)
        else:
            sig, pred = dec(stft_x, z_s, skiper, C, F, train=False, pad="sig")
    return dict(stft_x=stft_x, miu=mu_s, log_sigma=ls_s, delta=de_s, z_speech=z_s, z_noise=z_n, skiper=skiper,
                predict=torch.view_as_real(pred), recon_sig=sig)


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float(torch.linalg.norm((a - b).flatten()) / (torch.linalg.norm(b.flatten()) + 1e-30))
