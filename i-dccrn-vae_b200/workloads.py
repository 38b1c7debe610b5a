"""The five BASELINE.json configurations as ready-to-time step functions (synthetic 16 kHz input, deterministic
synthetic weights from synth.py - there are no checkpoints or datasets offline).  bench.py's ``configs`` object and the
tools/ scripts are built on these; nothing here touches the oracle.

  config 1  CVAE reconstruction          pvae_dccrn_encoder_skip_prepare + pvae_dccrn_decoder_skip_prepare, B = 4 x 4 s
  config 2  NSVAE enhancement            nsvae_pvae_dccrn_encoder_twophase(latent_num=1) + CVAE decoder (zero skips), B = 64 x 4 s
  config 2b shipped final system         latent_num=2 (H = 768) + nsvae_pvae_dccrn_decoder_twophase(pad='sig', mask)
  config 3  supervised DCCRN             DCCRN_(mask), 256 x 10 s split over the ranks (strong scaling), <= 64 utterances per pass
  config 4  NSVAE training step          frozen clean / noise encoders + noisy encoder and decoder fwd/bwd, KL + SI-SNR,
                                         gradient all-reduce, Adam; 32 x 4 s per GPU
  config 5  frame streaming              causal DCCRN-VAE, 128 streams per GPU, one hop per step
"""
import torch

from . import modules as M
from . import lib
from .netconfig import get_net_params
from .synth import fill_state_dict, synth_waveform

FS, HOP, NFFT, WIN, ZDIM = 16000, 100, 512, 400, 128
SKIPS = [0, 1, 2, 3, 4, 5]


def algorithmic_gmac_per_utt(T, H=384, real_skip=False):
    """SURVEY §8(d) formulae (real MACs; complex conv = 4 real convs; zero-skip decoder counted at its
    effective, halved K)."""
    enc_c = [1, 32, 64, 128, 128, 256, 256]
    f = [257, 129, 65, 33, 17, 9, 5]
    dec_c = [256, 256, 128, 128, 64, 32, 1]
    g = {}
    g["enc"] = [4 * enc_c[i] * enc_c[i + 1] * 10 * f[i + 1] * T / 1e9 for i in range(6)]
    g["dec"] = [4 * (dec_c[i] + (enc_c[6 - i] if real_skip else 0)) * dec_c[i + 1] * 10 * f[6 - i] * T / 1e9
                for i in range(6)]
    g["lstm_inproj0"] = 4 * (4 * H * 1280) * T / 1e9
    g["lstm_inproj1"] = 4 * (4 * H * H) * T / 1e9            # runs inside the wavefront LSTM kernel
    g["lstm_inproj"] = g["lstm_inproj0"] + g["lstm_inproj1"]
    g["lstm_rec"] = 4 * (4 * H * 2 * H) * T / 1e9
    g["dense"] = 2 * (H if H == 128 else 128) * 1280 * T / 1e9
    g["stft"] = T * 512 * 514 / 1e9
    g["istft"] = T * 512 * 514 / 1e9
    # launches of idv_tapgemm_tc in one step: enc1-5, LSTM layer-0 in-proj, dense, dec0-4, iSTFT frames GEMM
    g["tapgemm"] = sum(g["enc"][1:]) + sum(g["dec"][:5]) + g["lstm_inproj0"] + g["dense"] + g["istft"]
    g["total"] = sum(g["enc"]) + sum(g["dec"]) + g["lstm_inproj"] + g["lstm_rec"] + g["dense"] + g["stft"] + g["istft"]
    return g


def build_vae(latent_num, S, dec_kind, recon_type, seed, device, causal=True, cvae_encoder=False):
    """(encoder, decoder) of the VAE system with synthetic weights, eval mode."""
    net = get_net_params(causal)
    if cvae_encoder:
        enc = M.pvae_dccrn_encoder_skip_prepare(net, causal, device, ZDIM, NFFT, HOP, WIN, S)
    else:
        enc = M.nsvae_pvae_dccrn_encoder_twophase(net, causal, device, ZDIM, NFFT, HOP, WIN, S, latent_num)
    enc.load_state_dict(fill_state_dict(enc.state_dict(), seed), strict=True)
    if dec_kind == "skip_prepare":
        dec = M.pvae_dccrn_decoder_skip_prepare(net, causal, device, S, ZDIM, NFFT, HOP, WIN, recon_type, SKIPS)
    else:
        dec = M.nsvae_pvae_dccrn_decoder_twophase(net, causal, device, S, ZDIM, NFFT, HOP, WIN, recon_type, True, SKIPS, False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), seed + 1), strict=True)
    return enc.to(device).eval(), dec.to(device).eval()


def vae_step(enc, dec, pad=None):
    """x (B, L) -> enhanced waveform through the reference's call sequence (test_nsvae_se.py:L300-352)."""
    def step(x):
        with torch.no_grad():
            r = enc(x, train=False)
            stft_x, z, skiper, C, F = r[-1], r[0], r[-4], r[-3], r[-2]
            if pad is None:
                return dec(stft_x, z, skiper, C, F, train=False)[0]
            return dec(stft_x, z, skiper, C, F, train=False, pad=pad)[0]
    return step


def config1(device, rank=0, batch=4, seconds=4):
    enc, dec = build_vae(1, 1, "skip_prepare", "real_imag", 0, device, cvae_encoder=True)
    x = synth_waveform(batch, FS * seconds, rank=rank).to(device)
    f = vae_step(enc, dec)
    return (lambda: f(x)), {"workload": "config1: CVAE reconstruction (pvae_dccrn_encoder_skip_prepare + "
                                        "pvae_dccrn_decoder_skip_prepare, zero skips, real_imag)",
                            "batch_per_gpu": batch, "utterance_s": seconds, "H": 384, "real_skip": False}


def config2(device, rank=0, batch=64, seconds=4):
    enc, dec = build_vae(1, 1, "skip_prepare", "real_imag", 0, device)
    x = synth_waveform(batch, FS * seconds, rank=rank).to(device)
    f = vae_step(enc, dec)
    return (lambda: f(x)), {"workload": "config2: NSVAE encoder (latent_num=1, H=384) + CVAE decoder (zero skips, real_imag)",
                            "batch_per_gpu": batch, "utterance_s": seconds, "H": 384, "real_skip": False}


def config2b(device, rank=0, batch=64, seconds=4):
    enc, dec = build_vae(2, 1, "twophase", "mask", 0, device)
    x = synth_waveform(batch, FS * seconds, rank=rank).to(device)
    f = vae_step(enc, dec, pad="sig")
    return (lambda: f(x)), {"workload": "config2b (shipped final system): NSVAE encoder latent_num=2 (H=768) + "
                                        "nsvae_pvae_dccrn_decoder_twophase (real skips, mask head)",
                            "batch_per_gpu": batch, "utterance_s": seconds, "H": 768, "real_skip": True}


def config3(device, world=1, rank=0, total_batch=256, seconds=10, chunk=128):
    """Strong scaling: the 256 utterances are split over the ranks; a rank runs its shard in passes of <= ``chunk``
    utterances (128: two 64-utterance LSTM chunks interleaved in one launch; ~75 GB of activations per pass)."""
    m = M.DCCRN_(NFFT, HOP, get_net_params(True), True, device, WIN, SKIPS, "mask", False, None, None)
    m.load_state_dict(fill_state_dict(m.state_dict(), 5), strict=True)
    m = m.to(device).eval()
    per = total_batch // world
    if per * world != total_batch:
        raise ValueError("config 3 splits %d utterances over %d ranks evenly" % (total_batch, world))
    sizes = [min(chunk, per - lo) for lo in range(0, per, chunk)]
    xs = [synth_waveform(n, FS * seconds, rank=rank, seed=77 + i).to(device) for i, n in enumerate(sizes)]

    def step():
        out = None
        with torch.no_grad():
            for x in xs:
                out = m(x, train=False)[0]
        return out
    return step, {"workload": "config3: supervised DCCRN_ (causal, mask head, real skips, H=128), %d x %d s split over %d "
                              "GPU(s), passes of <= %d utterances" % (total_batch, seconds, world, chunk),
                  "batch_per_gpu": per, "global_batch": total_batch, "utterance_s": seconds, "H": 128, "real_skip": True,
                  "scaling": "strong"}


def config4(device, world=1, rank=0, group=None, batch=32, seconds=4, latent_num=2, num_samples=1):
    """End-to-end NSVAE training step (nsvae_loss.py:L598-613 with recon weights (0,0,1)):
    returns (step, info, optimiser)."""
    from . import losses
    from .train import FlatAdam
    net = get_net_params()
    noisy = M.nsvae_pvae_dccrn_encoder_twophase(net, True, device, ZDIM, NFFT, HOP, WIN, num_samples, latent_num)
    noisy.load_state_dict(fill_state_dict(noisy.state_dict(), 0))
    noisy = noisy.to(device)
    frozen = []
    for j in range(2):
        e = M.pvae_dccrn_encoder_skip_prepare(net, True, device, ZDIM, NFFT, HOP, WIN, 1)
        e.load_state_dict(fill_state_dict(e.state_dict(), 1 + j))
        frozen.append(e.to(device).eval())
    dec = M.nsvae_pvae_dccrn_decoder_twophase(net, True, device, num_samples, ZDIM, NFFT, HOP, WIN, "mask", True, SKIPS, False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), 5))
    dec = dec.to(device)
    opt = FlatAdam(list(noisy.parameters()) + list(dec.parameters()), lr=1e-3, weight_decay=1e-3, process_group=group,
                   world_size=world)
    L = int(seconds * FS)
    xs = [synth_waveform(batch, L, seed=100 * rank + j).to(device) for j in range(3)]
    # the step is GPU-bound by a small margin (host ~60 ms ahead of 168 ms of kernels): a full (generation-2) garbage
    # collection over the long-lived module / pack objects is a 60-100 ms host pause that stalls the GPU every few steps.
    # Freezing what exists now keeps later collections to the per-step garbage.
    import gc
    gc.collect()
    gc.freeze()

    def step():
        with torch.no_grad():
            rc = frozen[0](xs[1], train=False)
            rn = frozen[1](xs[2], train=False)
        r = noisy(xs[0], train=True)
        sig, _ = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
        kl, _, _ = losses.nsvae_kl_loss(r, rc, rn, ZDIM, latent_num, 1.0)
        clean = xs[1] if num_samples == 1 else xs[1].repeat_interleave(num_samples, 0)
        loss = kl + losses.si_snr_loss(clean, sig)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss
    return step, {"workload": "config4: end-to-end NSVAE training step (2 frozen CVAE encoders fwd; noisy encoder latent_num=%d "
                              "+ twophase decoder (mask, real skips) fwd/bwd; KL + SI-SNR; flat-bucket gradient all-reduce; "
                              "Adam lr 1e-3 wd 1e-3)" % latent_num,
                  "batch_per_gpu": batch, "utterance_s": seconds, "num_samples": num_samples, "scaling": "weak"}, opt


def config5(device, rank=0, streams=128, frames_per_step=1, final=False):
    """Frame streaming: returns (enhancer, chunk source, info).  The enhancer is primed and its steady-state CUDA graph
    captured; ``chunk()`` yields the next (streams, hop * k) samples."""
    from .streaming import StreamingEnhancer
    enc, dec = build_vae(2, 1, "twophase", "mask", 0, device) if final else \
        build_vae(1, 1, "skip_prepare", "real_imag", 0, device)
    se = StreamingEnhancer(enc, dec, n_streams=streams, frames_per_step=frames_per_step, device=device)
    hop, k = se.hop, frames_per_step
    x = synth_waveform(streams, hop * k * 64 + hop, rank=rank, seed=1).to(device)
    se.prime(x[:, :hop].contiguous())
    j = [0]

    def chunk():
        lo = hop + (j[0] % 64) * hop * k
        j[0] += 1
        return x[:, lo:lo + hop * k].contiguous()
    while se._graph is None:                      # eager steps until the steady-state graph is captured
        se.step(chunk())
    for _ in range(20):
        se.step(chunk())
    return se, chunk, {"workload": "config5: causal DCCRN-VAE frame streaming (%s), %d streams per GPU, %d hop(s) per step"
                                   % ("latent_num=2, real skips, mask" if final else "latent_num=1, zero skips", streams, k),
                       "streams_per_gpu": streams, "frames_per_step": k, "scaling": "weak"}


def time_streaming(se, chunk, steps):
    """(sorted per-step latencies in ms, back-to-back ms per step) of the captured step graph (CUDA events)."""
    evs = []
    for _ in range(steps):
        se.x_in.copy_(chunk())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        se._graph.replay()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    lat = sorted(a.elapsed_time(b) for a, b in evs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        se._graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return lat, e0.elapsed_time(e1) / steps


def profile_step(step, reps=2):
    """Per-entry-point device time of one step: {name: {"launches", "ms"}} (CUDA events on the launching stream)."""
    prof = {}
    lib.set_profile_hook(lambda name, ev: prof.setdefault(name, []).append(ev))
    try:
        for _ in range(reps):
            prof.clear()
            step()
            torch.cuda.synchronize()
    finally:
        lib.set_profile_hook(None)
    return {k: {"launches": len(v), "ms": sum(v)} for k, v in lib.resolve_profile(prof).items()}
