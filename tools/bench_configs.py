"""Throughput of the BASELINE configs that fit one GPU (device-resident inputs, CUDA events, 3 warm-up + 5 timed).
Writes gpurun_out/configs.json.  Extra datapoints for DESIGN.md — bench.py stays the contract line (config 2)."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as C
import idccrn_b200 as M
from idccrn_b200.synth import fill_state_dict, synth_waveform


def time_it(fn, warm=3, steps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


out = {}


def vae(tag, B, L, latent_num, dec_kind, recon):
    enc, dec = C.build_vae(latent_num, 1, dec_kind, recon, 0, "cuda")
    x, _ = C.vae_inputs(B, L, 1, latent_num, 0, "cuda")

    def step():
        with torch.no_grad():
            r = enc(x, train=False)
            if dec_kind == "skip_prepare":
                dec(r[11], r[0], r[8], r[9], r[10], train=False)
            else:
                dec(r[11], r[0], r[8], r[9], r[10], train=False, pad="sig")
    ms = time_it(step)
    out[tag] = {"B": B, "seconds": L / 16000, "ms_per_step": ms, "audio_s_per_s": B * L / 16000 / (ms / 1e3)}
    print(tag, out[tag], flush=True)
    del enc, dec
    torch.cuda.empty_cache()


def dccrn(tag, B, L):
    m = M.DCCRN_(512, 100, M.get_net_params(), True, "cuda", 400, list(range(6)), "mask", False, None, None)
    m.load_state_dict(fill_state_dict(m.state_dict(), 5))
    m = m.cuda().eval()
    x = synth_waveform(B, L).cuda()

    def step():
        with torch.no_grad():
            m(x, train=False)
    ms = time_it(step)
    out[tag] = {"B": B, "seconds": L / 16000, "ms_per_step": ms, "audio_s_per_s": B * L / 16000 / (ms / 1e3)}
    print(tag, out[tag], flush=True)
    del m
    torch.cuda.empty_cache()


vae("config1_cvae_recon_B4x4s", 4, 64000, 1, "skip_prepare", "real_imag")
vae("config2_nsvae_cvae_B64x4s", 64, 64000, 1, "skip_prepare", "real_imag")
vae("config2b_finetuned_mask_latent2_B64x4s", 64, 64000, 2, "twophase", "mask")
dccrn("config3_dccrn_mask_B64x10s_(one of 4 shards of 256)", 64, 160000)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
