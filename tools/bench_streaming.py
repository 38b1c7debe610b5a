"""BASELINE config 5: causal DCCRN-VAE frame streaming, 128 concurrent streams per GPU (1024 on 8 GPUs: the streams
are independent, no collective).  Reports per-step latency (p50 / p99 / max over N replays of the captured CUDA graph,
each timed with its own pair of CUDA events; device time of a step whose input is already resident) and the aggregate
audio-seconds per second, for several frames-per-step settings.  Real-time bar: one hop (6.25 ms of audio) per step
and stream.  Writes gpurun_out/streaming.json.

    python tools/bench_streaming.py [--streams 128] [--steps 400] [--final]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as C
from idccrn_b200 import lib
from idccrn_b200.streaming import StreamingEnhancer

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=128)
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--final", action="store_true", help="final system: latent_num=2 (H=768), real skips, mask head")
args = ap.parse_args()

if args.final:
    enc, dec = C.build_vae(2, 1, "twophase", "mask", 0, "cuda")
    workload = "nsvae encoder latent_num=2 (H=768) + twophase decoder (real skips, mask head)"
else:
    enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", 0, "cuda")
    workload = "nsvae encoder latent_num=1 (H=384) + CVAE decoder (zero skips, real_imag)"
res = {"workload": workload, "streams": args.streams, "fs": 16000, "hop": 100, "cases": []}
for k in (1, 2, 4, 8):
    se = StreamingEnhancer(enc, dec, n_streams=args.streams, frames_per_step=k)
    hop = se.hop
    x = C.synth_waveform(args.streams, hop * k * 64 + hop, seed=1).cuda()
    se.prime(x[:, :hop].contiguous())
    j = 0

    def chunk():
        global j
        lo = hop + (j % 64) * hop * k
        j += 1
        return x[:, lo:lo + hop * k].contiguous()
    while se._graph is None:                      # eager warm-up steps until the steady-state graph is captured
        se.step(chunk())
    for _ in range(20):
        se.step(chunk())
    torch.cuda.synchronize()
    n0 = lib.LAUNCHES[0]
    se._step_impl(1 << 40, 1 << 20)               # one eager step to count our kernels per step
    launches = lib.LAUNCHES[0] - n0
    se.steps += 1
    torch.cuda.synchronize()
    evs = []
    for _ in range(args.steps):
        se.x_in.copy_(chunk())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        se._graph.replay()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    # back-to-back replays (throughput; no host gaps between steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        se._graph.replay()
    e1.record()
    torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / args.steps
    audio_per_step = args.streams * hop * k / 16000.0
    case = {"frames_per_step": k, "kernels_per_step": launches, "latency_ms_p50": ms[len(ms) // 2],
            "latency_ms_p99": ms[int(len(ms) * 0.99)], "latency_ms_max": ms[-1], "ms_per_step_back_to_back": b2b,
            "audio_s_per_s": audio_per_step / (b2b / 1e3), "realtime_factor_per_stream": (hop * k / 16000.0) / (b2b / 1e3),
            "algorithmic_delay_ms": (se.output_delay + hop * (k - 1)) / 16.0}
    assert torch.isfinite(se.y_out).all()
    res["cases"].append(case)
    print(case, flush=True)
    del se
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
name = "streaming_final.json" if args.final else "streaming.json"
json.dump(res, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
