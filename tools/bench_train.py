"""Phase-1 NSVAE training step (train_nsvae.py:L472-566) on B200: frozen clean + noise CVAE encoders (train=False),
noisy encoder train=True, closed-form KL loss, backward, gradient all-reduce (NCCL, when launched under torchrun), Adam
(lr 1e-3, weight_decay 1e-3).  Per-GPU batch 32 x 4-s synthetic utterances (BASELINE config 4's shape).  Device time per
step with CUDA events (max over ranks), phases timed with events as well.  Writes gpurun_out/train_step[_N].json.

    python tools/bench_train.py [--batch 32] [--seconds 4] [--latent-num 2] [--steps 3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import common as C
import idccrn_b200 as M
from idccrn_b200 import lib, losses
from idccrn_b200.synth import fill_state_dict, synth_waveform
from idccrn_b200.train import FlatAdam

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("--latent-num", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--phase", type=int, default=1, help="1: KL step of train_nsvae.py; 2: decoder step of "
                "train_second_phase_decoder.py (frozen NSVAE encoder, decoder train=True, SI-SNR); 3: end-to-end step of "
                "BASELINE config 4 (KL + SI-SNR, gradients into the encoder AND the decoder)")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
group = None
if world > 1:
    dist.init_process_group("nccl")
    group = dist.group.WORLD
dev = "cuda"
B, L, ln = args.batch, int(args.seconds * 16000), args.latent_num
net = M.get_net_params()
noisy = M.nsvae_pvae_dccrn_encoder_twophase(net, True, dev, C.ZDIM, C.NFFT, C.HOP, C.WIN, 1, ln)
noisy.load_state_dict(fill_state_dict(noisy.state_dict(), 0))
noisy = noisy.to(dev)
frozen = []
for j in range(2):
    e = M.pvae_dccrn_encoder_skip_prepare(net, True, dev, C.ZDIM, C.NFFT, C.HOP, C.WIN, 1)
    e.load_state_dict(fill_state_dict(e.state_dict(), 1 + j))
    frozen.append(e.to(dev).eval())
dec = None
if args.phase in (2, 3):
    dec = M.nsvae_pvae_dccrn_decoder_twophase(net, True, dev, 1, C.ZDIM, C.NFFT, C.HOP, C.WIN, "mask", True, C.SKIPS, False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), 5))
    dec = dec.to(dev)
    if args.phase == 2:
        noisy.eval()
train_params = {1: lambda: list(noisy.parameters()), 2: lambda: list(dec.parameters()),
                3: lambda: list(noisy.parameters()) + list(dec.parameters())}[args.phase]()
opt = FlatAdam(train_params, lr=1e-3, weight_decay=1e-3, process_group=group,
               world_size=world)
xs = [synth_waveform(B, L, seed=100 * rank + j).to(dev) for j in range(3)]
ev = lambda: torch.cuda.Event(enable_timing=True)


def step2(timers=None):
    """train_second_phase_decoder.py:L376-433 with the shipped recon_loss_weight '001' (SI-SNR only)."""
    marks = [ev() for _ in range(5)]
    marks[0].record()
    with torch.no_grad():
        r = noisy(xs[0], train=False)
    marks[1].record()
    sig, pred = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
    loss = losses.si_snr_loss(xs[1], sig)
    marks[2].record()
    opt.zero_grad()
    loss.backward()
    marks[3].record()
    opt.step()
    marks[4].record()
    if timers is not None:
        timers.append(marks)
    return loss


def step3(timers=None):
    """nsvae_loss_with_cvae_decoder_recon.kl_loss_and_recon_loss (model/nsvae_loss.py:L598-613), recon weights (0,0,1)."""
    marks = [ev() for _ in range(5)]
    marks[0].record()
    with torch.no_grad():
        rc = frozen[0](xs[1], train=False)
        rn = frozen[1](xs[2], train=False)
    marks[1].record()
    r = noisy(xs[0], train=True)
    sig, pred = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
    kl, kc, kn = losses.nsvae_kl_loss(r, rc, rn, C.ZDIM, ln, 1.0)
    loss = kl + losses.si_snr_loss(xs[1], sig)
    marks[2].record()
    opt.zero_grad()
    loss.backward()
    marks[3].record()
    opt.step()
    marks[4].record()
    if timers is not None:
        timers.append(marks)
    return loss


def step(timers=None):
    if args.phase == 2:
        return step2(timers)
    if args.phase == 3:
        return step3(timers)
    marks = [ev() for _ in range(5)]
    marks[0].record()
    with torch.no_grad():
        rc = frozen[0](xs[1], train=False)
        rn = frozen[1](xs[2], train=False)
    marks[1].record()
    r = noisy(xs[0], train=True)
    loss, kc, kn = losses.nsvae_kl_loss(r, rc, rn, C.ZDIM, ln, 1.0)
    marks[2].record()
    opt.zero_grad()
    loss.backward()
    marks[3].record()
    opt.step()
    marks[4].record()
    if timers is not None:
        timers.append(marks)
    return loss


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
n0 = lib.LAUNCHES[0]
timers = []
e0, e1 = ev(), ev()
e0.record()
for _ in range(args.steps):
    loss = step(timers)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms)
if rank == 0:
    ph = [sum(t[i].elapsed_time(t[i + 1]) for t in timers) / len(timers) for i in range(4)]
    wl = {1: "phase-1 NSVAE training step: 2 frozen CVAE encoders fwd + noisy encoder (latent_num=%d) fwd/bwd + KL + Adam",
          2: "phase-2 decoder training step: frozen NSVAE encoder (latent_num=%d) fwd + twophase decoder (mask, real skips) "
             "train fwd + SI-SNR + bwd + Adam",
          3: "end-to-end NSVAE training step (BASELINE config 4): 2 frozen CVAE encoders fwd + noisy encoder (latent_num=%d) "
             "and twophase decoder (mask, real skips) fwd/bwd + KL + SI-SNR + Adam"}[args.phase] % ln
    names = {1: ("frozen_encoders_fwd", "noisy_fwd_and_loss"), 2: ("frozen_encoder_fwd", "decoder_fwd_and_loss"),
             3: ("frozen_encoders_fwd", "encoder_decoder_fwd_and_loss")}[args.phase]
    out = {"workload": wl, "batch_per_gpu": B, "seconds": args.seconds, "n_gpus": world, "ms_per_step": ms,
           "audio_s_per_s": world * B * args.seconds / (ms / 1e3), "loss": float(loss),
           "phase_ms": {names[0]: ph[0], names[1]: ph[1], "backward": ph[2],
                        "allreduce_and_adam": ph[3]},
           "kernel_launches_per_step": (lib.LAUNCHES[0] - n0) / args.steps,
           "grad_bytes_allreduced": int(opt.gflat.numel() * 4), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "train_step%s%s.json" % ({1: "", 2: "_phase2", 3: "_e2e"}[args.phase], "" if world == 1 else "_%d" % world)), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
