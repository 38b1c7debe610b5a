#!/usr/bin/env python
"""bench.py — audio-seconds enhanced per second on the BASELINE config-2 workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): NSVAE encoder (nsvae_pvae_dccrn_encoder_twophase, latent_num=1,
H=384) + pretrained-CVAE decoder (pvae_dccrn_decoder_skip_prepare, zero skips, real_imag), batch 64 x 4 s
synthetic 16 kHz utterances PER GPU (weak scaling: utterances are independent, no collective on the data
path), random-init weights, on-device Philox eps.  One "step" = one forward of the whole path over one batch.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput, `e2e` = same metric through the public
module API with pinned-host input and output copies inside the timed region.  `--impl reference` times the
CPU oracle port of the reference (oracle/ref_port.py — the reference is pure PyTorch, so its CPU path is the
same library calls) on a bounded sample with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS, SECONDS, HOP, NFFT, WIN, ZDIM = 16000, 4, 100, 512, 400, 128
BATCH_PER_GPU = 64
CPU_SAMPLE_BATCH = 4


def algorithmic_gmac_per_utt(T, H=384, real_skip=False):
    from idccrn_b200.workloads import algorithmic_gmac_per_utt as f
    return f(T, H, real_skip)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_throughput(steps, warmup, batch=CPU_SAMPLE_BATCH):
    """The reference's CPU path (oracle port, same torch library calls) on batch x 4 s, all host threads."""
    from oracle import ref_port as P
    import idccrn_b200 as M
    from idccrn_b200.synth import fill_state_dict, synth_eps, synth_waveform
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = M.get_net_params()
    enc = M.nsvae_pvae_dccrn_encoder_twophase(net, True, "cpu", ZDIM, NFFT, HOP, WIN, 1, 1)
    dec = M.pvae_dccrn_decoder_skip_prepare(net, True, "cpu", 1, ZDIM, NFFT, HOP, WIN, "real_imag", list(range(6)))
    esd, dsd = fill_state_dict(enc.state_dict(), 0), fill_state_dict(dec.state_dict(), 1)
    L = FS * SECONDS
    x = synth_waveform(batch, L)
    eps = synth_eps((batch, 1, L // HOP + 1, ZDIM))

    def step():
        with torch.no_grad():
            st = P.vae_encoder_forward(esd, x, ZDIM, 1, 1, eps)
            return P.vae_decoder_forward(dsd, st["stft_x"], st["z_speech"], st["skiper"], st["C"], st["F"], 1,
                                         "real_imag", "zero")["recon_sig"]
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch * SECONDS / dt, dt, cores, "%d x %d s utterances per step, %d timed steps" % (batch, SECONDS, steps)


def workload_config(B, world, T, g):
    return {"workload": "config2: NSVAE encoder (nsvae_pvae_dccrn_encoder_twophase, latent_num=1, H=384) "
                        "+ CVAE decoder (pvae_dccrn_decoder_skip_prepare, zero skips, real_imag)",
            "batch_per_gpu": B, "global_batch": B * world, "utterance_s": SECONDS, "fs": FS, "frames": T,
            "eps": "on-device Philox", "parallelism": "replica x%d (utterance shards)" % world,
            "l2": "per-step activation working set ~12 GB >> 126 MB L2 (no flush needed)",
            "gflop_per_utt_algorithmic": 2 * g["total"],
            # identical in both arms: the CPU legs (this arm's cpu_baseline and the whole --impl reference arm) time
            # cpu_sample_batch utterances per step, not batch_per_gpu (CPU throughput is flat in the batch size)
            "cpu_sample_batch": CPU_SAMPLE_BATCH,
            "reference_arm": "oracle/ref_port.py (the reference's own torch ops) on the host cores, "
                             "%d x %d s utterances per step" % (CPU_SAMPLE_BATCH, SECONDS)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    val, dt, cores, sample = cpu_oracle_throughput(steps, warmup)
    print(json.dumps({
        "impl": "reference", "metric": "audio_seconds_enhanced_per_second", "value": val, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, args.gpus, FS * SECONDS // HOP + 1,
                                  algorithmic_gmac_per_utt(FS * SECONDS // HOP + 1)),
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch.distributed as dist
    import idccrn_b200 as M
    from idccrn_b200 import lib, shard
    from idccrn_b200.synth import fill_state_dict, synth_waveform

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1 (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, L = args.batch, FS * SECONDS
    T = L // HOP + 1
    net = M.get_net_params()
    enc = M.nsvae_pvae_dccrn_encoder_twophase(net, True, dev, ZDIM, NFFT, HOP, WIN, 1, 1)
    dec = M.pvae_dccrn_decoder_skip_prepare(net, True, dev, 1, ZDIM, NFFT, HOP, WIN, "real_imag", list(range(6)))
    enc.load_state_dict(fill_state_dict(enc.state_dict(), 0))
    dec.load_state_dict(fill_state_dict(dec.state_dict(), 1))
    enc, dec = enc.to(dev).eval(), dec.to(dev).eval()
    x_host = synth_waveform(B, L, rank=rank).pin_memory()
    x_dev = x_host.to(dev)

    def step(x):
        with torch.no_grad():
            r = enc(x, train=False)
            sig, _ = dec(r[11], r[0], r[8], r[9], r[10], train=False)
        return sig

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    from idccrn_b200.pipeline import HostPipeline, StreamPipeline

    def timed(fn, steps, n_streams=1, finish=None):
        """K steps bracketed by barrier + synchronize; device time by CUDA events on the current stream (the
        pipeline streams fork after e0 and join before e1)."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with StreamPipeline(dev, n_streams) as pipe:
            for _ in range(2 if n_streams > 1 else 0):           # packs for the pipelined kernel configuration
                with pipe.next_stream():
                    fn()
                torch.cuda.synchronize()
            pipe.join()
            torch.cuda.synchronize()
            e0.record()
            for s_ in pipe.streams:
                s_.wait_stream(torch.cuda.current_stream())      # the timed region starts after e0
            for _ in range(steps):
                with pipe.next_stream():
                    fn()
            pipe.join()
            if finish is not None:
                finish()
            e1.record()
        torch.cuda.synchronize()
        ms = shard.max_over_ranks(e0.elapsed_time(e1), dev)       # device time, MAX over ranks
        sync_all()
        return ms

    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    # ---- device-resident throughput (the `value`)
    lib.LAUNCHES[0] = 0
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(lambda: step(x_dev), args.steps, args.streams)
    launches = lib.LAUNCHES[0]
    ms_single = timed(lambda: step(x_dev), args.steps, 1) if args.streams > 1 else ms
    clk = clocks.stop() if rank == 0 else None
    value = world * B * SECONDS * args.steps / (ms / 1e3)

    # ---- end-to-end through the public API with host buffers
    out_hosts = [torch.empty((B, L), dtype=torch.float32).pin_memory() for _ in range(max(2, args.streams))]
    e2e_i = [0]

    host_pipe = HostPipeline(step, dev) if args.streams == 1 else None

    def e2e_step():
        # public API with host buffers: pinned H2D of the batch, forward, D2H of the enhanced waveforms, every step.
        # One stream of kernels: pipeline.HostPipeline runs the copies of the neighbouring batches on two copy streams
        # beside the compute; several streams: the copies of one batch overlap the compute of the other.  The results
        # are complete when the timed region ends (the pipelines join before the closing event, then the device is
        # synchronised).
        out = out_hosts[e2e_i[0] % len(out_hosts)]
        e2e_i[0] += 1
        if host_pipe is not None:
            host_pipe.submit(x_host, out)
            return
        xd = x_host.to(dev, non_blocking=True)
        sig = step(xd)
        out.copy_(sig, non_blocking=True)
    e2e_step()
    if host_pipe is not None:
        host_pipe.join()
    torch.cuda.synchronize()
    # (the latent noise is drawn on the device, so two forwards differ: the copied result is checked for sanity only)
    assert bool(torch.isfinite(out_hosts[0]).all()) and float(out_hosts[0].abs().max()) > 0, "e2e result did not arrive"
    e2e_steps = max(1, args.steps)
    ms_e2e = timed(e2e_step, e2e_steps, args.streams, finish=host_pipe.join if host_pipe is not None else None)
    e2e_val = world * B * SECONDS * e2e_steps / (ms_e2e / 1e3)

    # ---- per-kernel device time of one step (CUDA events on the launching stream), for the roofline
    prof = {}

    def hook(name, t_ms):
        prof.setdefault(name, []).append(t_ms)
    lib.set_profile_hook(hook)
    for _ in range(2):
        prof.clear()
        step(x_dev)
        torch.cuda.synchronize()
    lib.set_profile_hook(None)
    per_kernel = {k: {"launches": len(v), "ms": sum(v)} for k, v in lib.resolve_profile(prof).items()}
    step_ms_prof = sum(v["ms"] for v in per_kernel.values())

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    # ---- the other BASELINE configurations (all ranks take part; config 3 is split over them)
    enc = dec = None                 # (the step closure holds the only other reference)
    torch.cuda.empty_cache()
    which = [c for c in args.configs.split(",") if c]
    configs = measure_configs(args, dev, world, rank, dist.group.WORLD if world > 1 else None, peaks, which) if which else {}
    eager = None
    if rank == 0 and world == 1 and not args.no_eager:
        try:
            eager = eager_b200()
        except Exception as e:
            eager = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        torch.cuda.empty_cache()

    if rank == 0:
        g = algorithmic_gmac_per_utt(T)
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
        tc_mode = "idv_tapgemm_tc" in per_kernel
        # the dominant kernel's launches of one step: idv_tapgemm_tc (convs, transposed convs, LSTM in-proj, iSTFT DFT,
        # and the first encoder layer when it runs on the tensor cores) + idv_tapgemm_tc_b2 (dense composed with the first
        # decoder layer).  Algorithmic FLOPs = the reference's MAC counts of exactly those layers: the
        # composed launch is credited with the dense + first-decoder-layer MACs the reference executes.
        names = ["idv_tapgemm_tc", "idv_tapgemm_tc_b2"] if tc_mode else ["idv_tapgemm_f32"]
        tg = {"ms": sum(per_kernel[n]["ms"] for n in names if n in per_kernel),
              "launches": sum(per_kernel[n]["launches"] for n in names if n in per_kernel)}
        if not tg["launches"]:
            tg["ms"] = float("nan")
        gmac = g["tapgemm"] + (g["enc"][0] if "idv_enc0_fwd" not in per_kernel else 0.0)
        tg_flops = 2 * gmac * 1e9 * B
        achieved_tf = tg_flops / (tg["ms"] / 1e3) / 1e12 if tg["launches"] else None
        cpu_val, cpu_dt, cores, sample = cpu_oracle_throughput(2, 1) if not args.no_cpu else (None, None, 0, "skipped")
        traffic, traffic_src = ncu_dram_traffic(B, tc_mode)
        hbm_stages = hbm_stage_rooflines(per_kernel, B, L, T, peaks) if tc_mode else None
        line = {
            "metric": "audio_seconds_enhanced_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (bf16x3 split on the tensor pipe)" if tc_mode else "f32", "data": "synthetic",
            "config": workload_config(B, world, T, g),
            "e2e": {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": B * L * 4,
                    "d2h_bytes_per_step": B * L * 4, "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": launches,
            "streams": args.streams,
            "single_stream": {"value": world * B * SECONDS * args.steps / (ms_single / 1e3), "ms_per_step": ms_single / args.steps},
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "tapgemm_tc_kernel (complex conv / convT / LSTM layer-0 in-proj / dense+dec0 / iSTFT DFT)",
                         "kernel_launches_per_step": tg["launches"],
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (achieved_tf / peak_tf) if achieved_tf else None, "traffic": traffic,
                         "traffic_unit": "GB per step (dram__bytes_read.sum + dram__bytes_write.sum over the kernel's launches)",
                         "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_gflop_per_step": tg_flops / 1e9,
                         "kernel_ms_per_step": tg["ms"], "kernel_share_of_step": tg["ms"] / step_ms_prof if step_ms_prof else None,
                         "split_factor": 3 if tc_mode else None,
                         "tensor_pipe_frac_incl_split": (3 * achieved_tf / peak_tf) if (tc_mode and achieved_tf) else None,
                         "note": ("tcgen05 kind::f16, error-compensated bf16 split: 3 MMAs per algorithmic product; "
                                  "`achieved` counts ALGORITHMIC flops (SURVEY 8(d)), the tensor pipe executes 3x that")
                         if tc_mode else "fp32 SIMT implementation measured against the bf16 tensor-pipe peak"},
            "per_kernel_ms": per_kernel,
            "hbm_stages": hbm_stages,
            "configs": configs,
            "eager_b200": eager,
            "cpu_baseline": {"value": cpu_val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_configs(args, dev, world, rank, group, peaks, which):
    """The other BASELINE configurations on the same ranks (``configs`` object of the JSON line): device-resident
    inputs, >= 3 warm-up steps, CUDA events bracketed by barrier + synchronize, MAX over ranks.  Each entry:
    ms_per_step, whole-job audio-s/s and the algorithmic-FLOP fraction of the measured sustained bf16 peak
    (SURVEY 8(d) MAC counts; the tensor pipe executes 3x that with the bf16x3 split)."""
    import gc
    import torch.distributed as dist
    from idccrn_b200 import lib, shard, workloads as W
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    out = {}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step, steps, warm=3):
        for _ in range(warm):
            step()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.LAUNCHES[0]
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = shard.max_over_ranks(e0.elapsed_time(e1), dev) / steps
        sync_all()
        return ms, (lib.LAUNCHES[0] - n0) / steps

    def inference(tag, maker, **kw):
        step, info = maker(dev, **kw)
        ms, launches = timed(step, min(args.steps, 5))
        T = info["utterance_s"] * FS // HOP + 1
        g = algorithmic_gmac_per_utt(T, info["H"], info["real_skip"])
        strong = info.get("scaling") == "strong"
        utts = info["global_batch"] if strong else info["batch_per_gpu"] * world
        flops = 2 * g["total"] * 1e9 * info["batch_per_gpu"]                 # per GPU and step
        e = {"workload": info["workload"], "batch_per_gpu": info["batch_per_gpu"], "utterance_s": info["utterance_s"],
             "frames": T, "scaling": info.get("scaling", "weak"), "ms_per_step": ms,
             "audio_s_per_s": utts * info["utterance_s"] / (ms / 1e3), "gpu_launches_per_step": launches,
             "gflop_per_utt_algorithmic": 2 * g["total"],
             "algorithmic_tflops_per_gpu": flops / (ms / 1e3) / 1e12,
             "algorithmic_flop_frac_of_bf16_peak": flops / (ms / 1e3) / 1e12 / peak_tf}
        if args.config_kernels:
            pk = W.profile_step(step)
            e["per_kernel_ms"] = {k: round(v["ms"], 4) for k, v in sorted(pk.items(), key=lambda kv: -kv[1]["ms"])}
        out[tag] = e
        del step
        gc.collect()
        torch.cuda.empty_cache()

    def guarded(tag, fn):
        if tag not in which:
            return
        try:
            fn()
        except Exception as e:                      # a failing side config must not take the headline line with it
            out[tag] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            gc.collect()
            torch.cuda.empty_cache()

    guarded("1", lambda: inference("1", W.config1, rank=rank))
    guarded("2b", lambda: inference("2b", W.config2b, rank=rank))
    guarded("3", lambda: inference("3", W.config3, world=world, rank=rank))

    def training():
        step, info, opt = W.config4(dev, world=world, rank=rank, group=group)
        # (warm-up: lazy packs, the BPTT graphs are captured on the second call of a shape, allocator growth - the third
        #  step can still be 1.5-2 x the steady state, tools/train_steps.py)
        ms, launches = timed(step, min(args.steps, 3), warm=5)
        out["4"] = {"workload": info["workload"], "batch_per_gpu": info["batch_per_gpu"], "utterance_s": info["utterance_s"],
                    "scaling": "weak", "ms_per_step": ms,
                    "audio_s_per_s": world * info["batch_per_gpu"] * info["utterance_s"] / (ms / 1e3),
                    "gpu_launches_per_step": launches, "grad_bytes_allreduced": int(opt.gflat.numel() * 4),
                    "allreduce": "NCCL all_reduce of one flat fp32 bucket, mean over ranks" if world > 1 else "single rank"}
        del step, opt
        gc.collect()
        torch.cuda.empty_cache()
    guarded("4", training)

    def streaming():
        se, chunk, info = W.config5(dev, rank=rank)
        n0 = lib.LAUNCHES[0]
        se._step_impl(1 << 40, 1 << 20)               # one eager step to count our kernels per step
        launches = lib.LAUNCHES[0] - n0
        se.steps += 1
        sync_all()
        lat, b2b = W.time_streaming(se, chunk, 300)
        b2b = shard.max_over_ranks(b2b, dev)
        p50 = shard.max_over_ranks(lat[len(lat) // 2], dev)
        p99 = shard.max_over_ranks(lat[int(len(lat) * 0.99)], dev)
        hop_s = se.hop * se.k / float(FS)
        out["5"] = {"workload": info["workload"], "streams_per_gpu": se.NB, "streams_total": se.NB * world,
                    "frames_per_step": se.k, "scaling": "weak", "kernels_per_step": launches,
                    "latency_ms_p50": p50, "latency_ms_p99": p99, "ms_per_step_back_to_back": b2b,
                    "audio_s_per_s": world * se.NB * hop_s / (b2b / 1e3), "realtime_factor_per_stream": hop_s / (b2b / 1e3),
                    "algorithmic_delay_ms": se.output_delay / 16.0,
                    "note": "latencies: MAX over ranks of the per-rank p50 / p99 of 300 graph replays (CUDA events)"}
        del se
        gc.collect()
        torch.cuda.empty_cache()
    guarded("5", streaming)
    return out


def eager_b200(steps=2, batch=16):
    """Baseline leg, never the product: the reference-equivalent PyTorch modules (oracle/ref_port.py = the reference's
    own torch ops: cuDNN convs, nn.LSTM, torch.stft) run EAGERLY on this B200 for config 2, with PyTorch's default TF32
    convolution setting and in true fp32 - the practical bar of SURVEY 8(d), since the reference ships no Blackwell
    kernel.  Bounded sample: `batch` x 4 s utterances (throughput scaled to audio-s/s)."""
    from oracle import ref_port as P
    import idccrn_b200 as M
    from idccrn_b200.synth import fill_state_dict, synth_eps, synth_waveform
    net = M.get_net_params()
    enc = M.nsvae_pvae_dccrn_encoder_twophase(net, True, "cpu", ZDIM, NFFT, HOP, WIN, 1, 1)
    dec = M.pvae_dccrn_decoder_skip_prepare(net, True, "cpu", 1, ZDIM, NFFT, HOP, WIN, "real_imag", list(range(6)))
    esd = {k: v.cuda() for k, v in fill_state_dict(enc.state_dict(), 0).items()}
    dsd = {k: v.cuda() for k, v in fill_state_dict(dec.state_dict(), 1).items()}
    L = FS * SECONDS
    x = synth_waveform(batch, L).cuda()
    eps = [e.cuda() for e in synth_eps((batch, 1, L // HOP + 1, ZDIM))]

    def step():
        with torch.no_grad():
            st = P.vae_encoder_forward(esd, x, ZDIM, 1, 1, eps)
            return P.vae_decoder_forward(dsd, st["stft_x"], st["z_speech"], st["skiper"], st["C"], st["F"], 1,
                                         "real_imag", "zero")["recon_sig"]

    def run():
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"ms_per_step": ms, "audio_s_per_s": batch * SECONDS / (ms / 1e3)}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    out = {"what": "oracle port (the reference's torch ops) eager on this GPU, config 2, %d x %d s utterances per step"
                   % (batch, SECONDS), "batch": batch}
    try:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, False       # PyTorch defaults
        out["tf32_default"] = run()
        out["tf32_default"]["note"] = "TF32 convolutions miss the 1e-4 waveform gate (SURVEY F9: 9.7e-4)"
        torch.backends.cudnn.allow_tf32 = False
        out["fp32"] = run()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return out


def hbm_stage_rooflines(per_kernel, B, L, T, peaks, H=384, zdim=128, kpad_stft=448, kpad_istft=576, n_istft=512, c0=32):
    """Achieved HBM GB/s of the elementwise / framing / overlap-add stages (north_star): ALGORITHMIC bytes = the tensors
    a stage must read and write once (fp32 = 4 B, split bf16 = 2 x 2 B per value), divided by its live CUDA-event time."""
    R, Rf, nb = B * (T + 1), B * T, 257
    f1 = (nb + 4 - 5) // 2 + 1
    by = {
        "idv_stft_frames_split": B * L * 4 + Rf * kpad_stft * 4,                    # waveform -> split-bf16 frames
        "idv_enc0_fwd": B * nb * T * 2 * 4 + f1 * R * 2 * c0 * 4,                   # STFT (B,257,T,2) -> 129 planes x 64 ch
        "idv_lstm_combine_fwd": 4 * R * H * 4 + B * T * H * 2 * 4,                  # 4 streams -> latent (B,T,H,2)
        # fused latent stage: 4 LSTM streams -> latent (B,T,H,2) + z (B,T,zdim,2) + split-bf16 z planes
        "idv_latent_fwd": 4 * R * H * 4 + B * T * H * 2 * 4 + B * T * zdim * 2 * 4 + R * 2 * zdim * 4,
        "idv_reparam_fwd": B * T * H * 2 * 4 + B * T * zdim * 2 * 4,                # latent -> z
        "idv_z_to_planes": B * T * zdim * 2 * 4 + R * 2 * zdim * 4,                 # z -> split-bf16 rows
        "idv_spec_rows_split": B * nb * T * 2 * 4 + Rf * kpad_istft * 4,            # spectrum -> split-bf16 K-major rows
        "idv_ola_fwd": Rf * n_istft * 4 + B * L * 4,                                # frames -> waveform
    }
    peak = peaks.get("hbm_gbs") or 6548.8            # MEASURED_PEAKS.json copy bandwidth, else the recipe fallback
    out = {}
    for k, b in by.items():
        if k in per_kernel and per_kernel[k]["ms"] > 0:
            gbs = b / (per_kernel[k]["ms"] / 1e3) / 1e9
            out[k] = {"ms": per_kernel[k]["ms"], "algorithmic_mb": b / 1e6, "achieved_gbps": gbs, "frac_of_hbm_peak": gbs / peak}
    out["peak_gbps"] = peak
    return out


def ncu_dram_traffic(B, tc_mode):
    """DRAM bytes of the dominant kernel's launches in one step, from the committed `ncu --set full` capture of this
    workload (profiles/r02_ncu_tapgemm_full_final2.csv: `ncu --set full -k regex:tapgemm_tc` of ONE step of the final build
    at batch 64, tools/step_launches.py): a profiler figure, never measured here.  The capture lists every tcgen05 tap-GEMM
    launch of the step; the STFT DFT and the fused head (entry point idv_tapgemm_tc_head: the first launch and the
    N = 32 one) are not part of the kernel the roofline object describes and are left out."""
    name = "r02_ncu_tapgemm_full_final2.csv"
    path = os.path.join(ROOT, "profiles", name)
    if not tc_mode or B != 64 or not os.path.exists(path):
        return None, "no ncu capture for this configuration"
    import csv
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    body = [r for r in rows[hi + 2:] if len(r) > iw]
    sel = [r for i, r in enumerate(body) if i > 0 and "tapgemm_tc_kernel<32" not in r[ik]]     # not the STFT DFT, not the head
    gb = sum(float(r[ir].replace(",", "")) + float(r[iw].replace(",", "")) for r in sel) / 1e9
    return gb, "profiles/%s (%d launches)" % (name, len(sel))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="utterances per GPU (config 2: 64)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager_b200 baseline leg (N = 1 only)")
    ap.add_argument("--configs", default="1,2b,3,4,5",
                    help="other BASELINE configurations to add to the line's `configs` object ('' = none)")
    ap.add_argument("--config-kernels", action="store_true", help="per-kernel device times inside every `configs` entry")
    ap.add_argument("--streams", type=int, default=1,
                    help="batches in flight: consecutive steps alternate over this many CUDA streams (1 = serial)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
