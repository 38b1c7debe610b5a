"""Per-step device time of the end-to-end training step (config 4), 12 consecutive steps.  python tools/train_steps.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import idccrn_b200  # noqa: F401
from idccrn_b200 import workloads as W, lib
step, info, opt = W.config4(torch.device("cuda", 0))
ts, ws = [], []
for i in range(12):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0, t0 = lib.LAUNCHES[0], time.perf_counter()
    e0.record(); step(); e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1)); ws.append((t1 - t0) * 1e3)
print("device ms per step:", [round(t, 1) for t in ts])
print("host ms to issue a step:", [round(t, 1) for t in ws])
print("peak GB", torch.cuda.max_memory_allocated() / 2**30, "reserved GB", torch.cuda.memory_reserved() / 2**30,
      "num_alloc_retries", torch.cuda.memory_stats().get("num_alloc_retries"), "cudaMalloc calls", torch.cuda.memory_stats().get("num_device_alloc"))
