"""Where the end-to-end training step spends the time that is NOT in the package's kernels: torch.profiler over one step
(BPTT graphs on, as in the bench): CUDA time and call counts of ATen kernels (weight re-packing, gradient unfolding,
reductions), CPU time of the Python side.  python tools/train_host_profile.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import idccrn_b200  # noqa: F401
from idccrn_b200 import workloads as W
from torch.profiler import profile, ProfilerActivity

step, info, opt = W.config4(torch.device("cuda", 0))
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print("step %.1f ms" % e0.elapsed_time(e1))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted(ka, key=lambda k: -k.device_time_total)[:40]
print("%-60s %8s %12s %12s" % ("name", "calls", "cuda ms", "cpu ms"))
for k in rows:
    print("%-60s %8d %12.3f %12.3f" % (k.key[:60], k.count, k.device_time_total / 1e3, k.cpu_time_total / 1e3))
tot_cuda = sum(k.self_device_time_total for k in ka) / 1e3
aten_cuda = sum(k.self_device_time_total for k in ka if k.key.startswith("aten::") or "at::" in k.key or "elementwise" in k.key) / 1e3
print("total self CUDA %.1f ms; ATen-side self CUDA %.1f ms; kernels launched %d" % (tot_cuda, aten_cuda, sum(k.count for k in ka if k.device_time_total > 0)))
