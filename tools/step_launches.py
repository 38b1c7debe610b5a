"""One step of a BASELINE configuration between cudaProfilerStart / Stop, after warm-up: run under
    ncu --profile-from-start off --metrics gpu__time_duration.sum,... --csv --log-file gpurun_out/x.csv python tools/step_launches.py [config]
to get the launch list of exactly one step (config: 1, 2 (default), 2b, 3)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import idccrn_b200  # noqa: F401
from idccrn_b200 import workloads as W

cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
maker = {"1": W.config1, "2": W.config2, "2b": W.config2b, "3": lambda d: W.config3(d, total_batch=128)}[cfg]
step, info = maker(torch.device("cuda", 0))
for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(info["workload"])
