"""Every host-synchronising torch call of one end-to-end training step, with the Python line that made it
(torch.cuda.set_sync_debug_mode('warn')).  python tools/train_sync_debug.py"""
import collections, os, sys, traceback, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import idccrn_b200  # noqa: F401
from idccrn_b200 import workloads as W

step, info, opt = W.config4(torch.device("cuda", 0))
for _ in range(3):
    step()
torch.cuda.synchronize()
sites = collections.Counter()


def showwarning(message, category, filename, lineno, file=None, line=None):
    if "synchroniz" not in str(message):
        return
    for fr in reversed(traceback.extract_stack()[:-2]):
        if "i-dccrn-vae_b200" in fr.filename or fr.filename.endswith("workloads.py"):
            sites["%s:%d %s" % (os.path.basename(fr.filename), fr.lineno, fr.line)] += 1
            break
    else:
        sites["(outside the package)"] += 1


warnings.showwarning = showwarning
warnings.simplefilter("always")
torch.cuda.set_sync_debug_mode("warn")
step()
torch.cuda.set_sync_debug_mode("default")
torch.cuda.synchronize()
for k, v in sites.most_common():
    print("%4d  %s" % (v, k))
print("total synchronising calls in one step:", sum(sites.values()))
