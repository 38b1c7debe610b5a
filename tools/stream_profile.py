"""Per-kernel device time of one eager streaming step (CUDA events around every C-ABI call)."""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as C
from idccrn_b200 import lib
from idccrn_b200.streaming import StreamingEnhancer
k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", 0, "cuda")
se = StreamingEnhancer(enc, dec, n_streams=128, frames_per_step=k, use_graph=False)
x = C.synth_waveform(128, 100 * k * 40 + 100, seed=1).cuda()
se.prime(x[:, :100].contiguous())
for j in range(12):
    se.step(x[:, 100 + j * 100 * k:100 + (j + 1) * 100 * k].contiguous())
torch.cuda.synchronize()
prof = []
lib.set_profile_hook(lambda name, ev: prof.append((name, ev)))
for j in range(12, 17):
    se.step(x[:, 100 + j * 100 * k:100 + (j + 1) * 100 * k].contiguous())
torch.cuda.synchronize()
lib.set_profile_hook(None)
n = len(prof) // 5
last = prof[-n:]
tot = 0
for i, (name, (e0, e1)) in enumerate(last):
    ms = e0.elapsed_time(e1)
    tot += ms
    print("%2d %-28s %.1f us" % (i, name, ms * 1e3))
print("sum %.1f us" % (tot * 1e3))
