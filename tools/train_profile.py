"""Per-entry-point device time of one phase-1 training step (CUDA events around every C-ABI call)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as C
import idccrn_b200 as M
from idccrn_b200 import lib, losses
from idccrn_b200.synth import fill_state_dict, synth_waveform
B, L, ln = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 64000, int(sys.argv[2]) if len(sys.argv) > 2 else 2
net = M.get_net_params()
noisy = M.nsvae_pvae_dccrn_encoder_twophase(net, True, "cuda", 128, 512, 100, 400, 1, ln)
noisy.load_state_dict(fill_state_dict(noisy.state_dict(), 0)); noisy = noisy.cuda()
tgt = M.pvae_dccrn_encoder_skip_prepare(net, True, "cuda", 128, 512, 100, 400, 1)
tgt.load_state_dict(fill_state_dict(tgt.state_dict(), 1)); tgt = tgt.cuda().eval()
x = synth_waveform(B, L, seed=0).cuda()
with torch.no_grad():
    rc = tgt(x, train=False)
def step():
    r = noisy(x, train=True)
    loss, _, _ = losses.nsvae_kl_loss(r, rc, rc, 128, ln, 1.0)
    for p in noisy.parameters(): p.grad = None
    loss.backward()
step(); torch.cuda.synchronize()
prof = []
lib.set_profile_hook(lambda name, ev: prof.append((name, ev)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record()
torch.cuda.synchronize()
lib.set_profile_hook(None)
agg = collections.OrderedDict()
for name, (a, b) in prof:
    d = agg.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += a.elapsed_time(b)
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-30s %5d calls %9.2f ms" % (k, v[0], v[1]))
print("sum of kernels %.1f ms, wall (events) %.1f ms" % (tot, e0.elapsed_time(e1)))
