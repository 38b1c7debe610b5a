"""Per-entry-point device time of one training step (CUDA events around every C-ABI call; the BPTT CUDA graph is
switched off so that its launches are visible).  python tools/train_profile.py [batch] [latent_num] [phase 1|3]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["IDV_BPTT_GRAPH"] = "0"
import torch
import common as C
import idccrn_b200 as M
from idccrn_b200 import lib, losses
from idccrn_b200.synth import fill_state_dict, synth_waveform
lib.set_option("gemm_cta_pairs", int(os.environ.get("IDV_PAIRS", "1")))      # A/B of the CTA-pair tap-GEMM
B, L, ln = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 64000, int(sys.argv[2]) if len(sys.argv) > 2 else 2
phase = int(sys.argv[3]) if len(sys.argv) > 3 else 1
net = M.get_net_params()
noisy = M.nsvae_pvae_dccrn_encoder_twophase(net, True, "cuda", 128, 512, 100, 400, 1, ln)
noisy.load_state_dict(fill_state_dict(noisy.state_dict(), 0)); noisy = noisy.cuda()
tgt = M.pvae_dccrn_encoder_skip_prepare(net, True, "cuda", 128, 512, 100, 400, 1)
tgt.load_state_dict(fill_state_dict(tgt.state_dict(), 1)); tgt = tgt.cuda().eval()
dec = None
if phase == 3:
    dec = M.nsvae_pvae_dccrn_decoder_twophase(net, True, "cuda", 1, 128, 512, 100, 400, "mask", True, C.SKIPS, False)
    dec.load_state_dict(fill_state_dict(dec.state_dict(), 5)); dec = dec.cuda()
x = synth_waveform(B, L, seed=0).cuda()
with torch.no_grad():
    rc = tgt(x, train=False)
def step():
    r = noisy(x, train=True)
    loss, _, _ = losses.nsvae_kl_loss(r, rc, rc, 128, ln, 1.0)
    if dec is not None:
        sig, pred = dec(r[11], r[0], r[8], r[9], r[10], train=True, pad="sig")
        loss = loss + losses.si_snr_loss(x, sig)
        for p in dec.parameters(): p.grad = None
    for p in noisy.parameters(): p.grad = None
    loss.backward()
step(); torch.cuda.synchronize()
prof = []
shapes = []
_orig_call = lib.call
def _spy(name, *a, **kw):
    if name == "idv_tapgemm_tc":
        shapes.append((a[6], a[12], a[15], a[9]))          # rows, N, units, kc_max
    return _orig_call(name, *a, **kw)
lib.call = _spy
import idccrn_b200.ops as _ops, idccrn_b200.train as _tr
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
tg = [(a.elapsed_time(b), sh) for (name, (a, b)), sh in zip([p for p in prof if p[0] == "idv_tapgemm_tc"], shapes)]
small = [t for t, sh in tg if sh[0] <= 4 * B]
print("tap-GEMMs with <= %d rows (BPTT steps): %d calls %.2f ms" % (4 * B, len(small), sum(small)))
for t, sh in sorted([x for x in tg if x[1][0] > 4 * B], key=lambda x: -x[0])[:40]:
    print("  %8.3f ms  rows %7d N %4d units %4d kc_max %d" % ((t,) + sh))
