"""Dump a per-call timeline (start/end in ms relative to a base event, per stream) of a few pipelined steps."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as C
from idccrn_b200 import lib
from idccrn_b200.pipeline import StreamPipeline

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", 0, "cuda")
x, _ = C.vae_inputs(64, 64000, 1, 1, 0, "cuda")


def step():
    with torch.no_grad():
        r = enc(x, train=False)
        dec(r[11], r[0], r[8], r[9], r[10], train=False)


rec = []
with StreamPipeline("cuda", n_streams) as pipe:
    for _ in range(3):
        with pipe.next_stream():
            step()
        torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True)
    base.record()
    for s in pipe.streams:
        s.wait_stream(torch.cuda.current_stream())
    cur = [0]
    lib.set_profile_hook(lambda name, ev: rec.append((cur[0], name, ev[0], ev[1])))
    for i in range(steps):
        cur[0] = i
        with pipe.next_stream():
            step()
    lib.set_profile_hook(None)
    pipe.join()
torch.cuda.synchronize()
print("pipelined: %.2f ms/step over %d steps, %d streams" % (max(base.elapsed_time(b) for (_, _, _, b) in rec) / steps, steps, n_streams))
out = [(i, n, round(base.elapsed_time(a), 3), round(base.elapsed_time(b), 3)) for (i, n, a, b) in rec]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "timeline_%d.json" % n_streams), "w"))
for i in range(steps):
    ev = [e for e in out if e[0] == i]
    print("step", i, "start %.2f end %.2f" % (ev[0][2], ev[-1][3]))
    for e in ev:
        if "lstm_rec" in e[1] or "wave" in e[1] or e[1] == "idv_enc0_fwd" or "ola" in e[1]:
            print("    %-26s %8.2f -> %8.2f  (%.2f)" % (e[1], e[2], e[3], e[3] - e[2]))
