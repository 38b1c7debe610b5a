"""Device time of the two-layer ComplexLSTM recurrence for small batches: cluster kernel (DSMEM exchange, idv_lstm2_cluster_tc)
against the wavefront kernel (idv_lstm2_wave_tc), inside an encoder forward of 4-s utterances (T = 641)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as C
from idccrn_b200 import lib, ops

out = {"cases": []}
for model, H in (("vae", 384), ("dccrn", 128)):
    if model == "vae":
        enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", 0, "cuda")
        run = lambda x: enc(x, train=False)
    else:
        import idccrn_b200 as M
        from idccrn_b200.synth import fill_state_dict
        net = M.DCCRN_(C.NFFT, C.HOP, M.get_net_params(True), True, "cuda", C.WIN, C.SKIPS, "mask", False, None, None)
        net.load_state_dict(fill_state_dict(net.state_dict(), 3), strict=True)
        net = net.cuda().eval()
        run = lambda x: net(x)
    for NB in (4, 8, 16, 32, 48, 64, 96, 128):
        x = C.synth_waveform(NB, 64000, seed=1).cuda()
        row = {"model": model, "H": H, "NB": NB}
        for name, flag in (("cluster", True), ("wave", False)):
            ops.LSTM_CLUSTER[0] = flag
            if flag and ops.lstm2_cluster_supported(H, NB, 641, x.device) is None:
                # force the chunked cluster launch even where the heuristic prefers the wavefront kernel
                _sup = ops.lstm2_cluster_supported
                ops.lstm2_cluster_supported = lambda H_, NB_, T_, d_: lib.lstm2_cluster_config(H_, NB_, T_)
            else:
                _sup = None
            with torch.no_grad():
                for _ in range(3):
                    run(x)
                torch.cuda.synchronize()
                prof = []
                lib.set_profile_hook(lambda n, ev: prof.append((n, ev)))
                for _ in range(5):
                    run(x)
                torch.cuda.synchronize()
                lib.set_profile_hook(None)
            ms = [e0.elapsed_time(e1) for n, (e0, e1) in prof if n in ("idv_lstm2_cluster_tc", "idv_lstm2_wave_tc")]
            names = sorted({n for n, _ in prof if "lstm2" in n})
            row[name + "_ms"] = round(sum(ms) / max(len(ms), 1), 4)
            row[name + "_kernel"] = names
            if _sup is not None:
                ops.lstm2_cluster_supported = _sup
                row["cluster_forced"] = True
        ops.LSTM_CLUSTER[0] = True
        out["cases"].append(row)
        print(row, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "lstm_small.json"), "w"), indent=1)
