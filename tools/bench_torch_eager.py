"""Context number (not a bench line): the oracle port (= the reference's stock PyTorch ops) run eagerly ON the B200
for config 2, with PyTorch's default TF32 conv setting and with TF32 disabled.  SURVEY §8(d) names this as the
practical bar since the reference ships no Blackwell kernel.  Writes gpurun_out/torch_eager.json."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common as C
from oracle import ref_port as P
from idccrn_b200.synth import synth_eps

B, L = 64, 64000
enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", 0, "cpu")
esd = {k: v.cuda() for k, v in enc.state_dict().items()}
dsd = {k: v.cuda() for k, v in dec.state_dict().items()}
x, _ = C.vae_inputs(B, L, 1, 1, 0, "cuda")
eps = [e.cuda() for e in synth_eps((B, 1, L // 100 + 1, 128))]


def step():
    with torch.no_grad():
        st = P.vae_encoder_forward(esd, x, 128, 1, 1, eps)
        return P.vae_decoder_forward(dsd, st["stft_x"], st["z_speech"], st["skiper"], st["C"], st["F"], 1,
                                     "real_imag", "zero")["recon_sig"]


def run(tag):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    return {"ms_per_step": ms, "audio_s_per_s": B * L / 16000 / (ms / 1e3)}


out = {}
# ref_port builds nn.LSTM on the CPU dtype/device of the state dict: make it follow the tensors' device
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = False
out["eager_cudnn_tf32_default"] = run("tf32")
torch.backends.cudnn.allow_tf32 = False
out["eager_fp32_no_tf32"] = run("fp32")
print(out)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "torch_eager.json"), "w"), indent=1)
