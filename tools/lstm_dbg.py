"""Print the per-phase timestamps of the tensor-core LSTM recurrence (IDV_LSTM_DBG=1) for the config-2 shape."""
import os
import sys
os.environ["IDV_LSTM_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import common as C

enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", 0, "cuda")
x, eps = C.vae_inputs(64, 64000, 1, 1, 0, "cuda")
with torch.no_grad():
    enc(x, train=False, eps=eps)
torch.cuda.synchronize()
