import os, sys
os.environ["IDV_LSTM_DBG"] = "1"
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import common as C
enc, dec = C.build_vae(2, 1, "twophase", "mask", 0, "cuda")
x, eps = C.vae_inputs(64, 64000, 1, 2, 0, "cuda")
with torch.no_grad():
    enc(x, train=False, eps=eps)
torch.cuda.synchronize()
