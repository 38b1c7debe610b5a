#!/bin/bash
# Round-2 GPU session N (1 GPU): small-batch cluster LSTM (DSMEM exchange): unit test, in-kernel timeline, config 1 with / without.
mkdir -p gpurun_out
for epw in 4; do
IDV_CL_EPW=$epw timeout 600 python -m pytest tests/test_gpu_abi_units.py -m gpu -q -k "lstm2_cluster" -x 2>&1 | tail -3
IDV_CL_EPW=$epw timeout 300 python - > gpurun_out/r02_cluster_dbg_epw$epw.log 2>&1 <<PY
import os, sys
os.environ["IDV_LSTM_DBG"] = "1"
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import common as C
enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", 0, "cuda")
x, eps = C.vae_inputs(4, 64000, 1, 1, 0, "cuda")
with torch.no_grad():
    enc(x, train=False, eps=eps)
torch.cuda.synchronize()
PY
echo "epw $epw"; grep "dbg" gpurun_out/r02_cluster_dbg_epw$epw.log | head -12 | sed -n "1,2p;5,6p;9,10p"; tail -2 gpurun_out/r02_cluster_dbg_epw$epw.log | grep -v dbg
IDV_CL_EPW=$epw timeout 600 python bench.py --configs 1 --config-kernels --no-cpu --no-eager > gpurun_out/r02_bench_n_epw$epw.json 2> gpurun_out/r02_bench_n_epw$epw.err
tail -c 300 gpurun_out/r02_bench_n_epw$epw.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n_epw$epw.json"))
print("epw=$epw", d["ms_per_step"], d["configs"]["1"]["ms_per_step"], d["configs"]["1"].get("per_kernel_ms"))
PY
done
