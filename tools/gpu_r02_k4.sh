#!/bin/bash
# Round-2 GPU session K4 (1 GPU): split-K with the coalesced workspace, only for layers with few 64-column tiles.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_abi_units.py tests/test_streaming.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_tests_k4.log
grep -E "passed|failed|rror" gpurun_out/r02_tests_k4.log | tail -5
run() {   # name, IDV_OPTIONS
  IDV_OPTIONS=$2 python tools/bench_streaming.py --steps 300 > gpurun_out/r02_streaming_k4_$1.log 2>&1
  cp gpurun_out/streaming.json gpurun_out/r02_streaming_k4_$1.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_streaming_k4_$1.json"))
print("$1", [(c["frames_per_step"], c["kernels_per_step"], round(c["latency_ms_p50"], 4), round(c["latency_ms_p99"], 4)) for c in d["cases"]])
PY
}
run all ""
run no_splitk gemm_splitk=0
bash -c 'sed -n "/^for sk/,\$p" tools/gpu_r02_k3.sh | sed "s/k1_splitk/k4_splitk/g" > /tmp/k3.sh; bash /tmp/k3.sh' | grep -E "tapgemm|sum"
