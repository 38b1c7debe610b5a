#!/bin/bash
# Round-2 GPU session K (1 GPU): streaming step - live-rows tap-GEMMs (TapGemmPack.tc_stream), split-K (idv_tapgemm_tc_splitk),
# dense composed with the first decoder layer - tests, A/B, per-kernel times of an eager step.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_tests_k.log
grep -E "passed|failed|rror" gpurun_out/r02_tests_k.log | tail -5
run() {   # name, IDV_OPTIONS, IDV_STREAM_LIVE_ROWS
  IDV_OPTIONS=$2 IDV_STREAM_LIVE_ROWS=$3 python tools/bench_streaming.py --steps 300 > gpurun_out/r02_streaming_k_$1.log 2>&1
  cp gpurun_out/streaming.json gpurun_out/r02_streaming_k_$1.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_streaming_k_$1.json"))
print("$1", [(c["frames_per_step"], c["kernels_per_step"], round(c["latency_ms_p50"], 4), round(c["latency_ms_p99"], 4)) for c in d["cases"]])
PY
}
run all "" 1
run no_splitk gemm_splitk=0 1
run no_live_rows "" 0
run neither gemm_splitk=0 0
run all_no_pdl launch_pdl=0 1
python tools/stream_profile.py 1 > gpurun_out/r02_stream_profile_k_k1.log 2>&1
tail -28 gpurun_out/r02_stream_profile_k_k1.log
python tools/bench_streaming.py --steps 300 --final > gpurun_out/r02_streaming_k_final_system.log 2>&1
cp gpurun_out/streaming.json gpurun_out/r02_streaming_k_final_system.json
tail -3 gpurun_out/r02_streaming_k_final_system.log
