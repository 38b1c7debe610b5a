#!/bin/bash
# Round-2 GPU session K2 (1 GPU): split-K as a reduce-scatter through the L2 (plain stores, fixed summation order).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_tests_k2.log
grep -E "passed|failed|rror" gpurun_out/r02_tests_k2.log | tail -5
run() {   # name, IDV_OPTIONS, IDV_STREAM_LIVE_ROWS
  IDV_OPTIONS=$2 IDV_STREAM_LIVE_ROWS=$3 python tools/bench_streaming.py --steps 300 > gpurun_out/r02_streaming_k2_$1.log 2>&1
  cp gpurun_out/streaming.json gpurun_out/r02_streaming_k2_$1.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_streaming_k2_$1.json"))
print("$1", [(c["frames_per_step"], c["kernels_per_step"], round(c["latency_ms_p50"], 4), round(c["latency_ms_p99"], 4)) for c in d["cases"]])
PY
}
run all "" 1
run no_splitk gemm_splitk=0 1
run all_no_pdl launch_pdl=0 1
python tools/stream_profile.py 1 > gpurun_out/r02_stream_profile_k2_k1.log 2>&1
tail -26 gpurun_out/r02_stream_profile_k2_k1.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --configs 1 > gpurun_out/r02_bench_k2.json 2> gpurun_out/r02_bench_k2.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_k2.json"))
print(d["ms_per_step"], d["e2e"]["value"], {k: v.get("ms_per_step") for k, v in d["configs"].items()})
PY
