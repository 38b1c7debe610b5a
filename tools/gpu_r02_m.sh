#!/bin/bash
# Round-2 GPU session M (1 GPU): verification of HEAD after the container was re-created — all GPU tests, smoke,
# the default bench line, the streaming bench and the serialised launch list of one streaming step.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_gpu_tests_m.log
grep -E "passed|failed|rror" gpurun_out/r02_gpu_tests_m.log | tail -5
python __graft_entry__.py --smoke > gpurun_out/r02_smoke_m.log 2>&1; tail -1 gpurun_out/r02_smoke_m.log
python bench.py --config-kernels > gpurun_out/r02_bench_m.json 2> gpurun_out/r02_bench_m.err
tail -c 300 gpurun_out/r02_bench_m.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_m.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"], d["roofline"], d["clocks"])
print({k: (v.get("ms_per_step"), v.get("latency_ms_p50")) for k, v in d["configs"].items()})
PY
python tools/bench_streaming.py --steps 300 > gpurun_out/r02_streaming_k4_all.log 2>&1
cp gpurun_out/streaming.json gpurun_out/r02_streaming_k4_all.json
python - <<PY
import json
d = json.load(open("gpurun_out/r02_streaming_k4_all.json"))
print([(c["frames_per_step"], c["kernels_per_step"], round(c["latency_ms_p50"], 4), round(c["latency_ms_p99"], 4)) for c in d["cases"]])
PY
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv \
  --log-file gpurun_out/r02_ncu_stream_k4_splitk1.csv python tools/stream_profile.py 1 > gpurun_out/r02_ncu_stream_k4_splitk1.log 2>&1
tail -3 gpurun_out/r02_ncu_stream_k4_splitk1.log
