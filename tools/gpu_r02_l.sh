#!/bin/bash
# Round-2 GPU session L (1 GPU): all tests of the build with split-K / stream tables / HostPipeline, the default bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_tests_l.log
grep -E "passed|failed|rror" gpurun_out/r02_tests_l.log | tail -5
python bench.py > gpurun_out/r02_bench_l.json 2> gpurun_out/r02_bench_l.err
tail -c 300 gpurun_out/r02_bench_l.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_l.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"], d["roofline"], d["clocks"])
print({k: (v.get("ms_per_step"), v.get("latency_ms_p50")) for k, v in d["configs"].items()})
PY
