#!/bin/bash
# Round-2 GPU session T (1 GPU): last verification of HEAD: all GPU tests, smoke, the default bench line and the reference arm.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_gpu_tests_t.log; tail -2 gpurun_out/r02_gpu_tests_t.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py > gpurun_out/r02_bench_t.json 2> gpurun_out/r02_bench_t.err; tail -c 200 gpurun_out/r02_bench_t.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_t.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
print({k: (v.get("ms_per_step"), v.get("latency_ms_p50")) for k, v in d["configs"].items()})
PY
python bench.py --impl reference > gpurun_out/r02_bench_t_ref.json 2>> gpurun_out/r02_bench_t.err; tail -c 300 gpurun_out/r02_bench_t_ref.json
