#!/bin/bash
# Round-2 GPU session R (1 GPU): cluster LSTM as concurrent chunks of 16 utterances: unit tests, cluster vs wavefront per batch,
# all tests, bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_abi_units.py -m gpu -q -k "lstm2_cluster" 2>&1 | tail -4
timeout 900 python tools/bench_lstm_small.py > gpurun_out/r02_lstm_small_r.log 2>&1; tail -18 gpurun_out/r02_lstm_small_r.log | cut -c1-250
cp gpurun_out/lstm_small.json gpurun_out/r02_lstm_chunks_vs_wave.json
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_gpu_tests_r.log; tail -3 gpurun_out/r02_gpu_tests_r.log
python bench.py --config-kernels --no-cpu --no-eager > gpurun_out/r02_bench_r.json 2> gpurun_out/r02_bench_r.err
tail -c 300 gpurun_out/r02_bench_r.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_r.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
print({k: (v.get("ms_per_step"), v.get("latency_ms_p50")) for k, v in d["configs"].items()})
print(d["configs"]["3"].get("per_kernel_ms"))
PY
