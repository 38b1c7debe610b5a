#!/bin/bash
# Round-2 GPU session O (1 GPU): cluster LSTM generalised to <= 32 utterances: unit tests, cluster vs wavefront per batch size,
# then every GPU test and the default bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_abi_units.py -m gpu -q -k "lstm2_cluster" 2>&1 | tail -8
timeout 900 python tools/bench_lstm_small.py > gpurun_out/r02_lstm_small.log 2>&1; tail -18 gpurun_out/r02_lstm_small.log
cp gpurun_out/lstm_small.json gpurun_out/r02_lstm_small.json
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_gpu_tests_o.log; tail -4 gpurun_out/r02_gpu_tests_o.log
python bench.py --config-kernels > gpurun_out/r02_bench_o.json 2> gpurun_out/r02_bench_o.err
tail -c 300 gpurun_out/r02_bench_o.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_o.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
print({k: (v.get("ms_per_step"), v.get("latency_ms_p50")) for k, v in d["configs"].items()})
PY
