#!/bin/bash
# Round-2 GPU session Q (1 GPU): TMA producer of the wavefront LSTM under elect.sync: LSTM tests, timeline, bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_abi_units.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
python tools/lstm_dbg.py 2>&1 | grep "wave dbg" | head -12 > gpurun_out/r02_lstm_dbg_q.log; head -4 gpurun_out/r02_lstm_dbg_q.log
python bench.py --config-kernels --no-cpu --no-eager > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err
tail -c 300 gpurun_out/r02_bench_q.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_q.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
print(d["per_kernel_ms"])
print({k: (v.get("ms_per_step"), v.get("latency_ms_p50")) for k, v in d["configs"].items()})
PY
