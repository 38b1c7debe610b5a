#!/bin/bash
# Round-2 GPU session S (1 GPU): wavefront LSTM with 16 epilogue warps (N = 64): tests, timeline, LSTM time under ncu, bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_abi_units.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python tools/lstm_dbg.py 2>&1 | grep "wave dbg" | head -4 > gpurun_out/r02_lstm_dbg_s.log; cat gpurun_out/r02_lstm_dbg_s.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:lstm --csv --log-file gpurun_out/r02_ncu_lstm_s.csv \
    python tools/step_launches.py 2 > /dev/null 2>&1; grep lstm gpurun_out/r02_ncu_lstm_s.csv | cut -c1-200 | tail -2
python bench.py --config-kernels --no-cpu --no-eager > gpurun_out/r02_bench_s.json 2> gpurun_out/r02_bench_s.err
tail -c 300 gpurun_out/r02_bench_s.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_s.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
print(d["per_kernel_ms"]["idv_lstm2_wave_tc"])
print({k: (v.get("ms_per_step"), v.get("latency_ms_p50")) for k, v in d["configs"].items()})
PY
