#!/bin/bash
# Round-2 GPU session V (1 GPU): per-K-chunk step counters of the one-layer CTA-pair LSTM (H = 768, training forward): tests,
# timeline, config 2b / 4 with and without.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_abi_units.py tests/test_gpu_parity.py tests/test_train_step.py -m gpu -q -x 2>&1 | tail -3
python tools/lstm_dbg_2b.py 2>&1 | grep "layer dbg" | tail -6 | tee gpurun_out/r02_lstm_dbg_h768_chunk_sync.log
for o in 1 0; do
  IDV_OPTIONS=lstm_chunk_sync=$o python bench.py --config-kernels --no-cpu --no-eager --configs 2b,4 > gpurun_out/r02_bench_v_chunk$o.json 2> gpurun_out/r02_bench_v.err
  tail -c 200 gpurun_out/r02_bench_v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_v_chunk$o.json"))
print("chunk_sync=$o", d["ms_per_step"], {k: v.get("ms_per_step") for k, v in d["configs"].items()}, d["configs"]["2b"]["per_kernel_ms"].get("idv_lstm_layer_pair_tc"), d["clocks"])
PY
done
