#!/bin/bash
# Round-2 GPU session Y (1 GPU): release-only publish (lstm_sync_mode = 2) against fence + atomic (0).
mkdir -p gpurun_out
IDV_OPTIONS=lstm_sync_mode=2 timeout 900 python -m pytest tests/test_gpu_abi_units.py -m gpu -q -x -k "lstm" 2>&1 | tail -2
IDV_OPTIONS=lstm_sync_mode=2 python tools/lstm_dbg_2b.py 2>&1 | grep "layer dbg" | tail -6 | head -3 | tee gpurun_out/r02_lstm_dbg_h768_sync2.log
IDV_OPTIONS=lstm_sync_mode=2 python tools/lstm_dbg.py 2>&1 | grep "wave dbg" | head -3 | tee gpurun_out/r02_lstm_dbg_wave_sync2.log
for o in 2 0 2 0; do
  IDV_OPTIONS=lstm_sync_mode=$o python bench.py --config-kernels --no-cpu --no-eager --configs 2b > gpurun_out/r02_bench_y.json 2> gpurun_out/r02_bench_y.err
  tail -c 200 gpurun_out/r02_bench_y.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_y.json"))
print("sync_mode=$o", d["ms_per_step"], d["per_kernel_ms"]["idv_lstm2_wave_tc"]["ms"], {k: v.get("ms_per_step") for k, v in d["configs"].items()}, d["configs"]["2b"]["per_kernel_ms"].get("idv_lstm_layer_pair_tc"), d["clocks"]["sm_mhz"])
PY
done
