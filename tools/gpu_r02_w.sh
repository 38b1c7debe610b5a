#!/bin/bash
# Round-2 GPU session W (1 GPU): h(t) published with one tensor store per CTA and step (lstm_tma_publish): tests, timelines,
# configs 2 / 2b / 4 with and without.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_abi_units.py tests/test_gpu_parity.py tests/test_train_step.py -m gpu -q -x 2>&1 | tail -3
python tools/lstm_dbg_2b.py 2>&1 | grep "layer dbg" | tail -6 | head -4 | tee gpurun_out/r02_lstm_dbg_h768_tma_publish.log
python tools/lstm_dbg.py 2>&1 | grep "wave dbg" | head -4 | tee gpurun_out/r02_lstm_dbg_wave_tma_publish.log
for o in "lstm_tma_publish=1,lstm_chunk_sync=1" "lstm_tma_publish=1,lstm_chunk_sync=0" "lstm_tma_publish=0,lstm_chunk_sync=0"; do
  IDV_OPTIONS=$o python bench.py --config-kernels --no-cpu --no-eager --configs 2b,4 > gpurun_out/r02_bench_w.json 2> gpurun_out/r02_bench_w.err
  tail -c 200 gpurun_out/r02_bench_w.err
  cp gpurun_out/r02_bench_w.json "gpurun_out/r02_bench_w_$o.json"
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_w.json"))
print("$o", d["ms_per_step"], d["per_kernel_ms"]["idv_lstm2_wave_tc"]["ms"], {k: v.get("ms_per_step") for k, v in d["configs"].items()}, d["configs"]["2b"]["per_kernel_ms"].get("idv_lstm_layer_pair_tc"), d["clocks"]["sm_mhz"])
PY
done
