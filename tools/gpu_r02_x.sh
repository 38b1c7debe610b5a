#!/bin/bash
# Round-2 GPU session X (1 GPU): staged 16-byte stores of h (lstm_tma_publish = 2) against direct stores (0).
mkdir -p gpurun_out
export IDV_OPTIONS=lstm_tma_publish=2
timeout 1200 python -m pytest tests/test_gpu_abi_units.py tests/test_gpu_parity.py -m gpu -q -x -k "lstm or vae or dccrn or config" 2>&1 | tail -3
python tools/lstm_dbg_2b.py 2>&1 | grep "layer dbg" | tail -6 | head -4 | tee gpurun_out/r02_lstm_dbg_h768_staged_stores.log
python tools/lstm_dbg.py 2>&1 | grep "wave dbg" | head -4 | tee gpurun_out/r02_lstm_dbg_wave_staged_stores.log
for o in 2 0 2 0; do
  IDV_OPTIONS=lstm_tma_publish=$o python bench.py --config-kernels --no-cpu --no-eager --configs 2b > gpurun_out/r02_bench_x.json 2> gpurun_out/r02_bench_x.err
  tail -c 200 gpurun_out/r02_bench_x.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_x.json"))
print("publish=$o", d["ms_per_step"], d["per_kernel_ms"]["idv_lstm2_wave_tc"]["ms"], {k: v.get("ms_per_step") for k, v in d["configs"].items()}, d["configs"]["2b"]["per_kernel_ms"].get("idv_lstm_layer_pair_tc"), d["clocks"]["sm_mhz"])
PY
done
