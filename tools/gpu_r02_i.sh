#!/bin/bash
# Round-2 GPU session I (1 GPU): LSTM h ingest by cp.async (lstm_ingest) and small-batch boxes - tests, A/B, in-kernel timeline.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_tests_i.log
grep -E "passed|failed" gpurun_out/r02_tests_i.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --config-kernels"
for opt in lstm_ingest=1 lstm_ingest=0 lstm_ingest=1,lstm_sync_mode=2 lstm_ingest=0,lstm_sync_mode=2; do
  IDV_OPTIONS=$opt $B --configs 1,2b,3 > gpurun_out/r02_bench_i_$opt.json 2> gpurun_out/r02_bench_i_$opt.err
  tail -c 300 gpurun_out/r02_bench_i_$opt.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_i_$opt.json"))
print("$opt", d["ms_per_step"], {k: v.get("ms_per_step") for k, v in d["configs"].items()})
PY
done
for opt in lstm_ingest=1 lstm_ingest=0 lstm_ingest=1,lstm_sync_mode=2; do
  IDV_OPTIONS=$opt IDV_LSTM_DBG=1 python tools/step_launches.py 2 2> gpurun_out/r02_lstm_dbg_i_$opt.log > /dev/null
  head -16 gpurun_out/r02_lstm_dbg_i_$opt.log
done
