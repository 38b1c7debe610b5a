#!/bin/bash
# Round-2 GPU session J (1 GPU): streaming step 29 -> 24 launches (fused latent stage, head writes the iSTFT rows, one tail
# kernel) and programmatic dependent launch (launch_pdl) - tests, A/B, per-kernel times of an eager step.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_tests_j.log
grep -E "passed|failed" gpurun_out/r02_tests_j.log
for pdl in 1 0; do
  IDV_OPTIONS=launch_pdl=$pdl python tools/bench_streaming.py --steps 300 > gpurun_out/r02_streaming_j_pdl$pdl.log 2>&1
  cp gpurun_out/streaming.json gpurun_out/r02_streaming_j_pdl$pdl.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_streaming_j_pdl$pdl.json"))
print("pdl=$pdl", [(c["frames_per_step"], c["kernels_per_step"], round(c["latency_ms_p50"], 4)) for c in d["cases"]])
PY
done
python tools/stream_profile.py 1 > gpurun_out/r02_stream_profile_j_k1.log 2>&1
cat gpurun_out/r02_stream_profile_j_k1.log | tail -30
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --configs 1"
for pdl in 1 0; do
  IDV_OPTIONS=launch_pdl=$pdl $B > gpurun_out/r02_bench_j_pdl$pdl.json 2> gpurun_out/r02_bench_j_pdl$pdl.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_j_pdl$pdl.json"))
print("pdl=$pdl", d["ms_per_step"], d["e2e"]["value"], {k: v.get("ms_per_step") for k, v in d["configs"].items()})
PY
done
