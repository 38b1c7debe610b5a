#!/bin/bash
# Round-2 GPU session K3 (1 GPU): true kernel durations of a streaming step (ncu, serialised) with and without split-K.
mkdir -p gpurun_out
for sk in 1 0; do
  IDV_OPTIONS=gemm_splitk=$sk ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv \
    --log-file gpurun_out/r02_ncu_stream_k1_splitk$sk.csv python tools/stream_profile.py 1 > gpurun_out/r02_ncu_stream_k1_splitk$sk.log 2>&1
  python - <<PY
import csv
rows = []
with open("gpurun_out/r02_ncu_stream_k1_splitk$sk.csv") as f:
    lines = [l for l in f if l.startswith('"')]
r = list(csv.DictReader(lines))
ks = {}
order = []
for x in r:
    key = x["ID"]
    if key not in ks:
        ks[key] = {"name": x["Kernel Name"][:40]}
        order.append(key)
    ks[key][x["Metric Name"]] = x["Metric Value"]
last = order[-22:]
tot = 0
for k in last:
    d = float(ks[k]["gpu__time_duration.sum"].replace(",", "")) / 1000.0
    tot += d
    print("%-42s grid %6s  %7.2f us" % (ks[k]["name"], ks[k].get("launch__grid_size"), d))
print("splitk=$sk sum %.1f us" % tot)
PY
done
