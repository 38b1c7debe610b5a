#!/bin/bash
# Round-2 GPU session P (1 GPU): tcgen05 instructions issued under elect.sync (tap-GEMM, wavefront LSTM): all tests, bench
# with per-kernel times, streaming, one-step launch list of config 2.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_gpu_tests_p.log; tail -3 gpurun_out/r02_gpu_tests_p.log
python bench.py --config-kernels > gpurun_out/r02_bench_p.json 2> gpurun_out/r02_bench_p.err
tail -c 300 gpurun_out/r02_bench_p.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_p.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
print(d["per_kernel_ms"])
print({k: (v.get("ms_per_step"), v.get("latency_ms_p50")) for k, v in d["configs"].items()})
PY
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_p_config2.csv \
    python tools/step_launches.py 2 > gpurun_out/r02_ncu_launches_p.log 2>&1
python - <<PY
import csv
lines = [l for l in open("gpurun_out/r02_ncu_launches_p_config2.csv") if l.startswith('"')]
ks, order = {}, []
for x in csv.DictReader(lines):
    k = x["ID"]
    if k not in ks:
        ks[k] = {"name": x["Kernel Name"][:46]}; order.append(k)
    ks[k][x["Metric Name"]] = x["Metric Value"]
tot = 0
for k in order:
    d = float(ks[k]["gpu__time_duration.sum"].replace(",", "")) / 1e6
    tot += d
    print("%-48s %8.3f ms  tensor %5s%%" % (ks[k]["name"], d, ks[k].get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")))
print("sum %.3f ms over %d launches" % (tot, len(order)))
PY
