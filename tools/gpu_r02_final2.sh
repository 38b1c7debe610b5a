#!/bin/bash
# Round-2 final 1-GPU session of the build with the cluster LSTM and the elect.sync issue: full ncu capture of the tap-GEMM
# launches (source of roofline.traffic), tests, smoke, the full bench line, reference arm, launch lists, streaming, training.
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:tapgemm_tc --csv --page raw \
  --log-file gpurun_out/r02_ncu_tapgemm_full_final2.csv python tools/step_launches.py 2 > gpurun_out/r02_ncu_full_final2.log 2>&1
cp gpurun_out/r02_ncu_tapgemm_full_final2.csv profiles/r02_ncu_tapgemm_full_final2.csv
python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r02_gpu_tests_final2.log
grep -E "passed|failed" gpurun_out/r02_gpu_tests_final2.log
python __graft_entry__.py --smoke > gpurun_out/r02_smoke_final2.log 2>&1; tail -1 gpurun_out/r02_smoke_final2.log
python bench.py --steps 20 --warmup 5 --config-kernels > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err
tail -c 300 gpurun_out/r02_bench_final2.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_final2.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["clocks"])
print({k: (v.get("ms_per_step"), v.get("latency_ms_p50")) for k, v in d["configs"].items()})
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm2.json 2>> gpurun_out/r02_bench_final2.err
tail -c 400 gpurun_out/r02_bench_reference_arm2.json
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for c in 1 2 2b 3; do
  timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_final2_config$c.csv \
    python tools/step_launches.py $c > gpurun_out/r02_ncu_launches_final2_$c.log 2>&1
done
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"lstm_cluster|lstm_wave" --csv --page raw \
  --log-file gpurun_out/r02_ncu_lstm_full_final2.csv python tools/step_launches.py 1 > gpurun_out/r02_ncu_full2_final2.log 2>&1
python tools/bench_streaming.py --steps 300 > gpurun_out/r02_streaming_final2.log 2>&1
cp gpurun_out/streaming.json gpurun_out/r02_streaming_final2.json
python tools/train_steps.py > gpurun_out/r02_train_steps_final2.log 2>&1; tail -4 gpurun_out/r02_train_steps_final2.log
ls -la gpurun_out | tail -14
