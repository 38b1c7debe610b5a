#!/bin/bash
# Round-2 GPU session C (1 GPU): tests, full bench line, A/B of the head-written spectrum rows, launch list of one step of
# configs 2 / 2b, ncu --set full of the tap-GEMM launches of one step (DRAM traffic for roofline.traffic).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r02_tests_c.log
tail -5 gpurun_out/r02_tests_c.log
python bench.py --steps 10 --warmup 3 --config-kernels > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
tail -c 400 gpurun_out/r02_bench_c.err
IDV_FUSED_SPEC_ROWS=0 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --configs '' > gpurun_out/r02_bench_c_no_rows.json 2>> gpurun_out/r02_bench_c.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c_reference.json 2>> gpurun_out/r02_bench_c.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for c in 2 2b; do
  timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_config$c.csv \
    python tools/step_launches.py $c > gpurun_out/r02_ncu_launches_$c.log 2>&1
done
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:tapgemm_tc --csv --page raw \
  --log-file gpurun_out/r02_ncu_tapgemm_full_final.csv python tools/step_launches.py 2 > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
