#!/bin/bash
# Round-2 GPU session F (1 GPU): tests; bench (vectorised STFT rows, multi-buffer store staging, enc0 on tensor cores) with
# the SIMT enc0 as A/B; launch list of config 2.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r02_tests_f.log
grep -E "passed|failed" gpurun_out/r02_tests_f.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-eager"
$B --configs 1,2b,3 --config-kernels > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err
tail -c 400 gpurun_out/r02_bench_f.err
python tools/step_launches.py 2b > /dev/null 2>&1
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_f_config2.csv \
    python tools/step_launches.py 2 > gpurun_out/r02_ncu_launches_f_2.log 2>&1
ls -la gpurun_out | tail -6
