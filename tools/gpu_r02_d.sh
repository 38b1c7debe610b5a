#!/bin/bash
# Round-2 GPU session D (1 GPU): tests; bench with the first encoder layer on the tensor cores, 256 / N output planes per
# tile for every narrow strided layer, step counters on separate lines; A/B switches; launch list of configs 2 and 3.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r02_tests_d.log
tail -5 gpurun_out/r02_tests_d.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-eager"
$B --configs 1,2b,3 --config-kernels > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
tail -c 400 gpurun_out/r02_bench_d.err
IDV_ENC0_TC=0 $B --configs '' > gpurun_out/r02_bench_d_enc0_simt.json 2>> gpurun_out/r02_bench_d.err
IDV_PAIR_PLANES=0 $B --configs '' > gpurun_out/r02_bench_d_no_plane_groups.json 2>> gpurun_out/r02_bench_d.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for c in 2 3; do
  timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_d_config$c.csv \
    python tools/step_launches.py $c > gpurun_out/r02_ncu_launches_d_$c.log 2>&1
done
IDV_LSTM_DBG=1 python tools/step_launches.py 2 2> gpurun_out/r02_lstm_dbg.log > /dev/null
ls -la gpurun_out | tail -8
