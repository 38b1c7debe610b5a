#!/bin/bash
# Round-2 GPU session H (1 GPU): two interleaved LSTM chunks per launch - tests, config 3 in passes of 128 with A/B.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r02_tests_h.log
grep -E "passed|failed" gpurun_out/r02_tests_h.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-eager"
$B --configs 1,2b,3 --config-kernels > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err
tail -c 400 gpurun_out/r02_bench_h.err
IDV_OPTIONS=lstm_interleave=0 $B --configs 3 --config-kernels > gpurun_out/r02_bench_h_no_interleave.json 2>> gpurun_out/r02_bench_h.err
IDV_LSTM_DBG=1 python tools/step_launches.py 3 2> gpurun_out/r02_lstm_dbg_interleaved.log > /dev/null
ls -la gpurun_out | tail -5
