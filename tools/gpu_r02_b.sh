#!/bin/bash
# Round-2 GPU session B: tests, bench with A/B switches for the dense + first-decoder-layer composition, the TMA-store
# epilogue and the 8-bin head units, launch list of one step.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r02_tests_b.log
tail -5 gpurun_out/r02_tests_b.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-eager"
$B --configs 1,2b,3 --config-kernels > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
tail -c 400 gpurun_out/r02_bench_b.err
IDV_FUSED_DENSE=0 $B --configs '' > gpurun_out/r02_bench_b_nofuse_dense.json 2>> gpurun_out/r02_bench_b.err
IDV_OPTIONS=gemm_tma_store=0 $B --configs '' > gpurun_out/r02_bench_b_no_tma_store.json 2>> gpurun_out/r02_bench_b.err
IDV_HEAD_BINS=2 $B --configs '' > gpurun_out/r02_bench_b_head_bins2.json 2>> gpurun_out/r02_bench_b.err
IDV_HEAD_BINS=16 $B --configs '' > gpurun_out/r02_bench_b_head_bins16.json 2>> gpurun_out/r02_bench_b.err
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -s 21 -c 21 --csv --log-file gpurun_out/r02_ncu_launches_b.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-eager --configs '' > gpurun_out/r02_ncu_b.log 2>&1
ls -la gpurun_out | tail -12
