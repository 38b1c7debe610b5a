#!/bin/bash
# Round-2 GPU session A: full GPU test suite, bench line with the `configs` object and per-kernel times, tile-order and
# fused-latent A/B runs, ncu DRAM-traffic capture of the tap-GEMM launches under both tile orders.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r02_tests_a.log
tail -5 gpurun_out/r02_tests_a.log
python bench.py --steps 10 --warmup 3 --config-kernels > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -c 600 gpurun_out/r02_bench_a.err
IDV_OPTIONS=gemm_tile_order=0 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --configs 2b,3 --config-kernels \
  > gpurun_out/r02_bench_a_order0.json 2> gpurun_out/r02_bench_a_order0.err
IDV_FUSED_LATENT=0 python bench.py --steps 10 --warmup 3 --no-cpu --no-eager --configs '' \
  > gpurun_out/r02_bench_a_unfused_latent.json 2> gpurun_out/r02_bench_a_unfused.err
for o in 1 0; do
  IDV_OPTIONS=gemm_tile_order=$o timeout 600 ncu --set full --clock-control none -k regex:tapgemm_tc -s 15 -c 15 --csv --page raw \
    --log-file gpurun_out/r02_ncu_tapgemm_full_order$o.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-eager --configs '' \
    > gpurun_out/r02_ncu_order$o.log 2>&1
done
ls -la gpurun_out | tail -12
