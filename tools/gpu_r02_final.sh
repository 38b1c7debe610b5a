#!/bin/bash
# Round-2 final 1-GPU session: tests, smoke, the full bench line, reference arm, one-step launch lists, full ncu captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r02_gpu_tests_final.log
grep -E "passed|failed" gpurun_out/r02_gpu_tests_final.log
python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
python bench.py --steps 20 --warmup 5 --config-kernels > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 300 gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_final.err
python tools/train_steps.py > gpurun_out/r02_train_steps.log 2>&1; tail -4 gpurun_out/r02_train_steps.log
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for c in 2 2b 3; do
  timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_final_config$c.csv \
    python tools/step_launches.py $c > gpurun_out/r02_ncu_launches_final_$c.log 2>&1
done
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:tapgemm_tc --csv --page raw \
  --log-file gpurun_out/r02_ncu_tapgemm_full_final.csv python tools/step_launches.py 2 > gpurun_out/r02_ncu_full.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"lstm_wave|latent_fused|stft_frames|ola_kernel" --csv --page raw \
  --log-file gpurun_out/r02_ncu_lstm_small_full_final.csv python tools/step_launches.py 2 > gpurun_out/r02_ncu_full2.log 2>&1
ls -la gpurun_out | tail -12
