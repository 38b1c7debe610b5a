#!/bin/bash
# Round-2 GPU session Z (1 GPU): switch-coverage test of the one-layer LSTM kernel, then all tests / smoke / default bench of HEAD.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abi_units.py -m gpu -q -x -k "switches" 2>&1 | tail -3
bash tools/gpu_r02_t.sh
