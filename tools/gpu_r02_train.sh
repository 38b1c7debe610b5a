#!/bin/bash
mkdir -p gpurun_out
python tools/train_profile.py 32 2 3 > gpurun_out/r02_train_profile_e2e.log 2>&1
tail -70 gpurun_out/r02_train_profile_e2e.log
