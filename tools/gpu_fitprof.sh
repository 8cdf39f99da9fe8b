#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fit.csv python tools/profile_fit.py > gpurun_out/ncu_fit.log 2>&1
tail -2 gpurun_out/ncu_fit.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
