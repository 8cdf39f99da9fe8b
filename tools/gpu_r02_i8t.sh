#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_score_i8.py -x -q > gpurun_out/pytest_i8.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_i8.log
timeout 200 python tools/bench_i8.py > gpurun_out/bench_i8.log 2>&1; echo "bench exit $?"; tail -2 gpurun_out/bench_i8.log
VZGP_LIB=$PWD/vizier_b200/_lib/libvzgp_i8timing.so timeout 200 python tools/i8_timing.py > gpurun_out/i8_timing.log 2>&1; echo "timing exit $?"; tail -3 gpurun_out/i8_timing.log
