#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_score_i8.py -x -q > gpurun_out/pytest_i8.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_i8.log
timeout 200 python tools/bench_i8.py > gpurun_out/bench_i8.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_i8.log
timeout 200 python tools/bench_i8.py 2000 50 > gpurun_out/bench_i8_c4.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_i8_c4.log
