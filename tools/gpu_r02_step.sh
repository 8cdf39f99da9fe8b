#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stepped_eagle.py -x -q > gpurun_out/pytest_step.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_step.log
timeout 600 python tools/bench_suggest_n1000.py > gpurun_out/bench_suggest_r02.json 2> gpurun_out/bench_suggest_r02.err; echo "suggest exit $?"; cat gpurun_out/bench_suggest_r02.json | head -60; tail -3 gpurun_out/bench_suggest_r02.err
