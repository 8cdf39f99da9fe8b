#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py -x -q > gpurun_out/pytest_step.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_step.log
