#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stepped_eagle.py -x -q -k "like_the_reference or designer" > gpurun_out/pytest_step.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_step.log
