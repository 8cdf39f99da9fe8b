#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_transfer.py -q -k "like_the_reference" > gpurun_out/pytest_step.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_step.log
