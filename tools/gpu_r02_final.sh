#!/bin/bash
# last check of the round at HEAD: every -m gpu test and smoke()
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r02.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
