#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_eagle.py 3000 | tee gpurun_out/eagle_timing.json
