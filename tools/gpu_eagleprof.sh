#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_eagle.py 3000 | tee gpurun_out/eagle_timing.json
VZ_PROFILE_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_eagle.csv python tools/profile_eagle.py 8 > gpurun_out/ncu_eagle.log 2>&1
tail -2 gpurun_out/ncu_eagle.log
