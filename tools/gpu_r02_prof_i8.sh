#!/bin/bash
mkdir -p gpurun_out
timeout 240 ncu --clock-control none --target-processes application-only --set full --import-source on -k regex:k_score_i8 -s 2 -c 1 -f -o gpurun_out/score_i8_r02 python tools/profile_score.py 4 > gpurun_out/ncu_score_i8_r02.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_score_i8_r02.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "c2_full_pool" 2>&1 | tail -2
