#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_score_i8.py -x -q > gpurun_out/pytest_i8.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_i8.log
timeout 200 python tools/bench_i8.py > gpurun_out/bench_i8.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_i8.log
timeout 240 ncu --clock-control none --target-processes application-only --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:k_score_i8 -s 2 -c 2 --csv --log-file gpurun_out/i8_traffic.csv python tools/profile_score.py 5 > gpurun_out/ncu_i8_traffic.log 2>&1; echo "ncu exit $?"; grep -v "^==" gpurun_out/i8_traffic.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -8
