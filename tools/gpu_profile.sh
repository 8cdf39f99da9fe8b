#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score -s 2 -c 1 -f -o gpurun_out/score_$TAG python tools/profile_score.py 4 > gpurun_out/ncu_score_$TAG.log 2>&1
tail -3 gpurun_out/ncu_score_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench_$TAG.log 2>&1
tail -2 gpurun_out/ncu_bench_$TAG.log | cut -c1-300
