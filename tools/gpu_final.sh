#!/bin/bash
# End-of-round evidence: smoke, full GPU tests, bench line, launch lists (bench step + default Eagle loop), ncu of k_score.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest"; timeout 1300 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_final.json | cut -c1-400
echo "== launch list: bench step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== launch list: small-pool scoring (25 candidates, N=1000)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_small_pool.csv python tools/profile_small_score.py > /dev/null 2>&1
grep -E "k_cross|k_var|k_small" gpurun_out/launches_small_pool.csv | tail -3 | cut -c1-260
echo "== ncu k_score"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score -s 2 -c 1 -f -o gpurun_out/score_final python tools/profile_score.py 4 > gpurun_out/ncu_score_final.log 2>&1; tail -1 gpurun_out/ncu_score_final.log
