#!/bin/bash
# One gpurun session: tests, smoke, fp64 peak, bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== fp64 peak"; timeout 120 ./tools/_bin/fp64_peak | tee gpurun_out/fp64_peak.json
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
echo "== configs"; timeout 900 python tools/bench_configs.py 2>&1 | tail -40 | tee gpurun_out/bench_configs.json
