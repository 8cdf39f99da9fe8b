#!/bin/bash
# round 2, call A: full GPU test-suite (with the new C2/C3/C4 full-size parity tests), smoke, bench.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02a.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu_r02a.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; tail -c 1500 gpurun_out/bench_r02a.json
