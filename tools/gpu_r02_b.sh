#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench_r02b.json; tail -5 gpurun_out/bench_r02b.err
