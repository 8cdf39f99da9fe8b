#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "ard or designer or ucb_pe or bandit" 2>&1 | tail -4
timeout 900 python tools/bench_configs.py 2>&1 | tail -45 | tee gpurun_out/bench_configs.json
