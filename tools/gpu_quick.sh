#!/bin/bash
./tools/_bin/potf2_bench
if [ "$1" = "test" ]; then
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cholesky or fit_factor or nll_grad or ard_fit" 2>&1 | tail -4
fi
