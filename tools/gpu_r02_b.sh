#!/bin/bash
mkdir -p gpurun_out
echo "== parity with the dataflow factorisation"
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
echo "== exchange"
timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -x -q 2>&1 | tail -15
echo "== fit latency: dataflow"
timeout 600 python tools/bench_fit.py > gpurun_out/bench_fit_dataflow.json 2> gpurun_out/bench_fit_dataflow.err; cat gpurun_out/bench_fit_dataflow.json; tail -3 gpurun_out/bench_fit_dataflow.err
echo "== fit latency: round-1 panel kernels"
VZGP_DATAFLOW=0 timeout 600 python tools/bench_fit.py > gpurun_out/bench_fit_panels.json 2> gpurun_out/bench_fit_panels.err; cat gpurun_out/bench_fit_panels.json
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench_r02b.json; tail -5 gpurun_out/bench_r02b.err
