#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ard or nll" 2>&1 | tail -12
timeout 600 python tools/bench_fit.py > gpurun_out/bench_fit_dataflow8.json 2> gpurun_out/bench_fit_dataflow8.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_fit_dataflow8.json'))
for k,v in j.items(): print(k, v if not isinstance(v,dict) else v['median'])
PY
tail -5 gpurun_out/bench_fit_dataflow8.err
timeout 900 python -m pytest tests/test_gpu_designer.py tests/test_gpu_multimetric.py -m gpu -q -x 2>&1 | tail -6
