#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multimetric.py tests/test_gpu_dataflow.py -m gpu -q -x 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_designer.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python tools/bench_fit.py > gpurun_out/bench_fit_dataflow4.json 2> gpurun_out/bench_fit_dataflow4.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_fit_dataflow4.json'))
for k,v in j.items():
  if 'N1000' in k or 'N2000' in k: print(k, v if not isinstance(v,dict) else v['median'])
PY
tail -3 gpurun_out/bench_fit_dataflow4.err
