#!/bin/bash
mkdir -p gpurun_out
./tools/_bin/potf2_la_bench
timeout 900 python -m pytest tests/test_gpu_dataflow.py -m gpu -q -x 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python tools/bench_fit.py > gpurun_out/bench_fit_dataflow7.json 2> gpurun_out/bench_fit_dataflow7.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_fit_dataflow7.json'))
for k,v in j.items():
  if 'N1000' in k or 'N2000' in k: print(k, v if not isinstance(v,dict) else v['median'])
PY
