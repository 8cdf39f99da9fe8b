#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/df_diag.py 2000 6 2>&1 | grep -v "bad tiles 0" | tail -14; echo "diag done"
timeout 900 python -m pytest tests/test_gpu_dataflow.py -m gpu -q 2>&1 | tail -8
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/bench_fit.py > gpurun_out/bench_fit_dataflow.json 2> gpurun_out/bench_fit_dataflow.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_fit_dataflow.json'))
for k,v in j.items(): print(k, v if not isinstance(v,dict) else v['median'])
PY
