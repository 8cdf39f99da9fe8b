#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dataflow.py -m gpu -q 2>&1 | tail -5
timeout 300 python tools/df_timing.py 1000 2>&1 | grep -B2 -A12 '"step": 14'
timeout 300 python tools/df_timing.py 1000 2>&1 | grep -A9 sum_cycles
timeout 300 python tools/df_timing.py 2000 2>&1 | grep -A9 sum_cycles
timeout 600 python tools/bench_fit.py > gpurun_out/bench_fit_dataflow2.json 2> gpurun_out/bench_fit_dataflow2.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_fit_dataflow2.json'))
for k,v in j.items(): print(k, v if not isinstance(v,dict) else v['median'])
PY
for c in 24 40 96; do echo "== VZGP_DF_CTAS=$c"; VZGP_DF_CTAS=$c timeout 600 python tools/bench_fit.py 2>/dev/null | python -c "
import json,sys
j=json.load(sys.stdin)
for k,v in j.items():
  if 'N1000' in k or 'N2000' in k: print(k, v if not isinstance(v,dict) else v['median'])
"; done
