#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-suggest --no-cpu > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "bench exit $?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
print({k: j[k] for k in ('value','ms_per_step')}, 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
PY
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "host" 2>&1 | tail -2
