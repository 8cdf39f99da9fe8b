#!/bin/bash
# full single-GPU evidence: every -m gpu test, smoke, bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_r02.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
# the FP64 DMMA kernel on the large pools as well (it stays the reference implementation of the tcgen05 kernel)
VZGP_SCORE_I8=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k 'c2_full_pool or score' 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err; echo "bench exit $?"; python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r02_1gpu.json').read().strip().splitlines()[-1])
print({k: j[k] for k in ('value','ms_per_step')}, 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'])
print('cpu_baseline', j.get('cpu_baseline'))
s=j.get('suggest_e2e', {})
print('suggest_e2e', {k: s.get(k) for k in ('gpu_s','cpu_s','speedup','breakdown')}, s.get('gpu',{}).get('eagle_default',{}).get('gpu_s'), s.get('error'))
PY
tail -3 gpurun_out/bench_r02_1gpu.err
timeout 400 python bench.py --workload c5 --steps 10 --warmup 3 --no-suggest > gpurun_out/bench_r02_c5_1gpu.json 2> gpurun_out/bench_r02_c5_1gpu.err; echo "bench c5 exit $?"; tail -c 1500 gpurun_out/bench_r02_c5_1gpu.json | cut -c1-1500
