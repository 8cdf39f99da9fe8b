#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multi_gpu_torchrun.py -m gpu -q -x 2>&1 | tail -4
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 3 $2 > gpurun_out/bench_${N}gpu_$1.json 2> gpurun_out/bench_${N}gpu_$1.err
  python - <<PY
import json
try:
  j=json.loads(open('gpurun_out/bench_${N}gpu_$1.json').read().strip().splitlines()[-1])
  print('$1', {k: j[k] for k in ('value','ms_per_step','n_gpus')}, j['config']['per_step_ms'], 'lat', j['config']['suggest_latency_ms'], j['config']['ranks_agree'], j['config']['exchange_ok'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'kernel_ms', j['roofline']['kernel_ms'])
except Exception as e:
  print('ERR', e); print(open('gpurun_out/bench_${N}gpu_$1.err').read()[-1500:])
PY
}
run c2_peer ""
VZGP_EXCHANGE=torch run c2_torch ""
run c5_peer "--workload c5"
timeout 300 python bench.py --steps 20 --warmup 3 --no-suggest > gpurun_out/bench_1gpu_on8box.json 2>/dev/null; python -c "
import json; j=json.loads(open('gpurun_out/bench_1gpu_on8box.json').read().strip().splitlines()[-1]); print('1gpu', j['value'], j['ms_per_step'])"
