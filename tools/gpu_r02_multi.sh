#!/bin/bash
# multi-GPU validation: transports against a host merge, then the bench at N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_multi_gpu_torchrun.py -m gpu -q -x 2>&1 | tail -15
for t in peer nccl torch; do
  echo "== bench --gpus $N transport=$t"
  VZGP_EXCHANGE=$t timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_${N}gpu_$t.json 2> gpurun_out/bench_${N}gpu_$t.err
  python - <<PY
import json
try:
  j=json.loads(open('gpurun_out/bench_${N}gpu_$t.json').read().strip().splitlines()[-1])
  print({k: j[k] for k in ('value','ms_per_step','n_gpus')}, j['config']['per_step_ms'], j['config']['suggest_latency_ms'], j['config']['ranks_agree'], j['config']['exchange_ok'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'])
except Exception as e:
  print('ERR', e); print(open('gpurun_out/bench_${N}gpu_$t.err').read()[-2000:])
PY
done
