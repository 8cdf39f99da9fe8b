#!/bin/bash
# 2-GPU check with the tcgen05 scoring kernel as the default: torchrun tests + bench over the fused exchange
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_multi_gpu_torchrun.py -m gpu -q -x 2>&1 | tail -5; fi
VZGP_EXCHANGE=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_${N}gpu_i8.json 2> gpurun_out/bench_${N}gpu_i8.err
python - <<PY
import json
try:
  j=json.loads(open('gpurun_out/bench_${N}gpu_i8.json').read().strip().splitlines()[-1])
  print({k: j[k] for k in ('value','ms_per_step','n_gpus')}, j['config']['per_step_ms'], j['config']['ranks_agree'], j['config']['exchange_ok'], 'e2e', j['e2e']['value'], j['roofline']['kernel'], j['roofline']['frac'])
except Exception as e:
  print('ERR', e); print(open('gpurun_out/bench_${N}gpu_i8.err').read()[-2000:])
PY
if [ "$N" = "8" ]; then
VZGP_EXCHANGE=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --workload c5 --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_c5_i8.json 2> gpurun_out/bench_${N}gpu_c5_i8.err
python - <<PY
import json
try:
  j=json.loads(open('gpurun_out/bench_${N}gpu_c5_i8.json').read().strip().splitlines()[-1])
  print('c5', {k: j[k] for k in ('value','ms_per_step','n_gpus')}, j['config']['ranks_agree'], j['roofline']['kernel'])
except Exception as e:
  print('ERR', e); print(open('gpurun_out/bench_${N}gpu_c5_i8.err').read()[-1500:])
PY
fi
