#!/bin/bash
# N-GPU run of both bench arms exactly as the driver launches them (+ the single-GPU line for the ratio) and C5.
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n1.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n$N.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/bench_c5.py 2>&1 | tail -1 | tee gpurun_out/bench_c5_n$N.json
if [ "$2" = "ref" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref_n$N.log
fi
