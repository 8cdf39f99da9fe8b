"""BASELINE config C5: M = 1,000,000 candidates sharded 125,000 per GPU over 8 GPUs, N=2000, D=50,
global top-1 and top-`count`.  Launch like bench.py:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c5.py
(also runs on fewer GPUs: the per-GPU shard stays 125,000)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vizier_b200 import gp
from vizier_b200.multi_gpu import TopkExchange, trust_radius
import bench

N, D, M_SHARD, COUNT, STEPS, WARM = 2000, 50, 125_000, 8, 10, 3
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
dist = None
torch.cuda.set_device(local)
if world > 1:
  import torch.distributed as dist
  os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
  dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rng = np.random.default_rng(0)
x = rng.uniform(size=(N, D)); y = -np.sum((x - 0.3) ** 2, axis=1) + 0.05 * rng.normal(size=N); y = (y - y.mean()) / y.std()
dev = gp.DeviceGP(local)
t0 = time.perf_counter()
dev.fit(x, y, gp.GPHyperParams(1.0, 0.5 * (1 + np.arange(D) / D), 1e-3))
fit_s = time.perf_counter() - t0
acq = gp.Acquisition(1.8, True, trust_radius(N, D, 0))
pools = [dev.random_pool(M_SHARD, D, seed=5, index_base=(rank * 3 + i) * M_SHARD) for i in range(3)]   # 150 MB > L2
score = torch.empty(M_SHARD, dtype=torch.float64, device=dev.device)
out = {}
for count in (1, COUNT):
  ex = TopkExchange(dist, dev, D, count)
  for i in range(WARM):
    ex.step(i % 2, pools[i % 3], acq, index_base=(rank * 3 + i % 3) * M_SHARD, score_out=score)
  ex.result((WARM - 1) % 2)
  if dist is not None: dist.barrier()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(dev.stream)
  for i in range(STEPS):
    ex.step(i % 2, pools[i % 3], acq, index_base=(rank * 3 + i % 3) * M_SHARD, score_out=score)
    if i: ex.result((i - 1) % 2)
  e1.record(dev.stream)
  gi, gv, gx = ex.result((STEPS - 1) % 2)
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  if dist is not None:
    t = torch.tensor([ms], dtype=torch.float64, device=dev.device); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    chk = torch.tensor([float(gi[0]), float(gv[0])], dtype=torch.float64, device=dev.device)
    allc = torch.empty((world, 2), dtype=torch.float64, device=dev.device); dist.all_gather_into_tensor(allc, chk)
    out[f'top{count}_ranks_agree'] = bool((allc == allc[0]).all().item())
  out[f'top{count}_ms_per_step'] = ms / STEPS
  out[f'top{count}_candidates_per_s'] = M_SHARD * world * STEPS / (ms * 1e-3)
  out[f'top{count}_winner'] = [int(gi[0]), float(gv[0])]
flops = bench.algorithmic_flops_per_candidate(N, D) * M_SHARD
peak, _ = bench.fp64_peak_tflops()
out.update({'config': f'C5: N={N}, D={D}, M={M_SHARD} per GPU x {world} GPUs', 'fit_s_first_call': fit_s,
            'flops_per_candidate': bench.algorithmic_flops_per_candidate(N, D),
            'per_gpu_tflops_top1': flops / (out['top1_ms_per_step'] * 1e-3) * 1e-12,
            'frac_of_fp64_peak': flops / (out['top1_ms_per_step'] * 1e-3) * 1e-12 / peak})
if rank == 0:
  print(json.dumps(out))
if dist is not None:
  dist.destroy_process_group()
