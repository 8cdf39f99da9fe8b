"""Run under torchrun (one rank per GPU): every transport of TopkExchange must give the host merge of the
per-rank winners on every rank, over several pipelined steps.  Prints 'MULTI_GPU_CHECK OK' on rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from vizier_b200 import gp, multi_gpu
from vizier_b200.acquisitions import trust_radius

local = int(os.environ.get('LOCAL_RANK', '0'))
os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
n, d, m, count = 300, 7, 20_000, 4
rng = np.random.default_rng(0)
x = rng.uniform(size=(n, d)); y = rng.normal(size=n)
dev = gp.DeviceGP(local)
dev.fit(x, y, gp.GPHyperParams(1.0, np.full(d, 0.8), 1e-2))
acq = gp.Acquisition(1.8, True, trust_radius(n, d, 0))
pools = [dev.random_pool(m, d, seed=11, index_base=(rank * 3 + i) * m) for i in range(3)]
results = {}
for transport in ('peer', 'nccl', 'torch'):
  ex = multi_gpu.TopkExchange(dist, dev, d, count, slots=4, transport=transport)
  got = []
  steps = 9
  for i in range(steps):
    ex.step(i % 4, pools[i % 3], acq, index_base=(rank * 3 + i % 3) * m)
    if i >= 3:
      got.append(ex.result((i - 3) % 4))
  for i in range(steps - 3, steps):
    got.append(ex.result(i % 4))
  if ex.peer is not None:
    assert ex.peer.status() == 0, 'exchange timed out'
  results[transport] = got
  dist.barrier()
# cross-check the transports against each other and against a host merge through a gloo group
g = dist.new_group(backend='gloo')
for i in range(9):
  bx, bs, bi = dev.score_topk(pools[i % 3], acq, count)
  wi, wv, wx = multi_gpu.global_topk(dist, bi + (rank * 3 + i % 3) * m, bs, torch.from_numpy(bx), count, group=g)
  for transport in ('peer', 'nccl', 'torch'):
    idx, val, feat = results[transport][i]
    np.testing.assert_array_equal(idx, wi, err_msg=f'{transport} step {i}')
    np.testing.assert_array_equal(val, wv)
    np.testing.assert_array_equal(feat, wx)
# e2e host path
hx = pools[0].cpu().pin_memory(); hs = torch.empty(m, dtype=torch.float64).pin_memory()
ex = multi_gpu.TopkExchange(dist, dev, d, count, transport='peer')
idx, val, feat = ex.suggest_host(hx, acq, rank * 3 * m, hs)
np.testing.assert_array_equal(idx, results['peer'][0][0])
dist.barrier()
if rank == 0:
  print('MULTI_GPU_CHECK OK world=%d' % world, flush=True)
dist.destroy_process_group()
