"""World-size-2 gloo test of the only cross-rank step of the path: the global top-`count` merge
(vizier_b200/multi_gpu.py).  Each rank owns a shard of one scored candidate pool; after one
all-gather every rank must hold the same winners as a single-process sort of the whole pool."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vizier_b200 import multi_gpu


def _free_port():
  s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, count, q):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  rng = np.random.default_rng(123)
  m, d = 1000, 6
  scores = rng.normal(size=world * m)
  scores[[10, m + 10]] = 9.0            # a cross-rank tie: lower global index must win
  scores[5] = np.nan
  feats = rng.uniform(size=(world * m, d))
  lo = rank * m
  s_loc = np.where(np.isnan(scores[lo:lo + m]), -np.inf, scores[lo:lo + m])
  order = np.lexsort((np.arange(m), -s_loc))[:count]
  idx, val, x = multi_gpu.global_topk(dist, order + lo, scores[lo:lo + m][order], torch.from_numpy(feats[lo:lo + m][order]), count)
  q.put((rank, idx.tolist(), val.tolist(), x.tolist()))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('count', [1, 4])
def test_global_topk_world2(count):
  world, port = 2, _free_port()
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, world, port, count, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in range(world)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  rng = np.random.default_rng(123)
  scores = rng.normal(size=world * 1000); scores[[10, 1010]] = 9.0; scores[5] = np.nan
  feats = rng.uniform(size=(world * 1000, 6))
  v = np.where(np.isnan(scores), -np.inf, scores)
  want = np.lexsort((np.arange(v.size), -v))[:count]
  for rank, idx, val, x in res:
    assert idx == want.tolist()
    np.testing.assert_array_equal(val, scores[want])
    np.testing.assert_array_equal(np.asarray(x), feats[want])
  assert want[0] == 10


def test_merge_topk_is_deterministic():
  idx = np.array([7, 3, 9, 1]); val = np.array([1.0, 2.0, 2.0, np.nan]); f = np.arange(8.0).reshape(4, 2)
  i, v, x = multi_gpu.merge_topk(idx, val, f, 3)
  assert i.tolist() == [3, 9, 7]
