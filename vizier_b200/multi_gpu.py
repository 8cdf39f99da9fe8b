"""Candidate-pool sharding across the GPUs of one box.

The scoring path has no cross-candidate dependence: rank r scores candidates
[r*M, (r+1)*M) of one global Philox pool against its own copy of the (deterministically
recomputed) factorisation, so the only exchange is the global top-`count`: one all-gather of
count*(16+8D) bytes per rank followed by the same deterministic merge on every rank (larger
score first, ties -> lower global index).  The reference has no counterpart (single process).
"""

from __future__ import annotations

import numpy as np

from vizier_b200.acquisitions import trust_radius  # re-export for bench.py  # noqa: F401


def merge_topk(indices: np.ndarray, values: np.ndarray, features: np.ndarray, count: int):
  """Deterministic merge of gathered per-rank winners. indices [R*c], values [R*c], features [R*c, D]."""
  v = np.where(np.isnan(values), -np.inf, values)
  order = np.lexsort((indices, -v))[:count]
  return indices[order], values[order], features[order]


class TopkExchange:
  """Reusable buffers for the per-suggest collective: one pinned host staging array, one device
  payload, one device gather target.  Per call: 1 H2D (count*(D+2) doubles), 1 NCCL all-gather,
  1 D2H of world*count*(D+2) doubles."""

  def __init__(self, dist, device, dim: int, count: int):
    import torch
    self.dist, self.count, self.dim = dist, count, dim
    self.world = dist.get_world_size()
    self.host = torch.empty((count, dim + 2), dtype=torch.float64).pin_memory()
    self.payload = torch.empty((count, dim + 2), dtype=torch.float64, device=device)
    self.gathered = torch.empty((self.world * count, dim + 2), dtype=torch.float64, device=device)
    self.host_out = torch.empty((self.world * count, dim + 2), dtype=torch.float64).pin_memory()

  def __call__(self, idx: np.ndarray, val: np.ndarray, x_host: np.ndarray):
    h = self.host.numpy()
    h[:, 0] = val
    h[:, 1] = idx          # global indices < 2^53 are exact in fp64
    h[:, 2:] = x_host
    self.payload.copy_(self.host, non_blocking=True)
    self.dist.all_gather_into_tensor(self.gathered, self.payload)
    self.host_out.copy_(self.gathered, non_blocking=False)
    g = self.host_out.numpy()
    return merge_topk(g[:, 1].astype(np.int64), g[:, 0], g[:, 2:], self.count)


def global_topk(dist, idx: np.ndarray, val: np.ndarray, x, count: int):
  """All-gathers each rank's local top-`count` (global indices, scores, feature rows) and merges.

  idx/val: host arrays [count]; x: torch tensor [count, D] (device or host).  Returns host arrays.
  """
  import torch
  dev = x.device if x.is_cuda else torch.device('cpu')
  ex = TopkExchange(dist, dev, x.shape[1], count) if x.is_cuda else None
  if ex is not None:
    return ex(np.asarray(idx), np.asarray(val), x.cpu().numpy())
  world = dist.get_world_size()
  d = x.shape[1]
  payload = torch.empty((count, d + 2), dtype=torch.float64)
  payload[:, 0] = torch.from_numpy(np.asarray(val, np.float64))
  payload[:, 1] = torch.from_numpy(np.asarray(idx, np.float64))
  payload[:, 2:] = x
  gathered = torch.empty((world * count, d + 2), dtype=torch.float64)
  dist.all_gather_into_tensor(gathered, payload)
  g = gathered.numpy()
  return merge_topk(g[:, 1].astype(np.int64), g[:, 0], g[:, 2:], count)
