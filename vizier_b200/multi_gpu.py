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
  """The per-suggest collective on device: every rank's packed winners (`DeviceGP.score_topk_pack`
  rows [score, global index, features]) are all-gathered with NCCL on the GP handle's stream and
  merged by the same deterministic kernel on every rank (`DeviceGP.merge_topk`); the merged rows
  land in a pinned host buffer by an asynchronous copy.  Nothing here blocks the host, so
  consecutive suggest steps queue back to back; `result(slot)` waits for one step's event.

  Two buffer slots let step i+1 be enqueued while step i's result is still being read.
  """

  def __init__(self, dist, gp_dev, dim: int, count: int, slots: int = 2):
    import torch
    self.dist, self.dev, self.count, self.dim = dist, gp_dev, count, dim
    self.world = dist.get_world_size() if dist is not None else 1
    device = gp_dev.device
    w = dim + 2
    self.payload = [torch.empty((count, w), dtype=torch.float64, device=device) for _ in range(slots)]
    self.gathered = [torch.empty((self.world * count, w), dtype=torch.float64, device=device) for _ in range(slots)]
    self.merged = [torch.empty((count, w), dtype=torch.float64, device=device) for _ in range(slots)]
    self.host = [torch.empty((count, w), dtype=torch.float64).pin_memory() for _ in range(slots)]
    self.done = [torch.cuda.Event() for _ in range(slots)]
    self.slots = slots

  def step(self, slot: int, xs, acq, index_base: int, score_out=None) -> None:
    """Enqueue score -> local top-k -> all-gather -> merge -> D2H for one pool shard (asynchronous)."""
    import torch
    dev = self.dev
    dev.score_topk_pack(xs, acq, self.count, index_base, self.payload[slot], score_out=score_out)
    if self.world > 1:
      with torch.cuda.stream(dev.stream):   # NCCL orders itself against the handle's stream
        self.dist.all_gather_into_tensor(self.gathered[slot], self.payload[slot])
      rows = self.gathered[slot]
    else:
      rows = self.payload[slot]
    dev.merge_topk(rows, self.count, self.merged[slot], self.host[slot])
    self.done[slot].record(dev.stream)

  def result(self, slot: int):
    """(global indices [count] i64, scores [count], features [count, D]) of the step last enqueued in `slot`."""
    self.done[slot].synchronize()
    g = self.host[slot].numpy()
    return g[:, 1].astype(np.int64), g[:, 0].copy(), g[:, 2:].copy()


def global_topk(dist, idx: np.ndarray, val: np.ndarray, x, count: int):
  """Host-side variant (CPU tensors / gloo): all-gathers each rank's local top-`count` (global indices,
  scores, feature rows) and merges with `merge_topk`.  Returns host arrays."""
  import torch
  world = dist.get_world_size()
  x = x.cpu() if hasattr(x, 'cpu') else torch.as_tensor(x)
  d = x.shape[1]
  payload = torch.empty((count, d + 2), dtype=torch.float64)
  payload[:, 0] = torch.from_numpy(np.asarray(val, np.float64))
  payload[:, 1] = torch.from_numpy(np.asarray(idx, np.float64))
  payload[:, 2:] = x
  gathered = torch.empty((world * count, d + 2), dtype=torch.float64)
  dist.all_gather_into_tensor(gathered, payload)
  g = gathered.numpy()
  return merge_topk(g[:, 1].astype(np.int64), g[:, 0], g[:, 2:], count)
