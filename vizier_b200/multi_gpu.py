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


def global_topk(dist, idx: np.ndarray, val: np.ndarray, x, count: int):
  """All-gathers each rank's local top-`count` (global indices, scores, feature rows) and merges.

  idx/val: host arrays [count]; x: torch device tensor [count, D].  Returns host arrays.
  """
  import torch
  world = dist.get_world_size()
  d = x.shape[1]
  payload = torch.empty((count, d + 2), dtype=torch.float64, device=x.device)
  payload[:, 0] = torch.from_numpy(np.asarray(val, np.float64)).to(x.device)
  # global indices < 2^53 are exact in fp64
  payload[:, 1] = torch.from_numpy(np.asarray(idx, np.float64)).to(x.device)
  payload[:, 2:] = x
  gathered = torch.empty((world * count, d + 2), dtype=torch.float64, device=x.device)
  dist.all_gather_into_tensor(gathered, payload)
  g = gathered.cpu().numpy()
  return merge_topk(g[:, 1].astype(np.int64), g[:, 0], g[:, 2:], count)
