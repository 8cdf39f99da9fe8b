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


class PeerExchange:
  """libvzgp's fused peer-memory exchange (`vzgp_exchange`, csrc/exchange.cu) for one rank.

  `transport`:
    'peer'  one kernel launch per step: NVLink stores into every peer's buffer + release/acquire flags +
            the deterministic merge (CUDA IPC mappings across processes),
    'nccl'  the in-library fallback: ncclAllGather on a communicator owned by the exchange + merge kernel.
  Creation is collective over `dist` (handle / unique-id exchange).  `dist=None` is a world of one."""

  def __init__(self, dist, gp_dev, count: int, width: int, transport: str = 'peer'):
    import ctypes as C
    import torch
    from vizier_b200 import _lib
    self._lib = _lib.load()
    self.dev, self.count, self.width, self.transport = gp_dev, count, width, transport
    self.world = dist.get_world_size() if dist is not None else 1
    self.rank = dist.get_rank() if dist is not None else 0
    x = C.c_void_p()
    _lib.check('vzgp_exchange_create', self._lib.vzgp_exchange_create(gp_dev._h, self.rank, self.world, count, width, C.byref(x)))
    self._x = x
    if self.world > 1:
      if transport == 'peer':
        mine = (C.c_ubyte * 64)()
        _lib.check('vzgp_exchange_ipc_handle', self._lib.vzgp_exchange_ipc_handle(x, mine))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine))
        blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b''.join(handles))
        _lib.check('vzgp_exchange_open', self._lib.vzgp_exchange_open(x, blob))
      elif transport == 'nccl':
        ident = [None]
        if self.rank == 0:
          buf = (C.c_ubyte * 128)()
          _lib.check('vzgp_nccl_unique_id', self._lib.vzgp_nccl_unique_id(buf))
          ident[0] = bytes(buf)
        dist.broadcast_object_list(ident, src=0)
        buf = (C.c_ubyte * 128).from_buffer_copy(ident[0])
        _lib.check('vzgp_exchange_nccl_init', self._lib.vzgp_exchange_nccl_init(x, buf))
      else:
        raise ValueError(transport)
      dist.barrier()   # every rank has mapped every buffer before the first step publishes into them

  @classmethod
  def local_group(cls, gp_devs, count: int, width: int):
    """Exchanges for several handles driven by ONE process (rank r = gp_devs[r]; the handles may sit on
    different GPUs or, for tests, on different streams of one GPU): peers are mapped by plain pointers."""
    import ctypes as C
    from vizier_b200 import _lib
    lib = _lib.load()
    world = len(gp_devs)
    group = []
    for r, d in enumerate(gp_devs):
      self = cls.__new__(cls)
      self._lib, self.dev, self.count, self.width, self.transport = lib, d, count, width, 'peer'
      self.world, self.rank = world, r
      x = C.c_void_p()
      _lib.check('vzgp_exchange_create', lib.vzgp_exchange_create(d._h, r, world, count, width, C.byref(x)))
      self._x = x
      group.append(self)
    bases = (C.c_void_p * world)(*[lib.vzgp_exchange_base(g._x) for g in group])
    for g in group:
      _lib.check('vzgp_exchange_set_peers', lib.vzgp_exchange_set_peers(g._x, bases))
    return group

  def allgather_topk(self, payload, out, host_out=None) -> None:
    """payload [count, width] device rows of this rank -> out [count, width] merged rows (same on every
    rank); asynchronous on the GP handle's stream."""
    import ctypes as C
    from vizier_b200 import _lib
    _lib.check('vzgp_allgather_topk', self._lib.vzgp_allgather_topk(
        self.dev._h, self._x, C.c_void_p(payload.data_ptr()), C.c_void_p(out.data_ptr()),
        C.c_void_p(host_out.data_ptr()) if host_out is not None else None,
        1 if (self.transport == 'nccl' and self.world > 1) else 0))

  def status(self) -> int:
    """Synchronises; 1 if some fused step timed out waiting for a peer."""
    import ctypes as C
    from vizier_b200 import _lib
    st = C.c_int(0)
    _lib.check('vzgp_exchange_status', self._lib.vzgp_exchange_status(self.dev._h, self._x, C.byref(st)))
    return int(st.value)

  def close(self):
    if getattr(self, '_x', None):
      self._lib.vzgp_exchange_destroy(self._x)
      self._x = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


class TopkExchange:
  """The per-suggest collective, entirely on the GP handle's stream: `DeviceGP.score_topk_pack` (fused
  score -> device top-k -> rows [score, global index, features]) then `vzgp_allgather_topk` - by default
  ONE fused kernel that pushes the rows to every peer over NVLink, waits for theirs and merges (every rank
  computes the identical result) - then an asynchronous copy of the merged rows to pinned host memory.
  Nothing here blocks the host and no other process's host thread is involved, so consecutive steps
  queue back to back; `result(slot)` waits for one step's event.

  `transport`: 'peer' (default; env VZGP_EXCHANGE overrides), 'nccl' (in-library ncclAllGather + merge
  kernel) or 'torch' (round 1: torch.distributed all_gather_into_tensor issued from Python + merge).
  `slots` result buffers: the host may lag `slots - 1` steps behind the device before it has to wait.
  """

  def __init__(self, dist, gp_dev, dim: int, count: int, slots: int = 8, transport: str | None = None):
    import os
    import torch
    self.dist, self.dev, self.count, self.dim = dist, gp_dev, count, dim
    self.world = dist.get_world_size() if dist is not None else 1
    self.transport = transport or os.environ.get('VZGP_EXCHANGE', 'peer')
    device = gp_dev.device
    w = dim + 2
    self.payload = [torch.empty((count, w), dtype=torch.float64, device=device) for _ in range(slots)]
    self.merged = [torch.empty((count, w), dtype=torch.float64, device=device) for _ in range(slots)]
    self.host = [torch.empty((count, w), dtype=torch.float64).pin_memory() for _ in range(slots)]
    self.done = [torch.cuda.Event() for _ in range(slots)]
    self.slots = slots
    if self.transport == 'torch':
      self.gathered = [torch.empty((self.world * count, w), dtype=torch.float64, device=device) for _ in range(slots)]
      self.peer = None
    else:
      self.peer = PeerExchange(dist, gp_dev, count, w, self.transport)

  def step(self, slot: int, xs, acq, index_base: int, score_out=None) -> None:
    """Enqueue score -> local top-k -> exchange + merge -> D2H for one pool shard (asynchronous)."""
    import torch
    dev = self.dev
    dev.score_topk_pack(xs, acq, self.count, index_base, self.payload[slot], score_out=score_out)
    if self.peer is not None:
      self.peer.allgather_topk(self.payload[slot], self.merged[slot], self.host[slot])
    else:
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

  def suggest_host(self, host_x, acq, index_base: int, host_scores=None):
    """The sharded suggest END TO END from host memory (what bench.py's e2e times): one synchronous
    `vzgp_suggest_host` call - pinned candidates -> device, score, local top-k, exchange + merge, winner
    rows (and, if `host_scores` is given, this shard's scores) back to the host."""
    if self.peer is None:
      raise NotImplementedError("suggest_host needs the in-library transports ('peer' or 'nccl')")
    return self.dev.suggest_host(host_x, acq, self.count, index_base, exchange=self.peer, score_out=host_scores)


def global_topk(dist, idx: np.ndarray, val: np.ndarray, x, count: int, group=None):
  """Host-side variant (CPU tensors / gloo): all-gathers each rank's local top-`count` (global indices,
  scores, feature rows) and merges with `merge_topk`.  Returns host arrays."""
  import torch
  world = dist.get_world_size(group)
  x = x.cpu() if hasattr(x, 'cpu') else torch.as_tensor(x)
  d = x.shape[1]
  payload = torch.empty((count, d + 2), dtype=torch.float64)
  payload[:, 0] = torch.from_numpy(np.asarray(val, np.float64))
  payload[:, 1] = torch.from_numpy(np.asarray(idx, np.float64))
  payload[:, 2:] = x
  gathered = torch.empty((world * count, d + 2), dtype=torch.float64)
  dist.all_gather_into_tensor(gathered, payload, group=group)
  g = gathered.numpy()
  return merge_topk(g[:, 1].astype(np.int64), g[:, 0], g[:, 2:], count)
