"""Minimal runtime profiler with the reference's hook names (vizier/utils/profiler.py:138-260):
`record_runtime` stores per-scope durations that `collect_events()` exposes, so callers that read
e.g. the 'VizierGPBandit.suggest' key keep working."""

from __future__ import annotations

import contextlib
import functools
import time
from typing import Dict, List

_events: List[Dict[str, float]] = []


@contextlib.contextmanager
def collect_events():
  store: Dict[str, List[float]] = {}
  _events.append(store)
  try:
    yield store
  finally:
    _events.remove(store)


def _record(name: str, seconds: float):
  for store in _events:
    store.setdefault(name, []).append(seconds)


@contextlib.contextmanager
def timeit(name: str):
  t0 = time.perf_counter()
  try:
    yield
  finally:
    _record(name, time.perf_counter() - t0)


def record_runtime(fn):
  name = fn.__qualname__

  @functools.wraps(fn)
  def wrapper(*args, **kwargs):
    t0 = time.perf_counter()
    try:
      return fn(*args, **kwargs)
    finally:
      _record(name, time.perf_counter() - t0)

  return wrapper
