"""Philox4x32-10 known-answer tests (Random123 kat_vectors) for the shared RNG."""
import numpy as np

from oracle import eagle_oracle as eo


def _kat(ctr, key):
  out = eo.philox4x32(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
  return [int(v) for v in out]


def test_philox_kat_zero():
  assert _kat([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_philox_kat_ones():
  f = 0xFFFFFFFF
  assert _kat([f, f, f, f], [f, f]) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_philox_kat_pi():
  assert _kat([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
      0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_uniform_range_and_moments():
  u = eo.philox_uniform(1234, eo.STREAM_INIT_POOL, 0, 200000)
  assert u.min() >= 0.0 and u.max() < 1.0
  assert abs(u.mean() - 0.5) < 5e-3
  assert abs(u.var() - 1 / 12) < 2e-3
  # different stream / iteration give different draws
  assert not np.array_equal(u[:16], eo.philox_uniform(1234, eo.STREAM_TRIM, 0, 16))
  assert not np.array_equal(u[:16], eo.philox_uniform(1234, eo.STREAM_INIT_POOL, 1, 16))
