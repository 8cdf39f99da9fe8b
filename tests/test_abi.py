"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/vzgp.h declares (no compute calls: there is no GPU in the build container)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
  text = open(os.path.join(ROOT, 'include', 'vzgp.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(vzgp_[a-z_0-9]+)\s*\(', text)))


def test_header_declares_expected_entry_points():
  syms = _header_symbols()
  for name in ['vzgp_create', 'vzgp_destroy', 'vzgp_kernel_matrix', 'vzgp_cross_kernel',
               'vzgp_cholesky_retry', 'vzgp_fit', 'vzgp_nll_grad', 'vzgp_score', 'vzgp_score_host',
               'vzgp_topk', 'vzgp_eagle_run', 'vzgp_random_search']:
    assert name in syms


def test_library_exports_every_declared_symbol():
  from vizier_b200 import _lib
  if not os.path.exists(_lib.LIB_PATH):
    import __graft_entry__ as g
    g.build()
  lib = _lib.load()
  syms = _header_symbols()
  assert sorted(_lib.SIGNATURES) == syms, 'ctypes table and header disagree'
  for s in syms:
    assert hasattr(lib, s), f'{s} not exported by libvzgp.so'
  assert lib.vzgp_version() == 3
  assert isinstance(lib.vzgp_device_count(), int)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
  from vizier_b200 import _lib
  monkeypatch.setattr(_lib, '_lib', None)
  monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
  with pytest.raises(ImportError):
    _lib.load()


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, 'vizier_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith('.py'):
        src = open(os.path.join(dirpath, f)).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
