"""CPU-only: the C-ABI library loads and exports every symbol that
include/wb2b200.h declares; the ctypes prototypes cover the same set; creating
a context without a GPU fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  text = open(os.path.join(ROOT, 'include', 'wb2b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(wb2_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_hot_path_entry_points():
  syms = _declared_symbols()
  for name in ('wb2_det_metrics', 'wb2_det_metrics_host', 'wb2_ens_metrics',
               'wb2_regrid_conservative', 'wb2_zonal_spectrum',
               'wb2_last_error', 'wb2_version'):
    assert name in syms


def test_library_exports_every_declared_symbol():
  from weatherbench2_b200 import _lib
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for name in _declared_symbols():
    assert hasattr(lib, name), f'{name} missing from libwb2b200.so'
  assert set(_lib.PROTOTYPES) == set(_declared_symbols())
  assert _lib.load_library().wb2_version() == 100
  assert _lib.load_library().wb2_has_cuda() == 1


def test_no_cpu_fallback_without_gpu():
  import torch
  from weatherbench2_b200 import _lib
  if torch.cuda.is_available():
    pytest.skip('a GPU is present')
  with pytest.raises(_lib.Wb2Error):
    _lib.Context(0)
  from weatherbench2_b200 import metrics, xarray_lite as xl
  import numpy as np
  ds = xl.Dataset({'a': (('time', 'latitude', 'longitude'),
                         np.zeros((1, 3, 4), np.float32))},
                  {'latitude': np.array([-45., 0., 45.]),
                   'longitude': np.arange(4) * 90., 'time': np.arange(1)})
  with pytest.raises(_lib.Wb2Error):
    metrics.MSE().compute_chunk(ds, ds)


def test_product_never_imports_the_oracle():
  pkg = os.path.join(ROOT, 'weatherbench2_b200')
  for dirpath, _, files in os.walk(pkg):
    for fn in files:
      if fn.endswith(('.py', '.cu', '.cuh', '.h')):
        src = open(os.path.join(dirpath, fn)).read()
        assert 'import oracle' not in src and 'from oracle' not in src, fn
