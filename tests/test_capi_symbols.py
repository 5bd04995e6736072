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


def _declarations():
  """{name: (return type, [parameter types])} parsed from the header, types
  normalised to a small vocabulary."""
  text = open(os.path.join(ROOT, 'include', 'wb2b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  text = re.sub(r'//[^\n]*', '', text)
  out = {}
  for ret, name, params in re.findall(
      r'\n\s*([A-Za-z_][\w\s\*]*?)\s*\b(wb2_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;',
      text):
    plist = []
    params = ' '.join(params.split())
    if params not in ('', 'void'):
      for p in params.split(','):
        p = p.strip()
        if '*' in p:
          base = p[:p.rindex('*') + 1]
        else:
          base = ' '.join(p.split()[:-1])  # drop the parameter name
        plist.append(' '.join(base.replace('*', ' * ').split()))
    out[name] = (' '.join(ret.replace('*', ' * ').split()), plist)
  return out


def _ctype_class(c_type: str):
  """The ctypes type a C parameter type must be bound as."""
  import ctypes as C
  from weatherbench2_b200 import _lib
  t = c_type.replace('const ', '').strip()
  scalars = {'int': C.c_int, 'int32_t': C.c_int32, 'int64_t': C.c_int64,
             'uint64_t': C.c_uint64, 'size_t': C.c_size_t, 'float': C.c_float,
             'double': C.c_double}
  if t in scalars:
    return [scalars[t]]
  if t.endswith('* *'):                      # out-parameters: void** / ctx**
    return [C.POINTER(C.c_void_p)]
  if t == 'char *':
    return [C.c_char_p]
  if t in ('wb2_weights *',):
    return [C.POINTER(_lib.Weights)]
  if t in ('wb2_csr *',):
    return [C.POINTER(_lib.Csr)]
  typed = {'int64_t *': C.POINTER(C.c_int64), 'int32_t *': C.POINTER(C.c_int32),
           'double *': C.POINTER(C.c_double), 'float *': C.POINTER(C.c_float)}
  # data buffers are passed as raw addresses (c_void_p); descriptor arrays may
  # be bound as typed pointers
  if t in typed:
    return [C.c_void_p, typed[t]]
  if t.endswith('*'):
    return [C.c_void_p]
  raise AssertionError(f'unmapped C type {c_type!r}')


def test_ctypes_prototypes_match_the_header_declarations():
  """Every argument of every entry point: same count, and a ctypes type of the
  right width / kind (an int32 bound as int64, or a missing parameter, would
  corrupt the call silently)."""
  import ctypes as C
  from weatherbench2_b200 import _lib
  decls = _declarations()
  assert set(decls) == set(_lib.PROTOTYPES)
  for name, (ret, params) in decls.items():
    restype, argtypes = _lib.PROTOTYPES[name]
    assert len(argtypes) == len(params), (name, params, argtypes)
    for i, (c_type, bound) in enumerate(zip(params, argtypes)):
      allowed = _ctype_class(c_type)
      assert any(bound is a or (
          C.sizeof(bound) == C.sizeof(a) and bound in (C.c_int, C.c_int32)
          and a in (C.c_int, C.c_int32)) for a in allowed), (
              name, i, c_type, bound)
    want = {'int': C.c_int, 'int64_t': C.c_int64, 'const char *': C.c_char_p,
            'void *': C.c_void_p}[ret]
    assert restype is want or (restype in (C.c_int, C.c_int32) and
                               want in (C.c_int, C.c_int32)), (name, ret)


def test_ctypes_structures_match_the_header_typedefs():
  """wb2_weights / wb2_csr: same member names, order and scalar widths."""
  import ctypes as C
  from weatherbench2_b200 import _lib
  text = open(os.path.join(ROOT, 'include', 'wb2b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  width = {'int32_t': 4, 'int64_t': 8, 'uint8_t': 1, 'float': 4, 'double': 8}
  for cname, struct in (('wb2_weights', _lib.Weights), ('wb2_csr', _lib.Csr)):
    body = re.search(r'typedef struct \{([^{}]*)\}\s*' + cname + ';', text,
                     re.S).group(1)
    members = []
    for decl in body.split(';'):
      decl = ' '.join(decl.split())
      if decl:
        name = re.search(r'(\w+)$', decl).group(1)
        members.append((name, decl[:-len(name)].strip()))
    assert [m[0] for m in members] == [f[0] for f in struct._fields_], cname
    for (name, c_type), (_, bound) in zip(members, struct._fields_):
      if '*' in c_type:
        assert C.sizeof(bound) == C.sizeof(C.c_void_p), (cname, name)
        target = c_type.replace('const', '').replace('*', '').strip()
        if hasattr(bound, '_type_') and not isinstance(bound._type_, str):
          assert C.sizeof(bound._type_) == width[target], (cname, name)
      else:
        assert C.sizeof(bound) == width[c_type], (cname, name)
