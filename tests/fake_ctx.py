"""TEST INFRASTRUCTURE: a NumPy stand-in for `_lib.Context` that lets the CPU
suite drive the Python operator layer end to end (evaluate_in_memory ->
_metric_and_region_loop -> metrics.batch -> _spatial.run_* -> C ABI call)
without a GPU.

It interprets the SAME raw arguments the operators hand to the C ABI -- base
addresses, element-offset tables, the wb2_weights factorisation, member strides
-- by reading host memory at those addresses, and produces the raw statistics
`include/wb2b200.h` documents.  What it checks is therefore the host logic
(offset tables, gathers, region factors, result assembly); the arithmetic of
the CUDA kernels is checked by the `-m gpu` tests against the oracle.  It is
never importable from the package: the product has no CPU path.
"""
from __future__ import annotations

import contextlib
import ctypes as C

import numpy as np

from weatherbench2_b200 import _lib


def _view(addr: int, n: int, dtype) -> np.ndarray:
  dtype = np.dtype(dtype)
  buf = (C.c_char * (n * dtype.itemsize)).from_address(int(addr))
  return np.frombuffer(buf, dtype=dtype, count=n)


def _slab(base: int, off: int, w: _lib.WeightSpec, dtype) -> np.ndarray:
  es = np.dtype(dtype).itemsize
  span = (w.nrow - 1) * w.row_stride + w.ncol
  flat = _view(base + int(off) * es, span, dtype)
  rows = np.lib.stride_tricks.as_strided(
      flat, shape=(w.nrow, w.ncol), strides=(w.row_stride * es, es),
      writeable=False)
  return rows.astype(np.float64)


class FakeContext:
  """'Device' memory is host memory; compute entries are NumPy."""

  def __init__(self):
    self._bufs: dict = {}
    self.calls: list = []
    self.h2d_bytes = 0

  # -- memory -------------------------------------------------------------------
  def malloc(self, nbytes: int) -> int:
    buf = np.zeros(max(int(nbytes), 8), dtype=np.uint8)
    self._bufs[buf.ctypes.data] = buf
    return buf.ctypes.data

  def free(self, ptr) -> None:
    self._bufs.pop(int(ptr), None)

  def to_device(self, arr: np.ndarray) -> int:
    arr = np.ascontiguousarray(arr)
    ptr = self.malloc(arr.nbytes)
    _view(ptr, arr.nbytes, np.uint8)[...] = arr.view(np.uint8).reshape(-1)
    self.h2d_bytes += arr.nbytes
    return ptr

  def from_device(self, ptr, shape, dtype) -> np.ndarray:
    n = int(np.prod(shape)) if len(shape) else 1
    return _view(ptr, n, dtype).reshape(shape).copy()

  def synchronize(self) -> None:
    pass

  def slab_cache(self, nbytes=None):
    return contextlib.nullcontext()

  def pinned_result(self, shape, dtype):
    return np.empty(shape, dtype=dtype)

  # -- weights ------------------------------------------------------------------
  def _weights(self, w: _lib.WeightSpec) -> np.ndarray:
    """W[r, row, col] exactly as the header defines it."""
    seg_of_col = np.zeros(w.ncol, dtype=int)
    for k in range(w.nseg):
      seg_of_col[w.seg_start[k]:w.seg_start[k + 1]] = k
    colw = (np.ones(w.ncol) if w.col_w is None
            else w.col_w.astype(np.float64))
    out = (w.row_w[:, :, None] * w.seg_w[:, seg_of_col][:, None, :] *
           colw[None, None, :])
    if w.cell_w_dev:
      cell = _view(w.cell_w_dev, w.nrow * w.ncol, np.float32).reshape(
          w.nrow, w.ncol).astype(np.float64)
      out = out * cell[None]
    return out

  @staticmethod
  def _wsum(W, value, skipna, zero_skip):
    """(sum W*value, sum W*[valid]) with the skipna / where(w>0, 0) rules."""
    if zero_skip:
      value = np.where(W > 0, value, 0.0)
    valid = ~np.isnan(value)
    if skipna:
      return (np.where(valid, value, 0.0) * W).sum(), (W * valid).sum()
    return (value * W).sum(), W.sum()

  # -- K1 -----------------------------------------------------------------------
  def det_metrics(self, f, t, c, dtype, off_f, off_t, off_c, weights, skipna,
                  out, host=False):
    self.calls.append(('det_metrics', int(off_f.size), weights.nregion, host))
    dt = np.float32 if dtype == _lib.F32 else np.float64
    W = self._weights(weights)
    res = np.zeros((off_f.size, weights.nregion, _lib.DET_NSTAT))
    for i in range(off_f.size):
      fs = _slab(f, off_f[i], weights, dt)
      ts = _slab(t, off_t[i], weights, dt)
      d = fs - ts
      vals = [d * d, np.abs(d), d]
      if c:
        cs = _slab(c, off_c[i], weights, dt)
        fa, ta = fs - cs, ts - cs
        vals += [fa * ta, fa * fa, ta * ta]
      for r in range(weights.nregion):
        zs = weights.zero_skip
        for k, v in enumerate(vals):
          s, ws = self._wsum(W[r], v, skipna, zs)
          res[i, r, k] = s
          if k == 0:
            res[i, r, 6] = ws
          elif k == 3:
            res[i, r, 7] = ws
          elif k == 4:
            res[i, r, 8] = ws
          elif k == 5:
            res[i, r, 9] = ws
    _view(out, res.size, np.float64)[...] = res.reshape(-1)

  # -- K2 -----------------------------------------------------------------------
  def ens_metrics(self, x, t, dtype, nmember, member_stride, off_x, off_t,
                  weights, skipna, out):
    self.calls.append(('ens_metrics', int(off_x.size), weights.nregion))
    dt = np.float32 if dtype == _lib.F32 else np.float64
    es = np.dtype(dt).itemsize
    W = self._weights(weights)
    M = int(nmember)
    res = np.zeros((off_x.size, weights.nregion, _lib.ENS_NSTAT))
    mean_fn = np.nanmean if skipna else np.mean
    for i in range(off_x.size):
      xs = np.stack([_slab(x + m * member_stride * es, off_x[i], weights, dt)
                     for m in range(M)])
      ts = _slab(t, off_t[i], weights, dt)
      with np.errstate(invalid='ignore'), _quiet():
        skill = mean_fn(np.abs(ts[None] - xs), axis=0)
        if M < 2:
          spread = np.zeros_like(ts)
        else:
          order = np.sort(xs, axis=0)  # NaN last, like np.argsort
          if skipna:
            n = (~np.isnan(xs)).sum(axis=0).astype(np.float64)
            rank = np.arange(1, M + 1, dtype=np.float64)[:, None, None]
            spread = 2.0 * np.nansum((2 * rank - M - 1) * order, axis=0) / np.where(
                n > 0, n, np.nan) / (M - 1)
          else:
            rank = np.arange(1, M + 1, dtype=np.float64)[:, None, None]
            spread = 2.0 * ((2 * rank - M - 1) * order).mean(axis=0) / (M - 1)
        xbar = mean_fn(xs, axis=0)
        mse = (ts - xbar) ** 2
        if M > 1:
          var = (np.nanvar if skipna else np.var)(xs, axis=0, ddof=1)
        else:
          var = np.full_like(ts, np.nan)
        deb = mse - var / M
      for r in range(weights.nregion):
        for k, v in enumerate([skill, spread, mse, var, deb]):
          s, ws = self._wsum(W[r], v, skipna, weights.zero_skip)
          res[i, r, k] = s
          res[i, r, 5 + k] = ws
    _view(out, res.size, np.float64)[...] = res.reshape(-1)

  def ens_metrics_host(self, x, t, nmember, member_stride, off_x, off_t,
                       weights, skipna, out):
    self.ens_metrics(x, t, _lib.F32, nmember, member_stride, off_x, off_t,
                     weights, skipna, out)

  def __getattr__(self, name):
    raise AttributeError(
        f'FakeContext has no emulation of {name!r}: this code path needs the '
        'GPU tests')


@contextlib.contextmanager
def _quiet():
  import warnings  # pylint: disable=import-outside-toplevel
  with warnings.catch_warnings():
    warnings.simplefilter('ignore', RuntimeWarning)
    yield


@contextlib.contextmanager
def installed():
  """Makes `_lib.default_context()` return a FakeContext inside the block."""
  fake = FakeContext()
  saved = _lib.default_context
  _lib.default_context = lambda device=None: fake
  try:
    yield fake
  finally:
    _lib.default_context = saved
