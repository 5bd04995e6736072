"""`jax.numpy` stand-in: NumPy functions with 32-bit results (see __init__)."""
import functools as _functools

import numpy as _np

newaxis = None
nan = _np.float32(_np.nan)
pi = _np.pi
inf = _np.inf


def _narrow(x):
  if isinstance(x, _np.ndarray):
    if x.dtype == _np.float64:
      return x.astype(_np.float32)
    if x.dtype == _np.int64:
      return x.astype(_np.int32)
  elif isinstance(x, _np.float64):
    return _np.float32(x)
  elif isinstance(x, (tuple, list)):
    return type(x)(_narrow(v) for v in x)
  return x


def _wrap(fn):
  @_functools.wraps(fn)
  def call(*args, **kwargs):
    kwargs.pop('precision', None)
    return _narrow(fn(*_narrow(list(args)),
                      **{k: _narrow(v) for k, v in kwargs.items()}))
  return call


def vectorize(pyfunc, *, signature=None, **kwargs):
  return _wrap(_np.vectorize(pyfunc, signature=signature, **kwargs))


def array(x, dtype=None):
  return _narrow(_np.array(x, dtype=dtype))


asarray = array


def __getattr__(name):
  return _wrap(getattr(_np, name))
