"""Stand-in `jax` for running weatherbench2/regridding.py in this container
(JAX is not installable here): `jax.numpy` is NumPy with JAX's default 32-bit
results (float64 -> float32, int64 -> int32 on the way in and out), `jit` is
the identity and `vmap` a Python loop.  Summation order inside einsum differs
from XLA's; comparisons against these vectors are tolerance-level (1e-5)."""
import numpy as _np

from . import numpy  # noqa: F401  pylint: disable=redefined-builtin

Array = _np.ndarray


def jit(fun=None, **kwargs):
  del kwargs
  return fun


def vmap(fun, in_axes=0, out_axes=0):

  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (
        (in_axes,) * len(args))
    n = next(_np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
    rows = []
    for i in range(n):
      rows.append(fun(*[a if ax is None else _np.take(a, i, axis=ax)
                        for a, ax in zip(args, axes)]))
    return numpy.stack(rows, axis=out_axes)

  return mapped
