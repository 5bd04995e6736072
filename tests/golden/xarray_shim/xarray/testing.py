"""xarray.testing for the stand-in: values, dimension names (in order) and the
labels of dimension coordinates must agree."""
import numpy as np

from weatherbench2_b200.xarray_lite import DataArray, Dataset


def _pairs(a, b):
  assert type(a) is type(b) or (isinstance(a, (DataArray, Dataset)) and
                                isinstance(b, (DataArray, Dataset))), (
                                    type(a), type(b))
  if isinstance(a, Dataset):
    assert set(a.keys()) == set(b.keys()), (sorted(a.keys()), sorted(b.keys()))
    return [(k, a[k], b[k]) for k in a.keys()]
  return [(getattr(a, 'name', None), a, b)]


def _check_coords(name, x, y, rtol=1e-7, atol=0.0):
  assert x.dims == y.dims, (name, x.dims, y.dims)
  for d in x.dims:
    if d in x.coords and d in y.coords:
      cx, cy = np.asarray(x.coords[d].values), np.asarray(y.coords[d].values)
      if cx.dtype.kind in 'fc':
        np.testing.assert_allclose(cx, cy, rtol=rtol, atol=atol,
                                   err_msg=f'{name}: coordinate {d}')
      else:
        np.testing.assert_array_equal(cx, cy,
                                      err_msg=f'{name}: coordinate {d}')


def assert_allclose(a, b, rtol=1e-5, atol=1e-8, decode_bytes=True):
  del decode_bytes
  for name, x, y in _pairs(a, b):
    _check_coords(name, x, y, rtol, atol)  # xarray applies them to coords too
    np.testing.assert_allclose(np.asarray(x.values), np.asarray(y.values),
                               rtol=rtol, atol=atol, equal_nan=True,
                               err_msg=str(name))


def assert_equal(a, b):
  for name, x, y in _pairs(a, b):
    _check_coords(name, x, y)
    np.testing.assert_array_equal(np.asarray(x.values), np.asarray(y.values),
                                  err_msg=str(name))


assert_identical = assert_equal
