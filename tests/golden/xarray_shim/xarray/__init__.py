"""Stand-in `xarray` used ONLY by tests/golden/make_reference_vectors.py to
execute the reference's own metrics.py / regions.py / thresholds.py / utils.py
in this container, where xarray itself cannot be installed (no network).

It is `weatherbench2_b200.xarray_lite` plus the rest of the xarray API those
four reference modules call (weighted means, the `.dt` accessor, apply_ufunc,
dot, cumsum, quantile, ...), each written to xarray's documented semantics.
The vectors it produces are therefore "the reference's code on a
re-implemented xarray subset": they pin the oracle's restatement of the
REFERENCE logic; xarray's own semantics remain restated (here and, separately,
in oracle/wb2_oracle.py, which does not use this module).
"""
import numpy as np
import pandas as pd

from weatherbench2_b200 import xarray_lite as _xl
from weatherbench2_b200.xarray_lite import (DataArray, Dataset, concat, merge,
                                            zeros_like)

__version__ = '0.0-shim'


# ---- free functions ----------------------------------------------------------
def where(cond, x, y):
  """xr.where: element-wise choice with broadcasting by dimension name."""
  base = next((o for o in (cond, x, y) if isinstance(o, (DataArray, Dataset))),
              None)
  if isinstance(base, Dataset) or any(isinstance(o, Dataset)
                                      for o in (cond, x, y)):
    ds = next(o for o in (cond, x, y) if isinstance(o, Dataset))
    out = Dataset(attrs=None)
    for k in ds.keys():
      pick = lambda o: o[k] if isinstance(o, Dataset) else o  # noqa: E731
      out[k] = where(pick(cond), pick(x), pick(y))
    return out
  cond = cond if isinstance(cond, DataArray) else DataArray(np.asarray(cond))
  zero = cond * 0  # carries dims / coords
  cx = (zero + x) if not isinstance(x, DataArray) else x
  cy = (zero + y) if not isinstance(y, DataArray) else y
  c, a = _xl._broadcast(cond, cx)[0:2], None  # pylint: disable=protected-access
  del c, a
  total = zero + cx * 0 + cy * 0  # the broadcast shape of all three
  cb = (total + cond).values != 0 if cond.dtype != bool else (
      (total + cond.astype(float)).values != 0)
  xb = (total * 0 + cx).values if np.isfinite(total.values).all() else None
  if xb is None:  # NaN / inf inside operands: broadcast without arithmetic
    xb = _bcast(cx, total)
    yb = _bcast(cy, total)
    cb = _bcast(cond.astype(float), total) != 0
  else:
    yb = (total * 0 + cy).values
  return DataArray(np.where(cb, xb, yb), total.dims, total.coords)


def _bcast(da, like):
  perm = [da.dims.index(d) for d in like.dims if d in da.dims]
  v = np.transpose(da.values, perm)
  v = v[tuple(slice(None) if d in da.dims else None for d in like.dims)]
  return np.broadcast_to(v, like.shape)


def apply_ufunc(func, *args, **kwargs):
  """Only the element-wise form the reference uses: func(array) -> array."""
  if kwargs:
    raise NotImplementedError(f'apply_ufunc kwargs {sorted(kwargs)}')
  first = args[0]
  if isinstance(first, Dataset):
    return first._map(lambda v: apply_ufunc(func, v, *args[1:]))  # pylint: disable=protected-access
  raw = [a.values if isinstance(a, DataArray) else a for a in args]
  return first._replace(np.asarray(func(*raw)))  # pylint: disable=protected-access


def dot(a, b, dims=None, dim=None):
  """xr.dot: sum of products over `dims` (NaN propagates)."""
  dims = dims if dims is not None else dim
  dims = (dims,) if isinstance(dims, str) else tuple(dims)
  return (a * b).sum(dims, skipna=False)


class _Weighted:
  """DatasetWeighted / DataArrayWeighted .mean (xarray/core/weighted.py):
  sum_of_weights uses the NOT-NULL mask of the data whatever `skipna`; the
  weighted sum fills NaN with 0 only when skipna; 0 weight-sum -> NaN."""

  def __init__(self, obj, weights):
    if np.isnan(np.asarray(weights.values, dtype=float)).any():
      raise ValueError('`weights` cannot contain missing values. Missing '
                       'values can be replaced by `weights.fillna(0)`.')
    self.obj, self.weights = obj, weights

  def _mean_da(self, da, dim, skipna):
    dim = tuple(d for d in dim if d in da.dims)
    w = self.weights
    if skipna or (skipna is None and da.dtype.kind in 'cfO'):
      data = da.fillna(0.0)
    else:
      data = da
    wsum = dot(data, w, dims=dim)
    mask = da.notnull()
    sow = dot(mask.astype(float) if hasattr(mask, 'astype') else mask, w,
              dims=dim)
    sow = sow.where(sow != 0.0)
    return wsum / sow

  def mean(self, dim=None, skipna=None, keep_attrs=None):
    del keep_attrs
    dim = (dim,) if isinstance(dim, str) else tuple(dim)
    if isinstance(self.obj, Dataset):
      return self.obj._map(lambda v: self._mean_da(v, dim, skipna))  # pylint: disable=protected-access
    return self._mean_da(self.obj, dim, skipna)


def _weighted(self, weights):
  return _Weighted(self, weights)


DataArray.weighted = _weighted
Dataset.weighted = _weighted


class _Dt:
  def __init__(self, da):
    self._da = da
    self._idx = pd.DatetimeIndex(np.asarray(da.values).ravel())

  def _field(self, name):
    v = np.asarray(getattr(self._idx, name)).reshape(self._da.shape)
    return self._da._replace(v)  # pylint: disable=protected-access

  dayofyear = property(lambda self: self._field('dayofyear'))
  hour = property(lambda self: self._field('hour'))
  year = property(lambda self: self._field('year'))


DataArray.dt = property(_Dt)

for _cls in (DataArray, Dataset):
  _cls.load = lambda self, **kw: self
  _cls.compute = lambda self, **kw: self
  _cls.chunk = lambda self, *a, **kw: self


# ---- methods xarray_lite does not need for the product ------------------------
def _ds_where(self, cond, other=np.nan):
  pick = lambda o, k: o[k] if isinstance(o, Dataset) else o  # noqa: E731
  out = Dataset(attrs=self.attrs)
  for k in self.keys():
    out[k] = self[k].where(pick(cond, k), pick(other, k))
  return out


def _da_where(self, cond, other=np.nan):
  """DataArray.where with broadcasting by name (cond may have fewer dims)."""
  if isinstance(cond, DataArray):
    return where(cond, self, other)
  return self._replace(np.where(np.asarray(cond), self.values, other))  # pylint: disable=protected-access


DataArray.where = _da_where
Dataset.where = _ds_where
Dataset.fillna = lambda self, value: self._map(lambda v: v.fillna(value))  # pylint: disable=protected-access
Dataset.isnull = lambda self: self._map(lambda v: v.isnull())  # pylint: disable=protected-access
Dataset.notnull = lambda self: self._map(lambda v: v.notnull())  # pylint: disable=protected-access
Dataset.astype = lambda self, dtype: self._map(lambda v: v.astype(dtype))  # pylint: disable=protected-access


def _ds_map(self, func, keep_attrs=None, args=(), **kwargs):
  del keep_attrs
  out = Dataset(attrs=self.attrs)
  for k in self.keys():
    out[k] = func(self[k], *args, **kwargs)
  return out


Dataset.map = _ds_map


def _moment(self, dim, skipna, ddof, root):
  dims = (dim,) if isinstance(dim, str) else tuple(dim)
  axes = tuple(self.dims.index(d) for d in dims)
  v = np.asarray(self.values, dtype=np.float64 if self.dtype.kind != 'f'
                 else self.dtype)
  nan_aware = skipna or (skipna is None and v.dtype.kind == 'f')
  import warnings
  with warnings.catch_warnings(), np.errstate(invalid='ignore',
                                              divide='ignore'):
    warnings.simplefilter('ignore', RuntimeWarning)
    r = (np.nanvar if nan_aware else np.var)(v, axis=axes, ddof=ddof)
  if root:
    r = np.sqrt(r)
  keep = tuple(d for d in self.dims if d not in dims)
  return self._replace(r, keep)  # pylint: disable=protected-access


DataArray.var = lambda self, dim=None, skipna=None, ddof=0, **kw: _moment(
    self, dim, skipna, ddof, False)
DataArray.std = lambda self, dim=None, skipna=None, ddof=0, **kw: _moment(
    self, dim, skipna, ddof, True)
Dataset.var = lambda self, dim=None, skipna=None, ddof=0, **kw: self._map(  # pylint: disable=protected-access
    lambda v: v.var(dim, skipna, ddof) if dim in v.dims else v)
Dataset.std = lambda self, dim=None, skipna=None, ddof=0, **kw: self._map(  # pylint: disable=protected-access
    lambda v: v.std(dim, skipna, ddof) if dim in v.dims else v)
Dataset.min = lambda self, dim=None, skipna=None, **kw: self._map(  # pylint: disable=protected-access
    lambda v: v.min(dim, skipna))
Dataset.max = lambda self, dim=None, skipna=None, **kw: self._map(  # pylint: disable=protected-access
    lambda v: v.max(dim, skipna))


def _da_cumsum(self, dim, skipna=None):
  del skipna
  return self._replace(np.cumsum(self.values, axis=self.dims.index(dim)))  # pylint: disable=protected-access


DataArray.cumsum = _da_cumsum
Dataset.cumsum = lambda self, dim, skipna=None: self._map(  # pylint: disable=protected-access
    lambda v: v.cumsum(dim) if dim in v.dims else v)


def _da_argmin(self, dim):
  ax = self.dims.index(dim)
  keep = tuple(d for d in self.dims if d != dim)
  return self._replace(np.argmin(self.values, axis=ax), keep)  # pylint: disable=protected-access


DataArray.argmin = _da_argmin


# ---- arithmetic aligns shared dimensions by label (join='inner') ----------------
_plain_binary = DataArray._binary  # pylint: disable=protected-access


def _align_inner(a, b):
  """xarray's default arithmetic join: along every shared dimension that both
  operands label, keep the labels present in both, in the order of `a`."""
  for d in a.dims:
    if d not in b.dims or d not in a.coords or d not in b.coords:
      continue
    if a.coords[d].dims != (d,) or b.coords[d].dims != (d,):
      continue  # a scalar coordinate of that name: nothing to align
    ca, cb = a.coords[d].values, b.coords[d].values
    if ca.shape == cb.shape and np.array_equal(ca, cb):
      continue
    keep = np.isin(ca, cb)
    a = a.isel({d: np.nonzero(keep)[0]})
    b = b.isel({d: _xl._lookup(cb, ca[keep])})  # pylint: disable=protected-access
  return a, b


def _aligned_binary(self, other, op, reflexive=False):
  if isinstance(other, DataArray):
    a, b = _align_inner(self, other)
    return _plain_binary(a, b, op, reflexive)
  return _plain_binary(self, other, op, reflexive)


DataArray._binary = _aligned_binary  # pylint: disable=protected-access

for _name, _op in (('__gt__', np.greater), ('__lt__', np.less),
                   ('__ge__', np.greater_equal), ('__le__', np.less_equal)):
  setattr(Dataset, _name,
          lambda self, o, _op=_op: self._binary(o, _op))  # pylint: disable=protected-access


# ---- .sel with several N-d indexers that share dimensions: pointwise -----------
_plain_da_sel = DataArray.sel


def _da_sel(self, indexers=None, method=None, drop=False, tolerance=None,
            **kw):
  idx = dict(indexers or {}, **kw)
  vec = {d: v for d, v in idx.items()
         if isinstance(v, DataArray) and d in self.dims and v.ndim >= 1 and
         v.dims != (d,)}
  rest = {d: v for d, v in idx.items() if d not in vec and d in self.dims}
  out = self
  if vec:
    maps, extra = {}, {}
    for d, lab in vec.items():
      pos = _xl._lookup(out.coords[d].values, lab.values.ravel(), method)  # pylint: disable=protected-access
      maps[d] = (lab.dims, pos.reshape(lab.shape))
      for k, c in lab.coords.items():
        extra.setdefault(k, c)
      extra[d] = _xl.Coord(lab.dims, lab.values)
    lazy = _xl.LazyGather(out, maps, extra_coords=extra)
    out = DataArray(lazy.values, lazy.dims, lazy.coords, out.name, out.attrs)
  for d, v in list(rest.items()):
    # pandas partial-string indexing on datetime coordinates: a string bound
    # (or label) covers its whole period
    if d in out.coords and out.coords[d].values.dtype.kind == 'M':
      def stamp(x, end):
        if not isinstance(x, str):
          return x
        period = pd.Period(x)
        return np.datetime64((period.end_time if end
                              else period.start_time).value, 'ns')
      if isinstance(v, slice):
        rest[d] = slice(stamp(v.start, False), stamp(v.stop, True))
      elif isinstance(v, str):
        rest[d] = slice(stamp(v, False), stamp(v, True))
  for d, v in list(rest.items()):
    # label slices go through pandas itself (Index.slice_indexer is what
    # xarray calls): both ends inclusive, bounds need not be labels
    if isinstance(v, slice) and d in out.coords:
      index = pd.Index(np.asarray(out.coords[d].values))
      if index.is_monotonic_increasing or index.is_monotonic_decreasing:
        plain = lambda b: (np.asarray(b.values)[()]  # noqa: E731
                           if isinstance(b, DataArray) else b)
        positions = index.slice_indexer(plain(v.start), plain(v.stop), v.step)
        out = out.isel({d: positions})
        del rest[d]
  if rest:
    if tolerance is not None:
      for d, v in rest.items():
        have = out.coords[d].values
        want = np.atleast_1d(np.asarray(v.values if isinstance(v, DataArray)
                                        else v, dtype=float))
        near = np.abs(have[None, :] - want[:, None]).min(axis=1)
        if (near > tolerance).any():
          raise KeyError(f'not all values found in index {d!r}')
    out = _plain_da_sel(out, rest, method=method, drop=drop)
  return out


DataArray.sel = _da_sel
Dataset.sel = lambda self, indexers=None, method=None, drop=False, \
    tolerance=None, **kw: self._map(lambda v: v.sel(  # pylint: disable=protected-access
        {d: k for d, k in dict(indexers or {}, **kw).items() if d in v.dims},
        method=method, drop=drop, tolerance=tolerance))


# ---- "zarr stores" and netCDF output are in-memory objects ---------------------
_STORE = {}


def register_store(path: str, dataset) -> None:
  _STORE[path] = dataset


def open_zarr(path, chunks=None, **kwargs):
  del chunks, kwargs
  return _STORE[path]


def _to_netcdf(self, path=None):
  import pickle
  payload = pickle.dumps({
      'vars': {k: (self[k].dims, np.asarray(self[k].values))
               for k in self.keys()},
      'coords': {k: (c.dims, np.asarray(c.values))
                 for k, c in self.coords.items()}})
  if path is None:
    return payload
  with open(path, 'wb') as f:
    f.write(payload)
  return None


Dataset.to_netcdf = _to_netcdf


# ---- the rest of what weatherbench2/evaluation.py and utils.py call ------------
DataArray.__eq__ = lambda self, o: self._binary(o, np.equal)  # pylint: disable=protected-access
DataArray.__ne__ = lambda self, o: self._binary(o, np.not_equal)  # pylint: disable=protected-access
DataArray.__hash__ = object.__hash__


def _thin(self, indexers=None, **kw):
  idx = dict(indexers or {}, **kw)
  return self.isel({d: slice(None, None, int(k)) for d, k in idx.items()})


DataArray.thin = _thin
Dataset.thin = _thin


def _da_swap_dims(self, mapping):
  """The coordinate `new` (which lies along `old`) becomes the dimension."""
  dims = tuple(mapping.get(d, d) for d in self.dims)
  coords = {}
  for k, c in self.coords.items():
    coords[k] = _xl.Coord(tuple(mapping.get(d, d) for d in c.dims), c.values,
                          c.attrs)
  return DataArray(self.data, dims, coords, self.name, self.attrs)


def _ds_swap_dims(self, mapping):
  out = Dataset(attrs=self.attrs)
  out._coords = {  # pylint: disable=protected-access
      k: _xl.Coord(tuple(mapping.get(d, d) for d in c.dims), np.asarray(c.values),
                   getattr(c, 'attrs', None))
      for k, c in self.coords.items()}
  for k in self.keys():
    out[k] = _da_swap_dims(self[k], mapping)
  return out


DataArray.swap_dims = _da_swap_dims
Dataset.swap_dims = _ds_swap_dims

_lite_concat = concat


def _reindex_like_union(objs, skip):
  """join='outer': every labelled dimension other than `skip` is extended to
  the sorted union of the labels, missing entries NaN."""
  dims = []
  for o in objs:
    for d in (o.dims if isinstance(o, DataArray) else o.sizes):
      if d != skip and d not in dims:
        dims.append(d)
  for d in dims:
    labels = [np.asarray(o.coords[d].values) for o in objs if d in o.coords]
    if len(labels) != len(objs) or all(
        l.shape == labels[0].shape and np.array_equal(l, labels[0])
        for l in labels):
      continue
    union = np.unique(np.concatenate(labels))
    objs = [_reindex(o, d, union) for o in objs]
  return objs


def _reindex(obj, dim, labels):
  if isinstance(obj, Dataset):
    out = Dataset(attrs=obj.attrs)
    for k in obj.keys():
      out[k] = _reindex(obj[k], dim, labels) if dim in obj[k].dims else obj[k]
    return out
  have = np.asarray(obj.coords[dim].values)
  ax = obj.dims.index(dim)
  shape = list(obj.shape)
  shape[ax] = labels.size
  v = np.asarray(obj.values)
  data = np.full(shape, np.nan, dtype=v.dtype if v.dtype.kind == 'f'
                 else np.float64)
  pos = np.searchsorted(labels, have)
  index = [slice(None)] * v.ndim
  index[ax] = pos
  data[tuple(index)] = v
  coords = {k: c for k, c in obj.coords.items() if dim not in c.dims}
  coords[dim] = _xl.Coord((dim,), labels)
  return DataArray(data, obj.dims, coords, obj.name, obj.attrs)


def concat(objs, dim, **kwargs):  # pylint: disable=function-redefined
  del kwargs
  objs = list(objs)
  if isinstance(dim, DataArray):
    name = dim.dims[0]
    objs = _reindex_like_union(objs, name)
    out = _lite_concat(objs, name)
    return out.assign_coords({name: np.asarray(dim.values)})
  return _lite_concat(_reindex_like_union(objs, dim), dim)


def _da_reindex(self, indexers=None, **kw):
  out = self
  for d, labels in dict(indexers or {}, **kw).items():
    labels = np.asarray(labels.values if isinstance(labels, DataArray)
                        else labels)
    out = out.isel({d: _xl._lookup(out.coords[d].values, labels)})  # pylint: disable=protected-access
  return out


DataArray.reindex = _da_reindex
Dataset.reindex = lambda self, indexers=None, **kw: self._map(  # pylint: disable=protected-access
    lambda v: v.reindex(indexers, **kw))

DataArray.all = lambda self, dim=None, axis=None, **kw: bool(
    np.all(self.values)) if dim is None and axis is None else self._replace(  # pylint: disable=protected-access
        np.all(self.values, axis=axis if axis is not None
               else self.dims.index(dim)),
        tuple(d for i, d in enumerate(self.dims)
              if i != (axis if axis is not None else self.dims.index(dim))))
DataArray.any = lambda self, dim=None, axis=None, **kw: bool(
    np.any(self.values))
DataArray.__bool__ = lambda self: bool(np.asarray(self.values))


# ---- what derived_variables.py (zonal spectrum, interpolation) and
# RankHistogram add ---------------------------------------------------------------
_elementwise_apply_ufunc = apply_ufunc


def apply_ufunc(func, *args, input_core_dims=None, output_core_dims=None,  # pylint: disable=function-redefined
                exclude_dims=None, **kwargs):
  """Element-wise form, or ONE core dimension moved last and allowed to change
  size (`exclude_dims`): what ZonalEnergySpectrum.compute asks for."""
  vectorize = kwargs.pop('vectorize', False)
  if input_core_dims is None:
    return _elementwise_apply_ufunc(func, *args, **kwargs)
  core, = input_core_dims
  out_core, = output_core_dims
  core, out_core = tuple(core), tuple(out_core)
  assert exclude_dims == set(core) and len(args) == 1 and not kwargs

  def one(da):
    if not all(d in da.dims for d in core):
      return da
    lead = tuple(d for d in da.dims if d not in core)
    moved = da.transpose(*lead, *core)
    v = np.asarray(moved.values)
    if vectorize:  # one call per slab, like np.vectorize with a signature
      flat = v.reshape((-1,) + v.shape[len(lead):])
      slabs = [np.asarray(func(flat[i])) for i in range(flat.shape[0])]
      data = np.stack(slabs).reshape(v.shape[:len(lead)] + slabs[0].shape)
    else:
      data = np.asarray(func(v))
    coords = {k: c for k, c in moved.coords.items()
              if not any(d in core for d in c.dims)}
    return DataArray(data, lead + out_core, coords, da.name, da.attrs)

  obj = args[0]
  return obj._map(one) if isinstance(obj, Dataset) else one(obj)  # pylint: disable=protected-access


def _rename_dims(self, mapping):
  if isinstance(self, Dataset):
    out = Dataset(attrs=self.attrs)
    out._coords = {k: _xl.Coord(tuple(mapping.get(d, d) for d in c.dims),  # pylint: disable=protected-access
                                np.asarray(c.values)) for k, c in
                   self.coords.items()}
    for k in self.keys():
      out[k] = _rename_dims(self[k], mapping)
    return out
  coords = {k: _xl.Coord(tuple(mapping.get(d, d) for d in c.dims), c.values)
            for k, c in self.coords.items()}
  return DataArray(self.data, tuple(mapping.get(d, d) for d in self.dims),
                   coords, self.name, self.attrs)


DataArray.rename_dims = _rename_dims
Dataset.rename_dims = _rename_dims

_lite_getattr = DataArray.__getattr__


def _da_getattr(self, name):
  try:
    return _lite_getattr(self, name)
  except AttributeError:
    dims = self.__dict__.get('dims', ())
    if name in dims:  # a dimension without coordinate reads as 0 .. n-1
      return DataArray(np.arange(self.shape[dims.index(name)]), (name,),
                       name=name)
    raise


DataArray.__getattr__ = _da_getattr


def _da_setitem(self, key, value):
  """da['coord'] = DataArray: (re)assigns a coordinate."""
  self._set_coord(key, value)  # pylint: disable=protected-access


DataArray.__setitem__ = _da_setitem
_lite_da_init = DataArray.__init__


def _da_init(self, data=None, coords=None, dims=None, name=None, attrs=None):
  """xarray's argument order (data, coords, dims, name, attrs); `dims` may be
  a single name; coordinates may be DataArrays."""
  if isinstance(coords, (tuple, list, str)) and not isinstance(dims, dict) and (
      dims is None or isinstance(dims, dict)):
    coords, dims = dims, coords  # called the xarray_lite way (data, dims, coords)
  if isinstance(dims, dict) and not isinstance(coords, dict):
    coords, dims = dims, coords
  if dims is None and not coords and np.ndim(data) > 0 and not isinstance(
      data, DataArray):
    dims = tuple(f'dim_{i}' for i in range(np.ndim(data)))  # xarray's default
  _lite_da_init(self, data, dims, coords, name, attrs)


DataArray.__init__ = _da_init


def _da_argsort(self, axis=-1, **kw):
  del kw
  return self._replace(np.argsort(self.values, axis=axis))  # pylint: disable=protected-access


DataArray.argsort = _da_argsort
_lite_expand_dims = DataArray.expand_dims


def _da_expand_dims(self, dim=None, axis=None, **kw):
  out = _lite_expand_dims(self, dim, **kw) if axis in (None, 0) else None
  if out is not None:
    return out
  out = _lite_expand_dims(self, dim, **kw)
  new = [d for d in out.dims if d not in self.dims]
  assert axis == -1
  return out.transpose(*self.dims, *new)


DataArray.expand_dims = _da_expand_dims


_outer_concat = concat


def _with_dim(obj, dim):
  """An object that holds `dim` only as a scalar coordinate becomes length 1
  along it (what xr.concat does before joining)."""
  if isinstance(obj, Dataset):
    out = Dataset(attrs=obj.attrs)
    label = obj.coords[dim].values if dim in obj.coords else None
    for k in obj.keys():
      out[k] = _with_dim(obj[k].assign_coords({dim: label})
                         if label is not None and dim not in obj[k].dims
                         else obj[k], dim)
    return out
  if dim in obj.dims:
    return obj
  label = np.asarray(obj.coords[dim].values).reshape(1) if (
      dim in obj.coords) else None
  coords = {k: c for k, c in obj.coords.items() if k != dim}
  if label is not None:
    coords[dim] = _xl.Coord((dim,), label)
  return DataArray(np.asarray(obj.values)[None], (dim,) + obj.dims, coords,
                   obj.name, obj.attrs)


def concat(objs, dim, **kwargs):  # pylint: disable=function-redefined
  objs = list(objs)
  if isinstance(dim, str) and any(
      dim in (o.dims if isinstance(o, DataArray) else o.sizes) for o in objs):
    objs = [_with_dim(o, dim) for o in objs]
  return _outer_concat(objs, dim, **kwargs)


class _GroupBy:
  def __init__(self, obj, dim):
    self.obj, self.dim = obj, dim

  def apply(self, func):
    n = self.obj.sizes[self.dim]
    parts = [func(self.obj.isel({self.dim: slice(i, i + 1)})) for i in range(n)]
    labels = self.obj.coords[self.dim].values
    parts = [p.assign_coords({self.dim: labels[i]}) if self.dim not in p.dims
             else p for i, p in enumerate(parts)]
    return concat(parts, self.dim)

  map = apply


DataArray.groupby = lambda self, group, squeeze=True, **kw: _GroupBy(self, group)


def _da_interp(self, coords=None, method='linear', **kw):
  """1-D linear interpolation along ONE dimension coordinate; NaN outside the
  coordinate's range (scipy.interpolate.interp1d, bounds_error=False)."""
  assert method == 'linear'
  (dim, new), = dict(coords or {}, **kw).items()
  new = np.asarray(new, dtype=np.float64)
  x = np.asarray(self.coords[dim].values, dtype=np.float64)
  ax = self.dims.index(dim)
  v = np.moveaxis(np.asarray(self.values, dtype=np.float64), ax, -1)
  from scipy import interpolate
  f = interpolate.interp1d(x, v, kind='linear', axis=-1, bounds_error=False,
                           fill_value=np.nan, assume_sorted=False)
  out = np.moveaxis(f(new), -1, ax)
  c = {k: cc for k, cc in self.coords.items() if dim not in cc.dims}
  c[dim] = _xl.Coord((dim,), new)
  return DataArray(out, self.dims, c, self.name, self.attrs)


DataArray.interp = _da_interp
DataArray.__floordiv__ = lambda self, o: self._binary(o, np.floor_divide)  # pylint: disable=protected-access
DataArray.__mod__ = lambda self, o: self._binary(o, np.mod)  # pylint: disable=protected-access


# ---- needed only by the reference's own TEST files (run on this stand-in to
# validate it: tests/golden/run_reference_tests.sh) -----------------------------------
from . import testing  # noqa: E402,F401  pylint: disable=wrong-import-position


def ones_like(obj):
  return zeros_like(obj) + 1


def full_like(obj, fill_value):
  return zeros_like(obj) + fill_value


_lite_broadcast = _xl._broadcast  # pylint: disable=protected-access


def _broadcast_with_arrays(a, b):
  """A bare ndarray operand follows NumPy's rules (trailing axes align)."""
  if isinstance(b, np.ndarray) and b.ndim > 0 and b.shape != a.shape:
    av = a.values
    np.broadcast_shapes(av.shape, b.shape)  # raises if incompatible
    assert b.ndim <= av.ndim
    return av, b, a.dims, dict(a.coords)
  return _lite_broadcast(a, b)


_xl._broadcast = _broadcast_with_arrays  # pylint: disable=protected-access


def _positions_for_missing_coords(obj, idx):
  """Dimensions without a coordinate are indexed by position in `.sel`."""
  dims = obj.dims if isinstance(obj, DataArray) else obj.sizes
  free = {d: v for d, v in idx.items() if d in dims and d not in obj.coords}
  return free, {d: v for d, v in idx.items() if d not in free}


_sel_with_labels = DataArray.sel


def _da_sel_any(self, indexers=None, method=None, drop=False, tolerance=None,
                **kw):
  idx = dict(indexers or {}, **kw)
  free, labelled = _positions_for_missing_coords(self, idx)
  out = self
  if free:
    vec = {d: v for d, v in free.items()
           if isinstance(v, DataArray) and v.ndim >= 1}
    if vec:  # N-d positional indexers: give the dimension a 0..n-1 coordinate
      out = out.assign_coords({d: np.arange(out.sizes[d]) for d in vec})
      labelled.update(vec)
    plain = {d: v for d, v in free.items() if d not in vec}
    if plain:
      out = out.isel(plain, drop=drop)
  if labelled:
    out = _sel_with_labels(out, labelled, method=method, drop=drop,
                           tolerance=tolerance)
  return out


DataArray.sel = _da_sel_any
_lite_ds_getitem = Dataset.__getitem__


def _ds_getitem(self, key):
  if isinstance(key, str) and key not in self.keys() and (
      key not in self.coords) and key in self.sizes:
    return DataArray(np.arange(self.sizes[key]), (key,), name=key)
  return _lite_ds_getitem(self, key)


Dataset.__getitem__ = _ds_getitem


class _Loc:
  """obj.loc[{dim: label}] for reading and for in-place assignment."""

  def __init__(self, obj):
    self.obj = obj

  def __getitem__(self, key):
    return self.obj.sel(key)

  def __setitem__(self, key, value):
    targets = ([self.obj[k] for k in self.obj.keys()]
               if isinstance(self.obj, Dataset) else [self.obj])
    for da in targets:
      index = [slice(None)] * len(da.dims)
      for d, label in key.items():
        if d in da.dims:
          index[da.dims.index(d)] = int(_xl._lookup(  # pylint: disable=protected-access
              da.coords[d].values, np.asarray([label]))[0])
      v = value[da.name] if isinstance(value, Dataset) else value
      v = np.asarray(v.values if isinstance(v, DataArray) else v)
      da.data[tuple(index)] = v


DataArray.loc = property(_Loc)
Dataset.loc = property(_Loc)


DataArray.__format__ = lambda self, spec: format(
    np.asarray(self.values)[()] if self.ndim == 0 else self.values, spec)

_moment_with_dims = _moment


def _moment(self, dim, skipna, ddof, root):  # pylint: disable=function-redefined
  return _moment_with_dims(self, self.dims if dim is None else dim, skipna,
                           ddof, root)


DataArray.var = lambda self, dim=None, skipna=None, ddof=0, **kw: _moment(
    self, dim, skipna, ddof, False)
DataArray.std = lambda self, dim=None, skipna=None, ddof=0, **kw: _moment(
    self, dim, skipna, ddof, True)

_concat_same_dims = concat


def _broadcast_to_common_dims(objs, dim):
  """xr.concat gives every object the union of the dimensions (a truth without
  prediction_timedelta is repeated along it)."""
  if not all(isinstance(o, DataArray) for o in objs):
    keys = list(objs[0].keys())
    per_var = {k: _broadcast_to_common_dims([o[k] for o in objs], dim)
               for k in keys}
    out = []
    for i, o in enumerate(objs):
      d = Dataset(attrs=o.attrs)
      for k in keys:
        d[k] = per_var[k][i]
      out.append(d)
    return out
  sizes = {}
  for o in objs:
    for d, n in o.sizes.items():
      if d != dim:
        sizes.setdefault(d, n)
  order = [d for d in objs[-1].dims if d != dim] + [
      d for d in sizes if d not in objs[-1].dims]
  done = []
  for o in objs:
    lead = [d for d in o.dims if d == dim]
    missing = [d for d in order if d not in o.dims]
    if missing:
      v = np.asarray(o.values)
      shape = tuple(sizes[d] for d in missing) + v.shape
      coords = dict(o.coords)
      for other in objs:
        for d in missing:
          if d in other.coords and d not in coords:
            coords[d] = other.coords[d]
      o = DataArray(np.broadcast_to(v, shape).copy(), tuple(missing) + o.dims,
                    coords, o.name, o.attrs)
    done.append(o.transpose(*lead, *order))
  return done


def concat(objs, dim, **kwargs):  # pylint: disable=function-redefined
  objs = list(objs)
  if isinstance(dim, str) and any(
      dim in (o.dims if isinstance(o, DataArray) else o.sizes) for o in objs):
    objs = [_with_dim(o, dim) for o in objs]
    same = lambda o: (tuple(o.dims) if isinstance(o, DataArray)  # noqa: E731
                      else tuple((k, o[k].dims) for k in o.keys()))
    if any(same(o) != same(objs[0]) for o in objs):
      objs = _broadcast_to_common_dims(objs, dim)
  return _outer_concat(objs, dim, **kwargs)


def _da_integrate(self, coord):
  if coord not in self.dims:  # a non-index coordinate: integrate along its dim
    (dim,) = self.coords[coord].dims
    return _da_integrate(self.swap_dims({dim: coord}), coord)
  ax = self.dims.index(coord)
  x = np.asarray(self.coords[coord].values, dtype=np.float64)
  keep = tuple(d for d in self.dims if d != coord)
  return self._replace(np.trapezoid(np.asarray(self.values), x, axis=ax), keep)  # pylint: disable=protected-access


DataArray.integrate = _da_integrate
Dataset.integrate = lambda self, coord: self._map(  # pylint: disable=protected-access
    lambda v: v.integrate(coord) if coord in v.dims else v)


def _da_argmax(self, dim):
  ax = self.dims.index(dim)
  return self._replace(np.argmax(self.values, axis=ax),  # pylint: disable=protected-access
                       tuple(d for d in self.dims if d != dim))


DataArray.argmax = _da_argmax


def _roll(self, shifts=None, roll_coords=False, **kw):
  shifts = dict(shifts or {}, **kw)
  if isinstance(self, Dataset):
    out = Dataset(attrs=self.attrs)
    for k in self.keys():
      out[k] = _roll(self[k], shifts, roll_coords)
    return out
  data = np.asarray(self.values)
  coords = dict(self.coords)
  for d, n in shifts.items():
    if d not in self.dims:
      continue
    data = np.roll(data, n, axis=self.dims.index(d))
    if roll_coords:
      for k, c in list(coords.items()):
        if d in c.dims:
          coords[k] = _xl.Coord(c.dims, np.roll(c.values, n,
                                                axis=c.dims.index(d)))
  return DataArray(data, self.dims, coords, self.name, self.attrs)


DataArray.roll = _roll
Dataset.roll = _roll


class _LabelGroupBy:
  """obj.groupby(<1-D DataArray of labels along one dim>).min() / .max()."""

  def __init__(self, obj, group):
    self.obj, self.group = obj, group
    (self.dim,) = group.dims
    self.name = group.name or self.dim

  def _reduce(self, fn):
    labels = np.asarray(self.group.values)
    uniq = np.unique(labels)

    def one(da):
      if self.dim not in da.dims:
        return da
      ax = da.dims.index(self.dim)
      v = np.asarray(da.values)
      parts = [fn(np.take(v, np.nonzero(labels == u)[0], axis=ax), axis=ax)
               for u in uniq]
      coords = {k: c for k, c in da.coords.items() if self.dim not in c.dims}
      dims = tuple(self.name if d == self.dim else d for d in da.dims)
      coords[self.name] = _xl.Coord((self.name,), uniq)
      return DataArray(np.stack(parts, axis=ax), dims, coords, da.name,
                       da.attrs)

    obj = self.obj
    return obj._map(one) if isinstance(obj, Dataset) else one(obj)  # pylint: disable=protected-access

  def min(self, *a, **k):
    return self._reduce(np.min)

  def max(self, *a, **k):
    return self._reduce(np.max)


def _groupby(self, group, squeeze=True, restore_coord_dims=None, **kw):
  del squeeze, restore_coord_dims, kw
  if isinstance(group, DataArray):
    return _LabelGroupBy(self, group)
  return _GroupBy(self, group)


DataArray.groupby = _groupby
Dataset.groupby = _groupby


def _to_zarr(self, path, **kwargs):
  del kwargs
  register_store(str(path), self)


Dataset.to_zarr = _to_zarr


def open_dataset(path, **kwargs):
  """Reads what Dataset.to_netcdf of this stand-in wrote."""
  del kwargs
  import pickle
  with open(path, 'rb') as f:
    payload = pickle.load(f)
  return Dataset({k: (d, v) for k, (d, v) in payload['vars'].items()},
                 {k: (d, v) for k, (d, v) in payload['coords'].items()})
