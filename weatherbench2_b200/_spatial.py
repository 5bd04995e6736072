"""Host-side planning for the spatial-reduction kernels.

Turns named-dimension arrays (NumPy on the host or torch tensors on a CUDA
device) plus a list of regions into what the C ABI wants: a base pointer and a
per-field element-offset table for every operand, and a `wb2_weights`
description of latitude weights x region masks.  Nothing numerical happens
here except building the (tiny) weight vectors with the reference's own
formulas (weatherbench2/metrics.py:35-60).
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Sequence

import numpy as np

from weatherbench2_b200 import _lib
from weatherbench2_b200 import xarray_lite as xl

LAT, LON = 'latitude', 'longitude'


# ---- latitude weights (metrics.py:35-60) -------------------------------------
def _assert_increasing(x: np.ndarray):
  if not (np.diff(x) > 0).all():
    raise ValueError(f'array is not increasing: {x}')


def _latitude_cell_bounds(x: np.ndarray) -> np.ndarray:
  pi_over_2 = np.array([np.pi / 2], dtype=x.dtype)
  return np.concatenate([-pi_over_2, (x[:-1] + x[1:]) / 2, pi_over_2])


def _cell_area_from_latitude(points: np.ndarray) -> np.ndarray:
  bounds = _latitude_cell_bounds(points)
  _assert_increasing(bounds)
  return np.sin(bounds[1:]) - np.sin(bounds[:-1])


def lat_weights(latitude_deg: np.ndarray) -> np.ndarray:
  """Area weights normalised to mean 1 (weatherbench2/metrics.py:55-60)."""
  w = _cell_area_from_latitude(np.deg2rad(np.asarray(latitude_deg)))
  return w / np.mean(w)


# ---- operand description -----------------------------------------------------
@dataclasses.dataclass
class Operand:
  """One named array prepared for the kernels."""
  data: object               # np.ndarray or torch.Tensor (kept alive)
  addr: int                  # address of element 0
  itemsize: int
  on_device: bool
  outer_dims: tuple          # non-spatial dims
  outer_shape: tuple
  outer_strides: tuple       # in elements
  layout: str                # 'lat_lon' (lon contiguous) | 'lon_lat'
  nrow: int
  ncol: int
  row_stride: int
  dtype: np.dtype


def _strides_elems(data) -> tuple:
  if xl._is_torch(data):  # pylint: disable=protected-access
    return tuple(int(s) for s in data.stride())
  return tuple(int(s // data.itemsize) for s in data.strides)


def _address(data) -> int:
  if xl._is_torch(data):  # pylint: disable=protected-access
    return int(data.data_ptr())
  return int(data.ctypes.data)


def _np_dtype(data) -> np.dtype:
  if xl._is_torch(data):  # pylint: disable=protected-access
    return np.dtype(str(data.dtype).replace('torch.', ''))
  return data.dtype


def prepare_operand(da: xl.DataArray, want_layout: Optional[str] = None,
                    want_dtype: Optional[np.dtype] = None) -> Operand:
  """Describes `da` for the kernels, copying only when its memory layout is
  not (…outer…, row, col) with `col` contiguous, or its dtype is neither
  float32 nor float64."""
  if LAT not in da.dims or LON not in da.dims:
    raise ValueError(f'{da.name!r} needs latitude and longitude dims, has '
                     f'{da.dims}')
  lazy = getattr(da, 'lazy_source', None)
  if lazy is not None:
    # a label gather (e.g. truth.sel(time=valid_time)): address the source
    # through the offset table instead of materialising the copy
    source, maps = lazy
    return gather_operand(prepare_operand(source, want_layout, want_dtype),
                          maps)
  data = da.data
  dims = da.dims
  torch_in = xl._is_torch(data)  # pylint: disable=protected-access
  dt = _np_dtype(data)
  target_dt = np.dtype(want_dtype) if want_dtype is not None else (
      dt if dt in (np.dtype('float32'), np.dtype('float64'))
      else np.dtype('float32') if dt.itemsize <= 4 and dt.kind == 'f'
      else np.dtype('float64'))
  if dt != target_dt:
    if torch_in:
      import torch  # pylint: disable=import-outside-toplevel
      data = data.to(getattr(torch, target_dt.name))
    else:
      data = data.astype(target_dt)
  strides = _strides_elems(data)
  ilat, ilon = dims.index(LAT), dims.index(LON)
  shape = tuple(data.shape)

  def slab_is_whole(irow, icol):
    # no other dimension may be interleaved with the (row, col) slab: every
    # outer stride spans a whole slab (or is 0: a broadcast view).  Otherwise
    # the slab of one field is scattered through memory -- two operands with
    # the same dims in a different order would disagree on the row stride, and
    # the kernels would stride through HBM -- so such arrays are compacted.
    span = (shape[irow] - 1) * strides[irow] + shape[icol]
    return all(shape[i] == 1 or strides[i] == 0 or strides[i] >= span
               for i in range(len(shape)) if i not in (irow, icol))

  def layout_ok():
    # rows must also be densely packed (row stride == ncol): operands of one
    # launch share ONE row stride in the C ABI, and a strided view on one side
    # (a longitude box, every second row, an integer pick of an inner
    # dimension) against a dense array on the other would otherwise be
    # refused.  Such views are compacted; whole-slab views stay zero-copy.
    if strides[ilon] == 1 or shape[ilon] == 1:
      lay = 'lat_lon'
      if (strides[ilat] == shape[ilon] or shape[ilat] == 1) and slab_is_whole(
          ilat, ilon):
        return lay
    if strides[ilat] == 1 or shape[ilat] == 1:
      lay = 'lon_lat'
      if (strides[ilon] == shape[ilat] or shape[ilon] == 1) and slab_is_whole(
          ilon, ilat):
        return lay
    return None

  lay = layout_ok()
  if any(s < 0 for s in strides):
    lay = None
  if lay is None or (want_layout is not None and lay != want_layout and not (
      shape[ilat] == 1 or shape[ilon] == 1)):
    # contiguous copy with the spatial dims last, in the wanted order
    order = want_layout or 'lat_lon'
    sp = (LAT, LON) if order == 'lat_lon' else (LON, LAT)
    perm_dims = tuple(d for d in dims if d not in sp) + sp
    perm = [dims.index(d) for d in perm_dims]
    if torch_in:
      data = data.permute(*perm).contiguous()
    else:
      data = np.ascontiguousarray(np.transpose(data, perm))
    dims = perm_dims
    strides = _strides_elems(data)
    shape = tuple(data.shape)
    ilat, ilon = dims.index(LAT), dims.index(LON)
    lay = order
  if want_layout is not None and lay != want_layout:
    lay = want_layout  # degenerate size-1 axis: either description is valid
  if lay == 'lat_lon':
    nrow, ncol, row_stride = shape[ilat], shape[ilon], strides[ilat]
  else:
    nrow, ncol, row_stride = shape[ilon], shape[ilat], strides[ilon]
  if nrow == 1:
    row_stride = max(row_stride, ncol)
  outer = [i for i in range(len(dims)) if i not in (ilat, ilon)]
  on_dev = bool(torch_in and data.is_cuda)
  if torch_in and not on_dev:
    data = data.numpy()
  return Operand(
      data=data, addr=_address(data), itemsize=target_dt.itemsize,
      on_device=on_dev, outer_dims=tuple(dims[i] for i in outer),
      outer_shape=tuple(shape[i] for i in outer),
      outer_strides=tuple(strides[i] for i in outer), layout=lay, nrow=nrow,
      ncol=ncol, row_stride=int(row_stride), dtype=target_dt)


def broadcast_dims(*operands: Operand) -> tuple[tuple, tuple]:
  """Result dims/shape of xarray arithmetic between the operands: dims of the
  first, then unseen dims of the next ones."""
  dims, shape = [], []
  for op in operands:
    for d, n in zip(op.outer_dims, op.outer_shape):
      if d in dims:
        if shape[dims.index(d)] != n:
          raise ValueError(
              f'size mismatch along {d!r}: {shape[dims.index(d)]} vs {n}')
      else:
        dims.append(d)
        shape.append(n)
  return tuple(dims), tuple(shape)


def offset_table(op: Operand, dims: tuple, shape: tuple) -> np.ndarray:
  """Element offset of the slab of `op` for every index of the broadcast outer
  grid (C order).  Missing dims get stride 0 (broadcast)."""
  off = np.zeros(shape, dtype=np.int64)
  gathered = getattr(op, 'gather_terms', None)
  if gathered is not None:
    # explicit (dims, offsets) terms: a label lookup folded into the table
    for tdims, arr in gathered:
      perm = [tdims.index(d) for d in dims if d in tdims]
      a = np.transpose(np.asarray(arr, dtype=np.int64), perm)
      a = a[tuple(slice(None) if d in tdims else None for d in dims)]
      off = off + a
    return np.ascontiguousarray(np.broadcast_to(off, shape)).reshape(-1)
  for ax, (d, n) in enumerate(zip(dims, shape)):
    if d in op.outer_dims:
      st = op.outer_strides[op.outer_dims.index(d)]
      sh = [1] * len(shape)
      sh[ax] = n
      off = off + (np.arange(n, dtype=np.int64) * st).reshape(sh)
  return np.ascontiguousarray(off).reshape(-1)


def gather_operand(op: Operand, index_maps: dict) -> Operand:
  """Re-addresses `op` through label lookups without copying data.

  index_maps: {source_dim: (new_dims, positions)} -- `source_dim` of `op` is
  indexed by the integer array `positions` whose dims are `new_dims`
  (xarray's vectorised `.sel`, e.g. the day-of-year / hour lookup of
  weatherbench2/metrics.py:398-404 or `truth.sel(time=valid_time)`,
  evaluation.py:475).  Untouched outer dims keep their stride addressing.
  """
  terms = []
  new_dims, new_shape = [], []

  def add_dim(d, n):
    if d in new_dims:
      if new_shape[new_dims.index(d)] != n:
        raise ValueError(f'size mismatch along {d!r}')
    else:
      new_dims.append(d)
      new_shape.append(n)

  for d, n, st in zip(op.outer_dims, op.outer_shape, op.outer_strides):
    if d in index_maps:
      tdims, pos = index_maps[d]
      pos = np.asarray(pos, dtype=np.int64)
      if pos.size and (pos.min() < 0 or pos.max() >= n):
        raise IndexError(f'gather positions out of range for {d!r}')
      terms.append((tuple(tdims), pos * st))
      for td, tn in zip(tdims, pos.shape):
        add_dim(td, tn)
    else:
      terms.append(((d,), np.arange(n, dtype=np.int64) * st))
      add_dim(d, n)
  out = dataclasses.replace(op, outer_dims=tuple(new_dims),
                            outer_shape=tuple(new_shape),
                            outer_strides=tuple(0 for _ in new_dims))
  out.gather_terms = terms
  return out


# ---- weights -----------------------------------------------------------------
def region_factors(regions: Sequence, latitude: np.ndarray,
                   longitude: np.ndarray):
  out = []
  for r in regions:
    if r is None:
      from weatherbench2_b200.regions import RegionFactors  # pylint: disable=import-outside-toplevel
      out.append(RegionFactors(np.ones(latitude.size), np.ones(longitude.size)))
    else:
      out.append(r.factors(latitude, longitude))
  return out


def _segments(factors_along_col: np.ndarray):
  """Column segments on which every region's factor is constant.
  factors_along_col: [R, ncol].  Returns (seg_start[nseg+1], seg_w[R, nseg])."""
  ncol = factors_along_col.shape[1]
  change = np.any(factors_along_col[:, 1:] != factors_along_col[:, :-1], axis=0)
  starts = np.concatenate([[0], np.nonzero(change)[0] + 1]).astype(np.int32)
  seg_start = np.concatenate([starts, [ncol]]).astype(np.int32)
  seg_w = factors_along_col[:, starts]
  return seg_start, np.ascontiguousarray(seg_w, dtype=np.float64)


def build_weights(ctx: _lib.Context, latitude: np.ndarray,
                  longitude: np.ndarray, regions: Sequence, layout: str,
                  row_stride: int, cell_cache: Optional[dict] = None,
                  max_regions: int = _lib.MAX_REGIONS
                  ) -> list[tuple[list[int], _lib.WeightSpec]]:
  """WeightSpecs for `regions` (entries may be None = global).

  Regions that carry a 2-D mask (LandRegion) cannot share a launch with a
  different mask, so the result is a list of (region indices, WeightSpec)
  groups; regions without a mask all go in the first group.
  """
  wlat = lat_weights(latitude)
  facs = region_factors(regions, latitude, longitude)
  if cell_cache is not None:
    # the cached values are DEVICE pointers: they belong to this context (its
    # device, its lifetime), whatever dict the caller passed to ask for caching
    cell_cache = ctx.__dict__.setdefault('_cell_weight_cache', {})
  groups: dict = {}
  for i, fc in enumerate(facs):
    key = None if fc.cell is None else fc.cell.tobytes()
    groups.setdefault(key, []).append(i)
  out = []
  nlat, nlon = latitude.size, longitude.size
  for key, idx in groups.items():
    for j0 in range(0, len(idx), max_regions):
      ids = idx[j0:j0 + max_regions]
      latf = np.stack([facs[i].lat for i in ids])  # [R, nlat]
      lonf = np.stack([facs[i].lon for i in ids])  # [R, nlon]
      zero_skip = any(regions[i] is not None for i in ids)
      cell_dev = None
      if key is not None:
        cell = facs[ids[0]].cell
        cell = cell if layout == 'lat_lon' else cell.T
        cache_key = (key, cell.shape, layout)
        if cell_cache is not None and cache_key in cell_cache:
          cell_dev = cell_cache[cache_key]
        else:
          cell_dev = ctx.to_device(np.ascontiguousarray(cell, np.float32))
          if cell_cache is not None:
            cell_cache[cache_key] = cell_dev
      if layout == 'lat_lon':
        seg_start, seg_w = _segments(lonf)
        spec = _lib.WeightSpec(nlat, nlon, latf * wlat[None, :], seg_start,
                               seg_w, None, cell_dev, zero_skip, row_stride)
      else:
        seg_start, seg_w = _segments(latf)
        spec = _lib.WeightSpec(nlon, nlat, lonf, seg_start, seg_w,
                               wlat.astype(np.float32), cell_dev, zero_skip,
                               row_stride)
      out.append((ids, spec))
  return out


# ---- launches ----------------------------------------------------------------
def _common_base(ops: Sequence[Operand]) -> int:
  return min(op.addr for op in ops)


def run_det_metrics(ctx: _lib.Context, f_ops: Sequence[Operand],
                    t_ops: Sequence[Operand],
                    c_ops: Optional[Sequence[Operand]], latitude, longitude,
                    regions: Sequence, skipna: bool, cell_cache=None):
  """Runs K1 for a list of variables that share dtype / layout / grid.

  Returns (stats, dims, shapes): stats[v] has shape outer_shape[v] +
  (len(regions), DET_NSTAT).
  """
  nvar = len(f_ops)
  first = f_ops[0]
  on_dev = first.on_device
  dtype_code = _lib.F32 if first.dtype == np.float32 else _lib.F64
  es = first.itemsize
  all_ops = list(f_ops) + list(t_ops) + (list(c_ops) if c_ops else [])
  for op in all_ops:
    if (op.on_device != on_dev or op.dtype != first.dtype or
        op.layout != first.layout or op.nrow != first.nrow or
        op.ncol != first.ncol or op.row_stride != first.row_stride):
      raise ValueError('operands of one launch must share device, dtype, '
                       'layout, grid and row stride')
  base = _common_base(all_ops)
  dims_list, shape_list, tabs = [], [], [[], [], []]
  for v in range(nvar):
    trio = [f_ops[v], t_ops[v]] + ([c_ops[v]] if c_ops else [])
    dims, shape = broadcast_dims(*trio)
    dims_list.append(dims)
    shape_list.append(shape)
    for k, op in enumerate(trio):
      rel = (op.addr - base)
      assert rel % es == 0
      tabs[k].append(offset_table(op, dims, shape) + rel // es)
  off_f = np.ascontiguousarray(np.concatenate(tabs[0]))
  off_t = np.ascontiguousarray(np.concatenate(tabs[1]))
  off_c = np.ascontiguousarray(np.concatenate(tabs[2])) if c_ops else None
  nfield = off_f.size
  nreg = len(regions)
  result = np.empty((nfield, nreg, _lib.DET_NSTAT), dtype=np.float64)
  groups = build_weights(ctx, np.asarray(latitude), np.asarray(longitude),
                         regions, first.layout, first.row_stride, cell_cache)
  for ids, spec in groups:
    R = len(ids)
    if on_dev:
      out_dev = ctx.malloc(nfield * R * _lib.DET_NSTAT * 8)
      try:
        ctx.det_metrics(base, base, base if c_ops else None, dtype_code, off_f,
                        off_t, off_c, spec, skipna, out_dev)
        part = ctx.from_device(out_dev, (nfield, R, _lib.DET_NSTAT),
                               np.float64)
      finally:
        ctx.free(out_dev)
    else:
      part = np.empty((nfield, R, _lib.DET_NSTAT), dtype=np.float64)
      ctx.det_metrics(base, base, base if c_ops else None, dtype_code, off_f,
                      off_t, off_c, spec, skipna, part.ctypes.data, host=True)
    result[:, ids, :] = part
  stats, pos = [], 0
  for v in range(nvar):
    n = int(np.prod(shape_list[v])) if shape_list[v] else 1
    stats.append(result[pos:pos + n].reshape(
        shape_list[v] + (nreg, _lib.DET_NSTAT)))
    pos += n
  return stats, dims_list, shape_list


# ---- K2: ensemble metrics ----------------------------------------------------
def split_member_dim(op: Operand, ens_dim: str):
  """Removes `ens_dim` from the outer dims of `op`; returns (op', M, stride)."""
  if ens_dim not in op.outer_dims:
    raise ValueError(f'ensemble_dim={ens_dim!r} not found in {op.outer_dims}')
  i = op.outer_dims.index(ens_dim)
  m, stride = op.outer_shape[i], op.outer_strides[i]
  keep = [k for k in range(len(op.outer_dims)) if k != i]
  out = dataclasses.replace(
      op, outer_dims=tuple(op.outer_dims[k] for k in keep),
      outer_shape=tuple(op.outer_shape[k] for k in keep),
      outer_strides=tuple(op.outer_strides[k] for k in keep))
  terms = getattr(op, 'gather_terms', None)
  if terms is not None:
    # a gathered operand addresses through explicit (dims, offsets) terms: the
    # member axis must be a plain strided term
    rest = []
    for tdims, arr in terms:
      if ens_dim in tdims:
        if tdims != (ens_dim,):
          raise ValueError('the ensemble dimension cannot be part of a gather')
        arr = np.asarray(arr)
        stride = int(arr[1] - arr[0]) if arr.size > 1 else 0
        if arr.size > 2 and not (np.diff(arr) == stride).all():
          raise ValueError('ensemble members must be evenly strided')
      else:
        rest.append((tdims, arr))
    out.gather_terms = rest
  return out, int(m), int(stride)


def _upload_gathered(ctx: _lib.Context, op: Operand, staged: list) -> Operand:
  """Uploads ONLY the slabs a gathered host operand references (a by-init truth
  gather or a day-of-year climatology lookup addresses a few slabs of a long
  record: uploading the whole backing array, per variable and chunk, is what
  the round-1 review flagged), packed contiguously, and re-addresses the
  operand.  Dimensions that were plain strided terms (e.g. the ensemble
  members) stay plain strided terms of the packed copy, so that
  split_member_dim still sees an evenly strided member axis; the looked-up
  dimensions become one explicit offset term over the DISTINCT slabs."""
  terms = op.gather_terms
  plain, looked = [], []
  for tdims, arr in terms:
    arr = np.asarray(arr, dtype=np.int64)
    if len(tdims) == 1 and (arr.size < 2 or
                            (np.diff(arr) == arr[1] - arr[0]).all()):
      plain.append((tdims[0], arr))
    else:
      looked.append((tuple(tdims), arr))
  size = dict(zip(op.outer_dims, op.outer_shape))
  gdims = tuple(d for d in op.outer_dims if any(d in td for td, _ in looked))
  gshape = tuple(size[d] for d in gdims)
  goff = np.zeros(gshape, dtype=np.int64)
  for tdims, arr in looked:
    perm = [tdims.index(d) for d in gdims if d in tdims]
    a = np.transpose(arr, perm)
    a = a[tuple(slice(None) if d in tdims else None for d in gdims)]
    goff = goff + a
  uniq, inv = np.unique(goff.reshape(-1), return_inverse=True)
  slab = (op.nrow - 1) * op.row_stride + op.ncol
  pad = (slab + 63) // 64 * 64
  es = op.itemsize
  pshape = tuple(arr.size for _, arr in plain)
  nplain = int(np.prod(pshape)) if pshape else 1
  packed = np.empty((nplain, uniq.size, pad), dtype=op.dtype)
  grids = np.meshgrid(*[arr for _, arr in plain], indexing='ij') if plain else []
  poff = sum(grids).reshape(-1) if plain else np.zeros(1, np.int64)
  for pi in range(nplain):
    for ui, off in enumerate(uniq):
      src = np.ctypeslib.as_array(
          (np.ctypeslib.ctypes.c_char * (slab * es)).from_address(
              op.addr + (int(poff[pi]) + int(off)) * es)).view(op.dtype)
      packed[pi, ui, :slab] = src
  dptr = ctx.to_device(packed)
  staged.append(dptr)
  out = dataclasses.replace(op, addr=dptr, on_device=True)
  new_terms = []
  stride = uniq.size * pad
  for i in range(len(plain) - 1, -1, -1):
    d, arr = plain[i]
    new_terms.insert(0, ((d,), np.arange(arr.size, dtype=np.int64) * stride))
    stride *= arr.size
  if gdims:
    new_terms.append((gdims, (inv.astype(np.int64) * pad).reshape(gshape)))
  out.gather_terms = new_terms
  return out


def _to_device_operand(ctx: _lib.Context, op: Operand, staged: list) -> Operand:
  """Uploads a host operand to the device: the smallest contiguous span of its
  backing array (same strides), or, for a gathered operand, only the slabs it
  references."""
  if op.on_device:
    return op
  if getattr(op, 'gather_terms', None) is not None:
    return _upload_gathered(ctx, op, staged)
  arr = op.data
  base = arr
  while isinstance(base, np.ndarray) and base.base is not None and isinstance(
      base.base, np.ndarray):
    base = base.base
  # upload the smallest contiguous span that covers the view
  lo = arr.ctypes.data
  span = 1 + sum((n - 1) * abs(s) for n, s in zip(arr.shape, arr.strides)
                 ) // arr.itemsize if arr.size else 1
  flat = np.ctypeslib.as_array(
      (np.ctypeslib.ctypes.c_char * (span * arr.itemsize)).from_address(lo))
  dptr = ctx.to_device(flat)
  staged.append(dptr)
  out = dataclasses.replace(op, addr=dptr, on_device=True)
  if hasattr(op, 'gather_terms'):
    out.gather_terms = op.gather_terms
  return out


def run_ens_metrics(ctx: _lib.Context, x_ops: Sequence[Operand],
                    t_ops: Sequence[Operand], ens_dim: str, latitude,
                    longitude, regions: Sequence, skipna: bool,
                    cell_cache=None):
  """Runs K2 for variables sharing layout / grid.  Returns (stats, dims,
  shapes, M): stats[v] has shape outer_shape[v] + (len(regions), ENS_NSTAT)."""
  staged: list = []
  # all-host float32 operands: stream member / truth slabs through the staging
  # buffers (wb2_ens_metrics_host; truth slabs go through the slab cache)
  host = all(not o.on_device and o.dtype == np.float32
             for o in list(x_ops) + list(t_ops))
  try:
    xs, ts, ms, strides = [], [], [], []
    for xo, to in zip(x_ops, t_ops):
      if ens_dim in to.outer_dims:
        raise ValueError(f'truth must not have the {ens_dim!r} dimension')
      if not host:
        xo = _to_device_operand(ctx, xo, staged)
        to = _to_device_operand(ctx, to, staged)
      xo, m, st = split_member_dim(xo, ens_dim)
      xs.append(xo)
      ts.append(to)
      ms.append(m)
      strides.append(st)
    first = xs[0]
    nreg = len(regions)
    groups = build_weights(ctx, np.asarray(latitude), np.asarray(longitude),
                           regions, first.layout, first.row_stride, cell_cache)
    stats, dims_list, shape_list = [], [], []
    for xo, to, m, st in zip(xs, ts, ms, strides):
      if (to.layout != xo.layout or to.row_stride != xo.row_stride or
          to.nrow != xo.nrow or to.ncol != xo.ncol):
        raise ValueError('forecast and truth must share layout and grid')
      # xarray puts truth's dims first for abs(truth - forecast)
      # (metrics.py:824); we keep the forecast-first order used by every other
      # metric of this module and of the reference's `mean/var` based ones.
      dims, shape = broadcast_dims(xo, to)
      base = min(xo.addr, to.addr)
      off_x = offset_table(xo, dims, shape) + (xo.addr - base) // 4
      off_t = offset_table(to, dims, shape) + (to.addr - base) // 4
      nfield = off_x.size
      res = np.empty((nfield, nreg, _lib.ENS_NSTAT), dtype=np.float64)
      for ids, spec in groups:
        if host:
          part = np.empty((nfield, len(ids), _lib.ENS_NSTAT), np.float64)
          ctx.ens_metrics_host(base, base, m, st, off_x, off_t, spec, skipna,
                               part.ctypes.data)
          res[:, ids, :] = part
          continue
        out_dev = ctx.malloc(nfield * len(ids) * _lib.ENS_NSTAT * 8)
        try:
          ctx.ens_metrics(base, base, _lib.F32, m, st, off_x, off_t, spec,
                          skipna, out_dev)
          res[:, ids, :] = ctx.from_device(
              out_dev, (nfield, len(ids), _lib.ENS_NSTAT), np.float64)
        finally:
          ctx.free(out_dev)
      stats.append(res.reshape(shape + (nreg, _lib.ENS_NSTAT)))
      dims_list.append(dims)
      shape_list.append(shape)
    return stats, dims_list, shape_list, ms
  finally:
    for p in staged:
      ctx.free(p)


# ---- K3: energy score --------------------------------------------------------
ENERGY_MAX_REGIONS = _lib.MAX_REGIONS


def run_energy_score(ctx: _lib.Context, x_ops: Sequence[Operand],
                     t_ops: Sequence[Operand], ens_dim: str, latitude,
                     longitude, regions: Sequence, cell_cache=None):
  """Runs K3.  Returns (stats, dims, M): stats[v] has shape
  outer_shape[v] + (len(regions), 4, M) -- see wb2_energy_score."""
  staged: list = []
  try:
    first = None
    stats, dims_list, ms = [], [], []
    for xo, to in zip(x_ops, t_ops):
      xo = _to_device_operand(ctx, xo, staged)
      to = _to_device_operand(ctx, to, staged)
      xo, m, st = split_member_dim(xo, ens_dim)
      if first is None:
        first = xo
        groups = build_weights(ctx, np.asarray(latitude),
                               np.asarray(longitude), regions, xo.layout,
                               xo.row_stride, cell_cache, ENERGY_MAX_REGIONS)
      dims, shape = broadcast_dims(xo, to)
      base = min(xo.addr, to.addr)
      off_x = offset_table(xo, dims, shape) + (xo.addr - base) // 4
      off_t = offset_table(to, dims, shape) + (to.addr - base) // 4
      nfield = off_x.size
      res = np.empty((nfield, len(regions), 4, m), dtype=np.float64)
      for ids, spec in groups:
        out_dev = ctx.malloc(nfield * len(ids) * 4 * m * 8)
        try:
          ctx.energy_score(base, base, _lib.F32, m, st, off_x, off_t, spec,
                           out_dev)
          res[:, ids] = ctx.from_device(out_dev, (nfield, len(ids), 4, m),
                                        np.float64)
        finally:
          ctx.free(out_dev)
      stats.append(res.reshape(shape + (len(regions), 4, m)))
      dims_list.append(dims)
      ms.append(m)
    return stats, dims_list, ms
  finally:
    for p in staged:
      ctx.free(p)


# ---- K6: map-output metrics (time mean fused in) -------------------------------
def _grouped_tables(ops: Sequence[Operand], reduce_dim: Optional[str]):
  """Offset tables of `ops` over their broadcast outer grid with `reduce_dim`
  (if present) moved last: returns (tables, out_dims, out_shape, ngroup)."""
  dims, shape = broadcast_dims(*ops)
  if reduce_dim is not None and reduce_dim in dims:
    i = dims.index(reduce_dim)
    ngroup = shape[i]
    out_dims = dims[:i] + dims[i + 1:]
    out_shape = shape[:i] + shape[i + 1:]
    dims, shape = out_dims + (reduce_dim,), out_shape + (ngroup,)
  else:
    ngroup, out_dims, out_shape = 1, dims, shape
  return ([offset_table(op, dims, shape) for op in ops], out_dims, out_shape,
          ngroup)


def _map_dims(op: Operand):
  return (LAT, LON) if op.layout == 'lat_lon' else (LON, LAT)


def _alloc_maps(ctx: _lib.Context, like: Operand, shape: tuple, dtype):
  """Device buffer for the output maps: a torch tensor when the inputs are
  torch CUDA tensors (the result stays on the device), else a raw allocation
  that `_fetch_maps` copies back."""
  if xl._is_torch(like.data):  # pylint: disable=protected-access
    import torch  # pylint: disable=import-outside-toplevel
    out = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name),
                      device=like.data.device)
    return out, int(out.data_ptr())
  nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
  return None, ctx.malloc(max(nbytes, 16))


def _fetch_maps(ctx: _lib.Context, tensor, ptr: int, shape: tuple, dtype):
  if tensor is not None:
    return tensor
  try:
    return ctx.from_device(ptr, shape, dtype)
  finally:
    ctx.free(ptr)


def run_det_maps(ctx: _lib.Context, f_op: Operand, t_op: Operand, stat: int,
                 reduce_dim: Optional[str], skipna: bool):
  """Runs K6 for one variable: maps of stat(f, t) averaged over `reduce_dim`
  (None: no averaging).  Returns (maps, dims)."""
  staged: list = []
  try:
    was_dev = f_op.on_device
    f_op = _to_device_operand(ctx, f_op, staged)
    t_op = _to_device_operand(ctx, t_op, staged)
    if (t_op.layout != f_op.layout or t_op.row_stride != f_op.row_stride or
        t_op.nrow != f_op.nrow or t_op.ncol != f_op.ncol or
        t_op.dtype != f_op.dtype):
      raise ValueError('forecast and truth must share dtype, layout and grid')
    (off_f, off_t), out_dims, out_shape, ngroup = _grouped_tables(
        [f_op, t_op], reduce_dim)
    es = f_op.itemsize
    base = min(f_op.addr, t_op.addr)
    off_f = off_f + (f_op.addr - base) // es
    off_t = off_t + (t_op.addr - base) // es
    nout = off_f.size // ngroup
    shape = tuple(out_shape) + (f_op.nrow, f_op.ncol)
    tensor, ptr = _alloc_maps(ctx, f_op if was_dev else t_op, shape,
                              f_op.dtype)
    code = _lib.F32 if f_op.dtype == np.float32 else _lib.F64
    try:
      ctx.det_maps(base, base, code, stat, nout, ngroup, off_f, off_t,
                   f_op.nrow, f_op.ncol, f_op.row_stride, skipna, ptr)
    except Exception:
      if tensor is None:
        ctx.free(ptr)
      raise
    maps = _fetch_maps(ctx, tensor, ptr, shape, f_op.dtype)
    return maps, tuple(out_dims) + _map_dims(f_op)
  finally:
    for p in staged:
      ctx.free(p)


def run_ens_maps(ctx: _lib.Context, x_op: Operand, t_op: Operand,
                 ens_dim: str, stat_mask: int, reduce_dim: Optional[str],
                 skipna: bool):
  """Runs K6e for one variable.  Returns (maps[nsel, ...], dims, M): the
  selected point-wise ensemble statistics (increasing bit order of
  `stat_mask`) averaged over `reduce_dim`."""
  staged: list = []
  try:
    if ens_dim in t_op.outer_dims:
      raise ValueError(f'truth must not have the {ens_dim!r} dimension')
    was_dev = x_op.on_device
    x_op = _to_device_operand(ctx, x_op, staged)
    t_op = _to_device_operand(ctx, t_op, staged)
    x_op, m, st = split_member_dim(x_op, ens_dim)
    if (t_op.layout != x_op.layout or t_op.row_stride != x_op.row_stride or
        t_op.nrow != x_op.nrow or t_op.ncol != x_op.ncol):
      raise ValueError('forecast and truth must share layout and grid')
    if x_op.dtype != np.float32 or t_op.dtype != np.float32:
      raise ValueError('ensemble kernels take float32 operands')
    (off_x, off_t), out_dims, out_shape, ngroup = _grouped_tables(
        [x_op, t_op], reduce_dim)
    base = min(x_op.addr, t_op.addr)
    off_x = off_x + (x_op.addr - base) // 4
    off_t = off_t + (t_op.addr - base) // 4
    nout = off_x.size // ngroup
    nsel = bin(stat_mask).count('1')
    shape = (nsel,) + tuple(out_shape) + (x_op.nrow, x_op.ncol)
    tensor, ptr = _alloc_maps(ctx, x_op if was_dev else t_op, shape,
                              np.float32)
    try:
      ctx.ens_maps(base, base, _lib.F32, m, st, nout, ngroup, off_x, off_t,
                   x_op.nrow, x_op.ncol, x_op.row_stride, stat_mask, skipna,
                   ptr)
    except Exception:
      if tensor is None:
        ctx.free(ptr)
      raise
    maps = _fetch_maps(ctx, tensor, ptr, shape, np.float32)
    return maps, tuple(out_dims) + _map_dims(x_op), m
  finally:
    for p in staged:
      ctx.free(p)


# ---- K7: threshold / Gaussian metrics -----------------------------------------
THR_NOUT = 8


def _threshold_tables(spec, dims, shape, base, staged, ctx):
  """Device addressing of a threshold spec over the broadcast outer grid.
  spec: ('field', [op per threshold]) or ('gaussian', mean_op, std_op, [z]).
  Returns (nq, ops_on_device, builder) where builder(base) -> kernel args."""
  if spec[0] == 'field':
    seen: dict = {}  # thresholds that share a climatology array share its upload

    def stage(op):
      if op.on_device:
        return op
      if getattr(op, 'gather_terms', None) is not None:
        # a gathered threshold field (day-of-year / quantile lookup): only the
        # slabs it references are uploaded, re-addressed by the returned operand
        return _to_device_operand(ctx, op, staged)
      key = id(op.data)
      if key not in seen:
        seen[key] = _to_device_operand(ctx, op, staged).addr
      return dataclasses.replace(op, addr=seen[key], on_device=True)

    ops = [stage(op) for op in spec[1]]
    return len(ops), ops, lambda b: (
        b, np.ascontiguousarray(np.concatenate([
            offset_table(op, dims, shape) + (op.addr - b) // 4
            for op in ops])), None, None, None)
  m_op = _to_device_operand(ctx, spec[1], staged)
  s_op = _to_device_operand(ctx, spec[2], staged)
  z = np.asarray(spec[3], dtype=np.float64)
  return z.size, [m_op, s_op], lambda b: (
      b, offset_table(m_op, dims, shape) + (m_op.addr - b) // 4, b,
      offset_table(s_op, dims, shape) + (s_op.addr - b) // 4, z)


def _check_same_grid(first: Operand, others: Sequence[Operand]):
  for op in others:
    if (op.layout != first.layout or op.row_stride != first.row_stride or
        op.nrow != first.nrow or op.ncol != first.ncol):
      raise ValueError('operands must share layout, grid and row stride '
                       f'({op.layout}, {op.nrow}x{op.ncol}, {op.row_stride}) vs '
                       f'({first.layout}, {first.nrow}x{first.ncol}, '
                       f'{first.row_stride})')
    if op.dtype != np.float32:
      raise ValueError('threshold / Gaussian kernels take float32 operands')


def run_ens_threshold_metrics(ctx: _lib.Context, x_op: Operand, t_op: Operand,
                              ens_dim: str, spec, latitude, longitude,
                              regions: Sequence, skipna: bool,
                              cell_cache=None):
  """Runs K7 (ensemble entry) for one variable.  Returns (stats, dims, M):
  stats has shape outer_shape + (nq, len(regions), 8)."""
  staged: list = []
  try:
    if ens_dim in t_op.outer_dims:
      raise ValueError(f'truth must not have the {ens_dim!r} dimension')
    x_op = _to_device_operand(ctx, x_op, staged)
    t_op = _to_device_operand(ctx, t_op, staged)
    x_op, m, st = split_member_dim(x_op, ens_dim)
    dims, shape = broadcast_dims(x_op, t_op)
    nq, thr_ops, build = _threshold_tables(spec, dims, shape, 0, staged, ctx)
    _check_same_grid(x_op, [t_op] + thr_ops)
    base = min(op.addr for op in [x_op, t_op] + thr_ops)
    off_x = offset_table(x_op, dims, shape) + (x_op.addr - base) // 4
    off_t = offset_table(t_op, dims, shape) + (t_op.addr - base) // 4
    thr_a, off_a, thr_b, off_b, z = build(base)
    nfield = off_x.size
    nreg = len(regions)
    groups = build_weights(ctx, np.asarray(latitude), np.asarray(longitude),
                           regions, x_op.layout, x_op.row_stride, cell_cache)
    res = np.empty((nfield, nq, nreg, THR_NOUT), dtype=np.float64)
    for ids, wspec in groups:
      out_dev = ctx.malloc(nfield * nq * len(ids) * THR_NOUT * 8)
      try:
        ctx.ens_threshold_metrics(base, base, m, st, off_x, off_t, nq, thr_a,
                                  off_a, thr_b, off_b, z, wspec, skipna,
                                  out_dev)
        res[:, :, ids, :] = ctx.from_device(
            out_dev, (nfield, nq, len(ids), THR_NOUT), np.float64)
      finally:
        ctx.free(out_dev)
    return res.reshape(shape + (nq, nreg, THR_NOUT)), dims, m
  finally:
    for p in staged:
      ctx.free(p)


def run_gaussian_metrics(ctx: _lib.Context, m_op: Operand, s_op: Operand,
                         t_op: Operand, spec, latitude, longitude,
                         regions: Sequence, skipna: bool, cell_cache=None):
  """Runs K7 (Gaussian entry) for one variable; spec None -> CRPS / variance.
  Returns (stats, dims): stats has shape outer_shape + (max(nq, 1), R, 8)."""
  staged: list = []
  try:
    m_op = _to_device_operand(ctx, m_op, staged)
    s_op = _to_device_operand(ctx, s_op, staged)
    t_op = _to_device_operand(ctx, t_op, staged)
    dims, shape = broadcast_dims(m_op, s_op, t_op)
    if spec is not None:
      nq, thr_ops, build = _threshold_tables(spec, dims, shape, 0, staged, ctx)
    else:
      nq, thr_ops, build = 0, [], lambda b: (None, None, None, None, None)
    _check_same_grid(m_op, [s_op, t_op] + thr_ops)
    base = min(op.addr for op in [m_op, s_op, t_op] + thr_ops)
    off_m = offset_table(m_op, dims, shape) + (m_op.addr - base) // 4
    off_s = offset_table(s_op, dims, shape) + (s_op.addr - base) // 4
    off_t = offset_table(t_op, dims, shape) + (t_op.addr - base) // 4
    thr_a, off_a, thr_b, off_b, z = build(base)
    nfield = off_m.size
    nreg = len(regions)
    nq_out = max(nq, 1)
    groups = build_weights(ctx, np.asarray(latitude), np.asarray(longitude),
                           regions, m_op.layout, m_op.row_stride, cell_cache)
    res = np.empty((nfield, nq_out, nreg, THR_NOUT), dtype=np.float64)
    for ids, wspec in groups:
      out_dev = ctx.malloc(nfield * nq_out * len(ids) * THR_NOUT * 8)
      try:
        ctx.gaussian_metrics(base, base, base, off_m, off_s, off_t, nq, thr_a,
                             off_a, thr_b, off_b, z, wspec, skipna, out_dev)
        res[:, :, ids, :] = ctx.from_device(
            out_dev, (nfield, nq_out, len(ids), THR_NOUT), np.float64)
      finally:
        ctx.free(out_dev)
    return res.reshape(shape + (nq_out, nreg, THR_NOUT)), dims
  finally:
    for p in staged:
      ctx.free(p)


def run_ens_threshold_maps(ctx: _lib.Context, x_op: Operand, t_op: Operand,
                           ens_dim: str, spec, stat: int,
                           reduce_dim: Optional[str], skipna: bool):
  """Runs the map-output form of K7 for one variable.  Returns (maps, dims,
  M): maps[nq, ...outer (without reduce_dim)..., nrow, ncol] float32."""
  staged: list = []
  try:
    if ens_dim in t_op.outer_dims:
      raise ValueError(f'truth must not have the {ens_dim!r} dimension')
    was_dev = x_op.on_device
    x_op = _to_device_operand(ctx, x_op, staged)
    t_op = _to_device_operand(ctx, t_op, staged)
    x_op, m, st = split_member_dim(x_op, ens_dim)
    dims, shape = broadcast_dims(x_op, t_op)
    if reduce_dim is not None and reduce_dim in dims:
      i = dims.index(reduce_dim)
      ngroup = shape[i]
      out_dims = dims[:i] + dims[i + 1:]
      out_shape = shape[:i] + shape[i + 1:]
      dims, shape = out_dims + (reduce_dim,), out_shape + (ngroup,)
    else:
      ngroup, out_dims, out_shape = 1, dims, shape
    nq, thr_ops, build = _threshold_tables(spec, dims, shape, 0, staged, ctx)
    _check_same_grid(x_op, [t_op] + thr_ops)
    base = min(op.addr for op in [x_op, t_op] + thr_ops)
    off_x = offset_table(x_op, dims, shape) + (x_op.addr - base) // 4
    off_t = offset_table(t_op, dims, shape) + (t_op.addr - base) // 4
    thr_a, off_a, thr_b, off_b, z = build(base)
    nout = off_x.size // ngroup
    out_full = (nq,) + tuple(out_shape) + (x_op.nrow, x_op.ncol)
    tensor, ptr = _alloc_maps(ctx, x_op if was_dev else t_op, out_full,
                              np.float32)
    try:
      ctx.ens_threshold_maps(base, base, m, st, nout, ngroup, off_x, off_t, nq,
                             thr_a, off_a, thr_b, off_b, z, x_op.nrow,
                             x_op.ncol, x_op.row_stride, stat, skipna, ptr)
    except Exception:
      if tensor is None:
        ctx.free(ptr)
      raise
    maps = _fetch_maps(ctx, tensor, ptr, out_full, np.float32)
    return maps, tuple(out_dims) + _map_dims(x_op), m
  finally:
    for p in staged:
      ctx.free(p)
