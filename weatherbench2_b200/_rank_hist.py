"""RankHistogram -- same class as weatherbench2/metrics.py:1894-2042 -- on top of
K10 (csrc/rank_hist.cu).  Imported into `weatherbench2_b200.metrics`.
"""
from __future__ import annotations

import typing as t

import numpy as np

from weatherbench2_b200 import _ensemble as ens
from weatherbench2_b200 import _spatial as sp
from weatherbench2_b200 import metrics as m
from weatherbench2_b200 import xarray_lite as xl

LAT, LON = sp.LAT, sp.LON


class RankHistogram(ens.EnsembleMetric):
  """Histogram of truth's rank with respect to the ensemble members.

  `compute_chunk` gives the one-hot encoding of the rank with a trailing `bins`
  dimension (K + 1 bins, or `num_bins` dividing K + 1); averaging over time
  (`compute`, fused into the kernel) gives the histogram.  NaN values are
  treated as larger than any other; `skipna` is ignored (metrics.py:1911).

  Tie-breaking: the reference perturbs truth and members by uniform noise
  smaller than a quarter of the smallest gap (:1960-1987), which only reorders
  exact ties; the kernel places the truth uniformly among the members equal to
  it.  The draws come from a hash of `seed` (entropy from the OS when None),
  not from NumPy's PCG64 stream: without ties the result is identical to the
  reference, with ties it has the same distribution.
  """

  def __init__(self, ensemble_dim: str = ens.REALIZATION,
               num_bins: t.Optional[int] = None,
               break_ties_randomly: bool = True,
               seed: t.Optional[int] = None):
    super().__init__(ensemble_dim=ensemble_dim)
    self.num_bins = num_bins
    self._break_ties_randomly = break_ties_randomly
    self._seed = seed

  def __hash__(self):
    return id(self)

  def _num_bins_actual(self, ensemble_size: int) -> int:
    """metrics.py:1938-1946."""
    default_n_bins = ensemble_size + 1
    if self.num_bins is None:
      return default_n_bins
    if default_n_bins % self.num_bins:
      raise ValueError(
          f'Cannot bin data with {ensemble_size=} into {self.num_bins} bins')
    return self.num_bins

  def _hist(self, forecast, truth, reduce_dim):
    forecast, truth, native = m._prep(forecast, truth)  # pylint: disable=protected-access
    n_ensemble = ens._get_n_ensemble(forecast, self.ensemble_dim)  # pylint: disable=protected-access
    nbins = self._num_bins_actual(n_ensemble)
    seed = self._seed
    if seed is None:
      seed = int(np.random.SeedSequence().entropy) & (2**64 - 1)
    ctx = m._context()  # pylint: disable=protected-access
    out = xl.Dataset()
    for vi, name in enumerate(m._common_vars(forecast, truth)):  # pylint: disable=protected-access
      f_da, t_da = forecast[name], truth[name]
      if LAT not in f_da.dims or LON not in f_da.dims:
        continue
      f_da, t_da = xl.align_inner(f_da, t_da)
      x_op = sp.prepare_operand(f_da, None, np.float32)
      t_op = sp.prepare_operand(t_da, x_op.layout, np.float32)
      hist, dims = _run(ctx, x_op, t_op, self.ensemble_dim, nbins, reduce_dim,
                        self._break_ties_randomly, seed + vi)
      coords = m._map_coords(dims, f_da, t_da)  # pylint: disable=protected-access
      coords.pop(self.ensemble_dim, None)
      coords['bins'] = xl.Coord(('bins',), np.arange(nbins))
      out[name] = xl.DataArray(hist, tuple(dims) + ('bins',), coords, name)
    return m._finish(out, native)  # pylint: disable=protected-access

  def compute_chunk(self, forecast, truth, region=None, skipna=False):
    del region, skipna  # ignored, like the reference
    return self._hist(forecast, truth, None)

  def compute(self, forecast, truth, region=None, skipna=False):
    del region, skipna
    fc = xl.from_xarray(forecast)
    result = self._hist(forecast, truth, m._avg_dim(fc))  # pylint: disable=protected-access
    return result.assign_attrs(ensemble_size=fc.sizes[self.ensemble_dim])


def _run(ctx, x_op, t_op, ens_dim, nbins, reduce_dim, random_ties, seed):
  """Launches K10 for one variable.  Returns (hist[..., nrow, ncol, nbins],
  dims without `bins`)."""
  staged: list = []
  try:
    if ens_dim in t_op.outer_dims:
      raise ValueError(f'truth must not have the {ens_dim!r} dimension')
    was_dev = x_op.on_device
    x_op = sp._to_device_operand(ctx, x_op, staged)  # pylint: disable=protected-access
    t_op = sp._to_device_operand(ctx, t_op, staged)  # pylint: disable=protected-access
    x_op, nmember, stride = sp.split_member_dim(x_op, ens_dim)
    if (t_op.layout != x_op.layout or t_op.row_stride != x_op.row_stride or
        t_op.nrow != x_op.nrow or t_op.ncol != x_op.ncol):
      raise ValueError('forecast and truth must share layout and grid')
    (off_x, off_t), out_dims, out_shape, ngroup = sp._grouped_tables(  # pylint: disable=protected-access
        [x_op, t_op], reduce_dim)
    nout = off_x.size // ngroup
    shape = tuple(out_shape) + (x_op.nrow, x_op.ncol, nbins)
    tensor, ptr = sp._alloc_maps(ctx, x_op if was_dev else t_op, shape,  # pylint: disable=protected-access
                                 np.float32)
    try:
      ctx.rank_histogram(x_op.addr, t_op.addr, nmember, stride, nout, ngroup,
                         off_x, off_t, x_op.nrow, x_op.ncol, x_op.row_stride,
                         nbins, random_ties, seed, ptr)
    except Exception:
      if tensor is None:
        ctx.free(ptr)
      raise
    hist = sp._fetch_maps(ctx, tensor, ptr, shape, np.float32)  # pylint: disable=protected-access
    return hist, tuple(out_dims) + sp._map_dims(x_op)  # pylint: disable=protected-access
  finally:
    for p in staged:
      ctx.free(p)


def central_reliability(hist):
  """Reliability diagram for central histogram probabilities
  (weatherbench2/metrics.py:2045-2126), computed on an already reduced rank
  histogram (a handful of numbers per variable -- result assembly like the
  final `sum / weight_sum`, not kernel work).

  For N bins the probability that truth was less extreme than the central
  2(k+1) - N%2 bins, k = 0 .. N//2 - 1 + N%2, indexed by the probability a
  perfectly calibrated forecast would give, `desired_prob` (with the integer
  `prob_index` as a coordinate along it).
  """
  native = xl.is_native_xarray(hist)
  ds = xl.from_xarray(hist)
  single = isinstance(ds, xl.DataArray)
  if single:
    ds = xl.Dataset({ds.name or '_': ds})
  n_bins = ds.sizes['bins']
  if n_bins < 3:
    raise ValueError(f'Too few bins. {n_bins=} but should be >= 3')
  half, odd = n_bins // 2, n_bins % 2
  # label-based like the reference's .sel (bins are 0 .. N-1)
  bins = np.asarray(ds.coords['bins'].values) if 'bins' in ds.coords else (
      np.arange(n_bins))
  left_idx = xl.label_slice_indices(bins, slice(None, half - 1))
  right_idx = xl.label_slice_indices(bins, slice(half + odd, None))
  desired = np.ones(left_idx.size)
  if odd:
    desired = np.concatenate(([0.5], desired))
  desired = np.cumsum(desired)
  desired = desired / desired[-1]
  out = xl.Dataset(attrs=ds.attrs)
  for name in ds.keys():
    da = ds[name]
    if 'bins' not in da.dims:
      out[name] = da
      continue
    ax = da.dims.index('bins')
    h = np.moveaxis(np.asarray(da.values), ax, -1)
    # cumulative sum from the centre outwards: left half reversed + right half
    probs = np.cumsum(h[..., left_idx[::-1]] + h[..., right_idx], axis=-1)
    other = tuple(d for d in da.dims if d != 'bins')
    if odd:
      center = h[..., int(xl._lookup(bins, np.array([half]))[0])]  # pylint: disable=protected-access
      probs = np.concatenate([center[..., None], center[..., None] + probs],
                             axis=-1)
      # xr.concat puts the expanded prob_index dimension first
      data, dims = np.moveaxis(probs, -1, 0), ('desired_prob',) + other
    else:
      data = np.moveaxis(probs, -1, ax)
      dims = da.dims[:ax] + ('desired_prob',) + da.dims[ax + 1:]
    coords = {k: c for k, c in da.coords.items() if 'bins' not in c.dims}
    coords['desired_prob'] = xl.Coord(('desired_prob',), desired)
    coords['prob_index'] = xl.Coord(('desired_prob',), np.arange(desired.size))
    out[name] = xl.DataArray(data, dims, coords, name, da.attrs)
  if single:
    result = out[list(out.keys())[0]]
    return xl.to_xarray(xl.Dataset({'_': result}))['_'] if native else result
  return xl.to_xarray(out) if native else out
