"""Randomised differential tests of the operator layer's HOST logic against the
oracle, on the NumPy stand-in context (tests/fake_ctx.py): dimension orders no
dataset convention guarantees (spatial dims anywhere, truth with fewer or
differently ordered dims than the forecast, the ensemble axis in the middle),
float32 / float64, strided views, NaNs with and without skipna, regions.
xarray arithmetic does not care about memory layout; the offset tables, layout
decisions and broadcast rules of `_spatial.py` must not either."""
import numpy as np
import pytest

import fake_ctx
from oracle import wb2_oracle as orc

REGIONS = [
    lambda R: (None, None),
    lambda R: (R.SliceRegion(lat_slice=slice(-40, 60)),
               orc.SliceRegion(lat_slice=slice(-40, 60))),
    lambda R: (R.ExtraTropicalRegion(), orc.ExtraTropicalRegion()),
]


def _case(rs, outer, forecast_only=()):
  from weatherbench2_b200 import regions as R, xarray_lite as xl
  nlat, nlon = int(rs.choice([5, 7, 12])), int(rs.choice([6, 8, 16]))
  lat = np.linspace(-90, 90, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  fnames = list(forecast_only) + [d for d in outer if d not in forecast_only
                                  and rs.rand() < 0.8]
  tnames = [d for d in fnames if d not in forecast_only and rs.rand() < 0.7]
  fdims = fnames + ['latitude', 'longitude']
  tdims = tnames + ['latitude', 'longitude']
  rs.shuffle(fdims)
  rs.shuffle(tdims)
  size = dict(outer, latitude=nlat, longitude=nlon)
  dtype = rs.choice([np.float32, np.float64]) if not forecast_only else (
      np.float32)
  f = rs.normal(size=[size[d] for d in fdims]).astype(dtype)
  t = rs.normal(size=[size[d] for d in tdims]).astype(dtype)
  if rs.rand() < 0.3:
    f[rs.rand(*f.shape) < 0.05] = np.nan
  fv = f
  if rs.rand() < 0.3:  # every second element of a larger array
    big = np.zeros([2 * s for s in f.shape], dtype=dtype)
    view = tuple(slice(0, 2 * s, 2) for s in f.shape)
    big[view] = f
    fv = big[view]
  coords = {'latitude': lat, 'longitude': lon}
  coords.update({d: np.arange(n) for d, n in outer.items()})
  fds = xl.Dataset({'z': (tuple(fdims), fv)},
                   {k: v for k, v in coords.items() if k in fdims})
  tds = xl.Dataset({'z': (tuple(tdims), t)},
                   {k: v for k, v in coords.items() if k in tdims})
  preg, oreg = REGIONS[rs.randint(len(REGIONS))](R)
  return (fds, tds, f, tuple(fdims), t, tuple(tdims), lat, lon, preg, oreg,
          bool(rs.rand() < 0.5))


def _same(got, want, wd, atol):
  assert set(got.dims) == set(wd), (got.dims, wd)
  np.testing.assert_allclose(got.transpose(*wd).values, want, rtol=1e-5,
                             atol=atol, equal_nan=True)


@pytest.mark.parametrize('seed', range(4))
def test_deterministic_metrics_any_dimension_order(seed):
  from weatherbench2_b200 import metrics
  rs = np.random.RandomState(seed)
  for _ in range(25):
    outer = {'time': rs.randint(1, 4), 'level': rs.randint(1, 3),
             'lead_time': rs.randint(1, 3)}
    fds, tds, f, fd, t, td, lat, lon, preg, oreg, skipna = _case(rs, outer)
    with fake_ctx.installed():
      mse = metrics.MSE().compute_chunk(fds, tds, region=preg,
                                        skipna=skipna)['z']
      bias = metrics.Bias().compute_chunk(fds, tds, region=preg,
                                          skipna=skipna)['z']
    want, wd = orc.mse(f, fd, t, td, lat, lon, region=oreg, skipna=skipna)
    _same(mse, want, wd, 1e-7)
    want, wd = orc.bias(f, fd, t, td, lat, lon, region=oreg, skipna=skipna)
    _same(bias, want, wd, 1e-6)


@pytest.mark.parametrize('seed', range(4))
def test_ensemble_metrics_any_dimension_order(seed):
  from weatherbench2_b200 import metrics
  rs = np.random.RandomState(100 + seed)
  for _ in range(20):
    outer = {'time': rs.randint(1, 4), 'level': rs.randint(1, 3),
             'realization': int(rs.choice([1, 2, 3, 5, 10]))}
    fds, tds, f, fd, t, td, lat, lon, preg, oreg, skipna = _case(
        rs, outer, forecast_only=('realization',))
    with fake_ctx.installed():
      crps = metrics.CRPS().compute_chunk(fds, tds, region=preg,
                                          skipna=skipna)['z']
      var = metrics.EnsembleVariance().compute_chunk(fds, tds, region=preg,
                                                     skipna=skipna)['z']
      maps = metrics.SpatialCRPS().compute_chunk(fds, tds, skipna=skipna)['z']
    want, wd = orc.crps(f, fd, t, td, 'realization', lat, lon, region=oreg,
                        skipna=skipna)
    _same(crps, want, wd, 1e-6)
    want, wd = orc.ensemble_variance(f, fd, 'realization', lat, lon,
                                     region=oreg, skipna=skipna)
    _same(var, want, wd, 1e-6)
    want, wd = orc.spatial_ens_maps(f, fd, t, td, 'realization',
                                    skipna)['crps']
    _same(maps, want, wd, 1e-6)


def test_scattered_slabs_are_compacted_not_rejected():
  """(time, latitude, level, longitude): another dimension sits INSIDE the
  (row, col) slab; a longitude box or every second row leaves gaps between
  rows.  Forecast and truth that differed in such details used to be refused
  ('must share ... row stride'); now every operand whose slab is not one dense
  block is packed, and views made of whole dense slabs stay zero-copy."""
  from weatherbench2_b200 import _spatial as sp, xarray_lite as xl
  a = np.arange(2 * 5 * 3 * 8, dtype=np.float32).reshape(2, 5, 3, 8)
  op = sp.prepare_operand(xl.DataArray(
      a, ('time', 'latitude', 'level', 'longitude')))
  assert (op.layout, op.nrow, op.ncol, op.row_stride) == ('lat_lon', 5, 8, 8)
  assert not np.shares_memory(op.data, a)
  assert op.outer_dims == ('time', 'level')
  np.testing.assert_array_equal(op.data, np.moveaxis(a, 2, 1))
  box = np.zeros((2, 3, 5, 16), np.float32)[..., 4:12]  # longitude box
  op = sp.prepare_operand(xl.DataArray(
      box, ('time', 'level', 'latitude', 'longitude')))
  assert op.row_stride == op.ncol == 8 and not np.shares_memory(op.data, box)
  # whole dense slabs, strided / broadcast OUTER dimensions: nothing is copied
  b = np.zeros((6, 3, 5, 8), np.float32)
  op = sp.prepare_operand(xl.DataArray(
      b[::2, 1:], ('time', 'level', 'latitude', 'longitude')))
  assert np.shares_memory(op.data, b) and op.outer_strides == (240, 40)
  bc = np.broadcast_to(b[0, 0][None], (4, 5, 8))  # stride-0 outer dim
  op = sp.prepare_operand(xl.DataArray(bc, ('lead_time', 'latitude',
                                            'longitude')))
  assert op.outer_strides == (0,) and np.shares_memory(op.data, b)


@pytest.mark.parametrize('seed', range(3))
def test_spectrum_and_regridder_any_dimension_order(seed):
  """ZonalEnergySpectrum.compute moves the transformed dimension last
  (apply_ufunc's core-dim rule, derived_variables.py:597-626) and
  Regridder.regrid_dataset keeps the input's order (regridding.py:193-209),
  wherever latitude / longitude sit."""
  from weatherbench2_b200 import (derived_variables as dvs, regridding as rg,
                                  xarray_lite as xl)
  rs = np.random.RandomState(200 + seed)
  for _ in range(12):
    nlat, nlon = int(rs.choice([9, 13, 19])), int(rs.choice([16, 24, 36]))
    lat = np.linspace(-90, 90, nlat)
    lon = np.linspace(0, 360, nlon, endpoint=False)
    dims = ['time', 'level', 'latitude', 'longitude']
    rs.shuffle(dims)
    dims = tuple(dims)
    size = dict(time=rs.randint(1, 4), level=rs.randint(1, 3), latitude=nlat,
                longitude=nlon)
    x = rs.standard_normal([size[d] for d in dims]).astype(
        rs.choice([np.float32, np.float64]))
    ds = xl.Dataset({'u': (dims, x)}, dict(
        time=np.arange(size['time']), level=np.arange(size['level']),
        latitude=lat, longitude=lon))
    x32 = x.astype(np.float32)
    with fake_ctx.installed():
      got = dvs.ZonalEnergySpectrum('u').compute(ds)
    want, wd, _, _ = orc.zonal_energy_spectrum(x32, dims, lat, lon)
    assert got.dims == wd
    total = np.abs(want).sum(axis=-1, keepdims=True)
    assert np.max(np.abs(got.values - want) / total) < 1e-5
    target = rg.Grid.from_degrees(
        np.linspace(0, 360, nlon // 2, endpoint=False),
        np.linspace(-90, 90, (nlat + 1) // 2))
    with fake_ctx.installed():
      out = rg.ConservativeRegridder(rg.Grid.from_degrees(lon, lat),
                                     target).regrid_dataset(ds)['u']
    assert out.dims == dims
    outer = [d for d in dims if d not in ('longitude', 'latitude')]
    moved = np.moveaxis(x32, [dims.index('longitude'), dims.index('latitude')],
                        [-2, -1])
    want = orc.conservative_regrid(
        moved, orc.Grid(lon, lat),
        orc.Grid(np.asarray(target.longitudes), np.asarray(target.latitudes)))
    np.testing.assert_allclose(
        out.transpose(*outer, 'longitude', 'latitude').values, want, rtol=2e-5,
        atol=2e-6)


@pytest.mark.parametrize('seed', range(3))
def test_energy_score_rank_histogram_and_maps_any_dimension_order(seed):
  """K3, K10 and K6 through their operators; the deterministic map is fed an
  integer pick of the member axis (a strided view with gaps between rows)."""
  from weatherbench2_b200 import metrics
  rs = np.random.RandomState(300 + seed)
  for _ in range(15):
    outer = {'time': rs.randint(1, 4), 'level': rs.randint(1, 3),
             'realization': int(rs.choice([2, 3, 5, 9]))}
    fds, tds, f, fd, t, td, lat, lon, _, _, _ = _case(
        rs, outer, forecast_only=('realization',))
    f = np.nan_to_num(f)  # K3 has no skipna path of its own
    fds = type(fds)({'z': (fd, f)}, fds.coords)
    with fake_ctx.installed():
      es = metrics.EnergyScore().compute_chunk(fds, tds)['z']
      rh = metrics.RankHistogram().compute_chunk(fds, tds)['z']
      sm = metrics.SpatialMSE().compute_chunk(fds.isel(realization=0),
                                              tds)['z']
    want, wd = orc.energy_score(f, fd, t, td, 'realization', lat, lon)
    _same(es, want, wd, 1e-6)
    want, wd = orc.rank_histogram_one_hot(f, fd, t, td, 'realization')
    assert set(rh.dims) == set(wd)
    np.testing.assert_array_equal(rh.transpose(*wd).values, want)
    f0 = np.take(f, 0, axis=fd.index('realization'))
    want, wd = orc.spatial_det_map(
        'mse', f0, tuple(d for d in fd if d != 'realization'), t, td)
    _same(sm, want, wd, 1e-7)


@pytest.mark.parametrize('seed', range(2))
def test_acc_climatology_lookup_any_dimension_order(seed):
  """ACC (metrics.py:387-414): climatology with or without `hour`, with more
  levels than the forecast, every array with its own dimension order; the
  day-of-year / hour / level lookups are folded into the offset table."""
  import pandas as pd
  from weatherbench2_b200 import metrics, xarray_lite as xl
  rs = np.random.RandomState(400 + seed)
  hour12 = np.timedelta64(12, 'h')
  for _ in range(15):
    nlat, nlon = int(rs.choice([5, 7])), int(rs.choice([6, 8]))
    lat = np.linspace(-90, 90, nlat)
    lon = np.linspace(0, 360, nlon, endpoint=False)
    ntime, nlev = rs.randint(1, 5), rs.randint(1, 3)
    times = np.datetime64('2020-12-30T00', 'ns') + np.arange(ntime) * hour12
    levels = np.array([500, 850])[:nlev]
    clev = np.array([300, 500, 850]) if rs.rand() < 0.5 else levels
    has_hour = rs.rand() < 0.5
    base = ['time', 'level', 'latitude', 'longitude']
    fdims, tdims = list(base), list(base)
    cdims = (['hour'] if has_hour else []) + ['dayofyear'] + base[1:]
    for dims in (fdims, tdims, cdims):
      rs.shuffle(dims)
    size = dict(time=ntime, level=nlev, latitude=nlat, longitude=nlon, hour=2,
                dayofyear=366)
    f = rs.normal(size=[size[d] for d in fdims]).astype(np.float32)
    t = rs.normal(size=[size[d] for d in tdims]).astype(np.float32)
    c = rs.normal(size=[dict(size, level=clev.size)[d] for d in cdims]).astype(
        np.float32)
    coords = dict(latitude=lat, longitude=lon, time=times, level=levels)
    ccoords = dict(latitude=lat, longitude=lon, level=clev,
                   dayofyear=np.arange(1, 367))
    if has_hour:
      ccoords['hour'] = np.array([0, 12])
    skipna = bool(rs.rand() < 0.5)
    with fake_ctx.installed():
      got = metrics.ACC(climatology=xl.Dataset({'z': (tuple(cdims), c)},
                                               ccoords)).compute_chunk(
          xl.Dataset({'z': (tuple(fdims), f)}, coords),
          xl.Dataset({'z': (tuple(tdims), t)}, coords), skipna=skipna)['z']
    lead = (['hour'] if has_hour else []) + ['dayofyear'] + base[1:]
    cc = np.transpose(c, [cdims.index(d) for d in lead])
    pick = [list(clev).index(v) for v in levels]
    rows = []
    for stamp in pd.DatetimeIndex(times):
      slab = cc[stamp.hour // 12] if has_hour else cc
      rows.append(slab[stamp.dayofyear - 1][pick])
    want, wd = orc.acc(f, tuple(fdims), t, tuple(tdims), np.stack(rows),
                       tuple(base), lat, lon, skipna=skipna)
    _same(got, want, wd, 1e-6)
