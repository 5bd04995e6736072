"""GPU parity tests for K9 (SEEPS / SpatialSEEPS, weatherbench2/metrics.py:417-528)
against the oracle and the reference's known answers
(weatherbench2/metrics_test.py:1392-1440)."""
import numpy as np
import pandas as pd
import pytest

from oracle import wb2_oracle as orc
import wb2_testdata as td

pytestmark = pytest.mark.gpu
NAME = 'total_precipitation_24hr'


def _ds(vars, coords):  # pylint: disable=redefined-builtin
  from weatherbench2_b200 import xarray_lite as xl
  return xl.Dataset({k: (d, v) for k, (d, v) in vars.items()}, coords)


def _pair(nday=10, lead_stop='0 day', res=30):
  from weatherbench2_b200 import evaluation
  kw = dict(variables_3d=[], variables_2d=[NAME], time_start='2022-01-01',
            time_stop=str(np.datetime64('2022-01-01') + nday),
            spatial_resolution_in_degrees=res)
  forecast = td.mock_forecast_data(lead_stop=lead_stop, **kw)
  truth = td.mock_truth_data(**kw)
  fds = evaluation.apply_time_conventions(_ds(**forecast), by_init=True)
  tds = _ds(**truth)
  return fds, tds, forecast, truth


def _climatology(truth, dry_fraction, threshold, rs=None):
  dims, arr = truth['vars'][NAME]
  first = np.take(arr, 0, axis=dims.index('time')).astype(np.float32)
  sdims = tuple(d for d in dims if d != 'time')
  shape = (4, 366) + first.shape
  coords = {k: v for k, v in truth['coords'].items() if k != 'time'}
  coords['hour'] = np.array([0, 6, 12, 18])
  coords['dayofyear'] = np.arange(1, 367)
  if rs is None:
    frac = np.broadcast_to(first + dry_fraction, shape).astype(np.float32)
    thr = np.broadcast_to(first + threshold, shape).astype(np.float32)
  else:
    frac = rs.uniform(0.0, 1.0, size=shape).astype(np.float32)
    thr = rs.uniform(0.0005, 0.004, size=shape).astype(np.float32)
  cdims = ('hour', 'dayofyear') + sdims
  return _ds({NAME + '_seeps_dry_fraction': (cdims, frac.copy()),
              NAME + '_seeps_threshold': (cdims, thr.copy())}, coords), frac, thr


def test_seeps_known_answers():
  """metrics_test.py:1392-1440."""
  from weatherbench2_b200 import evaluation, metrics
  fds, tds, forecast, truth = _pair()
  tsel = evaluation.select_truth_at_valid_time(tds, fds)
  clim, _, _ = _climatology(truth, 0.4, 1.0)
  seeps = metrics.SEEPS(climatology=clim)
  res = seeps.compute(fds, tsel)[NAME]
  np.testing.assert_allclose(res.values, 0, atol=1e-4)
  d, v = forecast['vars'][NAME]
  forecast['vars'][NAME] = (d, v + 0.5)
  fds2 = evaluation.apply_time_conventions(_ds(**forecast), by_init=True)
  res = seeps.compute(fds2, tsel)[NAME]
  assert res.dims == ('lead_time',)
  np.testing.assert_allclose(res.values, 1.25, atol=1e-4)
  maps = metrics.SpatialSEEPS(climatology=clim).compute(fds2, tsel)[NAME]
  assert set(maps.dims) == {'lead_time', 'latitude', 'longitude'}
  np.testing.assert_allclose(maps.values, 1.25, atol=1e-4)


@pytest.mark.parametrize('skipna', [False, True])
def test_seeps_random_fields_match_oracle(skipna):
  from weatherbench2_b200 import evaluation, metrics, regions as R
  fds, tds, forecast, truth = _pair(nday=6, lead_stop='1 day', res=20)
  rs = np.random.RandomState(12)
  fd, f = forecast['vars'][NAME]
  tdm, t = truth['vars'][NAME]
  # precipitation-like: many exact zeros, values around both thresholds and
  # values exactly AT the dry threshold (which belong to no category)
  f = np.where(rs.uniform(size=f.shape) < 0.3, 0.0,
               rs.gamma(0.6, 0.002, size=f.shape)).astype(np.float32)
  t = np.where(rs.uniform(size=t.shape) < 0.3, 0.0,
               rs.gamma(0.6, 0.002, size=t.shape)).astype(np.float32)
  f.flat[::17] = np.float32(0.00025)
  t.flat[::23] = np.float32(0.00025)
  f.flat[5::41] = np.nan
  t.flat[7::53] = np.nan
  forecast['vars'][NAME] = (fd, f)
  truth['vars'][NAME] = (tdm, t)
  fds = evaluation.apply_time_conventions(_ds(**forecast), by_init=True)
  tds = _ds(**truth)
  clim, frac, thr = _climatology(truth, 0, 0, rs)
  # oracle inputs on the (lead, init, lon, lat) grid of the forecast
  times = truth['coords']['time']
  vt = (forecast['coords']['time'][None, :] +
        forecast['coords']['prediction_timedelta'][:, None])
  pos = np.searchsorted(times, vt)
  ok = (vt <= times.max()).all(axis=0)
  nok = int(ok.sum())
  fds = fds.isel(init_time=slice(0, nok))
  tsel = evaluation.select_truth_at_valid_time(tds, fds)
  f_o = f[:, :nok]
  t_o = t[pos[:, :nok]]
  stamps = pd.DatetimeIndex(vt[:, :nok].ravel())
  doy = (stamps.dayofyear.values - 1).reshape(vt[:, :nok].shape)
  hour = (stamps.hour.values // 6).reshape(vt[:, :nok].shape)
  wet = thr[hour, doy]
  p1 = frac.mean(axis=(0, 1))
  point = orc.seeps_pointwise(f_o, t_o, wet, wet, p1)
  dims = ('lead_time', 'init_time', 'longitude', 'latitude')
  got = metrics.SpatialSEEPS(climatology=clim).compute_chunk(fds, tsel)[NAME]
  a, b, _ = orc.align(np.asarray(got.values), got.dims, point, dims)
  np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
  np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6, equal_nan=True)
  # fused time mean
  want, wd = orc.time_mean(point, dims, skipna=skipna, avg_dim='init_time')
  got = metrics.SpatialSEEPS(climatology=clim).compute(fds, tsel,
                                                       skipna=skipna)[NAME]
  a, b, _ = orc.align(np.asarray(got.values), got.dims, want, wd)
  np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6, equal_nan=True)
  # SEEPS: weighted spatial mean with skipna = True, with a region
  lat, lon = truth['coords']['latitude'], truth['coords']['longitude']
  avg, ad = orc.spatial_average(point, dims, lat, lon,
                                orc.SliceRegion(lat_slice=slice(-40, 60)), True)
  got = metrics.SEEPS(climatology=clim).compute_chunk(
      fds, tsel, region=R.SliceRegion(lat_slice=slice(-40, 60)))[NAME]
  a, b, _ = orc.align(np.asarray(got.values), got.dims, avg, ad)
  np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6, equal_nan=True)
