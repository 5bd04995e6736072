"""GPU parity tests for K6 / K6e: the map-output (Spatial*) metrics and their
fused time mean, against the oracle (weatherbench2/metrics.py:304-374, 718-772,
1244-1266, 1366-1399 and Metric.compute :117-138)."""
import numpy as np
import pytest

from oracle import wb2_oracle as orc
import wb2_testdata as td

pytestmark = pytest.mark.gpu


def _ds(vars_, coords):
  from weatherbench2_b200 import xarray_lite as xl
  return xl.Dataset({k: (d, v) for k, (d, v) in vars_.items()}, coords)


def _cmp(got, want, wd, **tol):
  a, b, _ = orc.align(np.asarray(got.values), got.dims, want, wd)
  np.testing.assert_allclose(a, b, **tol)


def _det_pair(nlat, nlon, dtype, nan_frac=0.0, seed=0, ntime=5, nlev=3):
  rs = np.random.RandomState(seed)
  dims = ('time', 'level', 'latitude', 'longitude')
  f = rs.normal(size=(ntime, nlev, nlat, nlon)).astype(dtype)
  t = rs.normal(size=(ntime, nlev, nlat, nlon)).astype(dtype)
  if nan_frac:
    f[rs.uniform(size=f.shape) < nan_frac] = np.nan
    f[:, 0, 0, 0] = np.nan  # a cell that is NaN at every time
  coords = {'time': np.arange(ntime), 'level': np.arange(nlev),
            'latitude': np.linspace(-90, 90, nlat),
            'longitude': np.linspace(0, 360, nlon, endpoint=False)}
  return dims, f, t, coords


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('nlat,nlon', [(9, 16), (7, 33), (32, 64)])
def test_spatial_det_maps_chunk_and_time_mean(dtype, nlat, nlon):
  from weatherbench2_b200 import metrics
  dims, f, t, coords = _det_pair(nlat, nlon, dtype, nan_frac=0.05)
  fds, tds = _ds({'a': (dims, f)}, coords), _ds({'a': (dims, t)}, coords)
  tol = dict(rtol=1e-6, atol=1e-6) if dtype == np.float32 else dict(rtol=1e-12)
  for cls, stat in ((metrics.SpatialBias, 'bias'), (metrics.SpatialMSE, 'mse'),
                    (metrics.SpatialMAE, 'mae')):
    want, wd = orc.spatial_det_map(stat, f, dims, t, dims)
    got = cls().compute_chunk(fds, tds)['a']
    assert got.dims == dims and got.dtype == dtype
    # per-time maps are exact: one subtraction (and one multiply) per cell
    np.testing.assert_array_equal(got.values, want)
    np.testing.assert_array_equal(got.coords['latitude'].values,
                                  coords['latitude'])
    for skipna in (False, True):
      mean, md = orc.time_mean(want, wd, skipna=skipna, avg_dim='time')
      got = cls().compute(fds, tds, skipna=skipna)['a']
      assert got.dims == ('level', 'latitude', 'longitude')
      _cmp(got, mean, md, equal_nan=True, **tol)
      assert np.isnan(got.values[0, 0, 0])


def test_spatial_det_maps_truth_gather_init_lead_layout():
  """forecast (init, lead, lat, lon) against truth.sel(time=valid_time)
  (evaluation.py:475): averaged over init_time, no copy of the gathered truth."""
  from weatherbench2_b200 import evaluation, metrics
  truth, forecast = td.get_random_truth_and_forecast(
      variables=['geopotential'], lead_stop='1 day',
      time_stop='2019-12-04', spatial_resolution_in_degrees=30)
  fds = _ds(forecast['vars'], forecast['coords'])
  tds = _ds(truth['vars'], truth['coords'])
  fds2 = evaluation.apply_time_conventions(fds, by_init=True)
  # forecasts whose valid time is still inside the truth record
  n_ok = int((fds2['valid_time'].values.max(axis=1) <=
              tds['time'].values.max()).sum())
  fds2 = fds2.isel(init_time=slice(0, n_ok))
  tds2 = evaluation.select_truth_at_valid_time(tds, fds2)
  got = metrics.SpatialMSE().compute(fds2, tds2)['geopotential']
  f = np.asarray(fds2['geopotential'].values)
  t = np.asarray(tds2['geopotential'].values)
  want, wd = orc.spatial_det_map('mse', f, fds2['geopotential'].dims, t,
                                 tds2['geopotential'].dims)
  mean, md = orc.time_mean(want, wd, avg_dim='init_time')
  _cmp(got, mean, md, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('ensemble_size', [1, 2, 5, 50])
@pytest.mark.parametrize('skipna', [False, True])
def test_spatial_ensemble_maps(ensemble_size, skipna):
  from weatherbench2_b200 import metrics
  truth, forecast = td.get_random_truth_and_forecast(
      variables=['geopotential'], ensemble_size=ensemble_size,
      lead_stop='1 day', time_stop='2019-12-01T12',
      spatial_resolution_in_degrees=30)
  fd, f = forecast['vars']['geopotential']
  f = f.astype(np.float32)
  tdm, t = truth['vars']['geopotential']
  t = t.astype(np.float32)
  if skipna:
    rs = np.random.RandomState(11)
    f[rs.uniform(size=f.shape) < 0.1] = np.nan
  forecast['vars']['geopotential'] = (fd, f)
  truth['vars']['geopotential'] = (tdm, t)
  fds = _ds(forecast['vars'], forecast['coords'])
  tds = _ds(truth['vars'], truth['coords'])
  want = orc.spatial_ens_maps(f.astype(np.float64), fd, t.astype(np.float64),
                              tdm, 'realization', skipna)
  classes = {
      'skill': metrics.SpatialCRPSSkill, 'spread': metrics.SpatialCRPSSpread,
      'crps': metrics.SpatialCRPS, 'mse': metrics.SpatialEnsembleMeanMSE,
      'variance': metrics.SpatialEnsembleVariance,
      'debiased': metrics.DebiasedSpatialEnsembleMeanMSE}
  tol = dict(rtol=2e-5, atol=2e-6, equal_nan=True)
  ctx = metrics._context()  # pylint: disable=protected-access
  with metrics.batch([None]):
    before = ctx.launch_count
    for key, cls in classes.items():
      w, wd = want[key]
      got = cls().compute_chunk(fds, tds, skipna=skipna)['geopotential']
      assert 'realization' not in got.dims
      assert set(got.dims[-2:]) == {'latitude', 'longitude'}
      _cmp(got, w, wd, **tol)
    assert ctx.launch_count - before == 1  # six maps from one pass
  # time mean fused in (EnsembleMetric.compute, metrics.py:598-607)
  for key in ('crps', 'variance'):
    w, wd = want[key]
    mean, md = orc.time_mean(w, wd, skipna=skipna, avg_dim='time')
    res = classes[key]().compute(fds, tds, skipna=skipna)
    assert res.attrs['ensemble_size'] == ensemble_size
    _cmp(res['geopotential'], mean, md, **tol)


def test_spatial_metrics_on_device_tensors_stay_on_device():
  import torch
  from weatherbench2_b200 import metrics, xarray_lite as xl
  dims, f, t, coords = _det_pair(16, 32, np.float32)
  dev = torch.device('cuda', 0)
  fds = xl.Dataset({'a': (dims, torch.from_numpy(f).to(dev))}, coords)
  tds = xl.Dataset({'a': (dims, torch.from_numpy(t).to(dev))}, coords)
  got = metrics.SpatialMAE().compute(fds, tds)['a']
  assert isinstance(got.data, torch.Tensor) and got.data.is_cuda
  want, wd = orc.spatial_det_map('mae', f, dims, t, dims)
  mean, _ = orc.time_mean(want, wd, avg_dim='time')
  np.testing.assert_allclose(got.data.cpu().numpy(), mean, rtol=1e-6)


def test_spatial_ensemble_more_than_64_members_raises():
  from weatherbench2_b200 import _lib, metrics
  truth, forecast = td.get_random_truth_and_forecast(
      variables=['geopotential'], ensemble_size=65, lead_stop='0 day',
      time_stop='2019-12-01T03', spatial_resolution_in_degrees=30)
  fds = _ds(forecast['vars'], forecast['coords'])
  tds = _ds(truth['vars'], truth['coords'])
  with pytest.raises(_lib.Wb2Error, match='64'):
    metrics.SpatialCRPS().compute_chunk(fds, tds)


def test_headline_grid_time_mean_properties():
  """721 x 1440: the fused time mean equals the mean of the per-time maps and
  SpatialMSE >= SpatialBias**2 cell by cell (Jensen)."""
  from weatherbench2_b200 import metrics
  dims, f, t, coords = _det_pair(721, 1440, np.float32, ntime=4, nlev=2,
                                 seed=3)
  fds, tds = _ds({'a': (dims, f)}, coords), _ds({'a': (dims, t)}, coords)
  mse = metrics.SpatialMSE()
  per_time = mse.compute_chunk(fds, tds)['a'].values
  fused = mse.compute(fds, tds)['a'].values
  np.testing.assert_allclose(fused, per_time.astype(np.float64).mean(axis=0),
                             rtol=1e-6)
  bias = metrics.SpatialBias().compute(fds, tds)['a'].values
  assert (fused >= bias ** 2 - 1e-6).all()
