"""GPU parity tests for K7: Gaussian-forecast and threshold metrics against the
oracle and the reference's own known answers
(weatherbench2/metrics_test.py:284-532, 987-1390)."""
import numpy as np
import pytest

from oracle import wb2_oracle as orc
import wb2_testdata as td

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _ds(vars, coords):  # pylint: disable=redefined-builtin
  from weatherbench2_b200 import xarray_lite as xl
  return xl.Dataset({k: (d, v) for k, (d, v) in vars.items()}, coords)


def _shift(ds, delta):
  return {'vars': {k: (d, v + delta) for k, (d, v) in ds['vars'].items()},
          'coords': ds['coords']}


def _clim_from_truth(truth, name, kind, quantiles=None, shifts=None):
  """The climatologies the reference's tests build from truth.isel(time=0)
  expanded over dayofyear=366 (metrics_test.py:397-426, 497-521)."""
  dims, arr = truth['vars'][name]
  first = np.take(arr, 0, axis=dims.index('time'))
  sdims = tuple(d for d in dims if d != 'time')
  coords = {k: v for k, v in truth['coords'].items() if k != 'time'}
  coords['dayofyear'] = np.arange(1, 367)
  if kind == 'gaussian':
    rep = np.broadcast_to(first, (366,) + first.shape).copy()
    return _ds({name: (('dayofyear',) + sdims, rep),
                name + '_std': (('dayofyear',) + sdims, rep.copy())}, coords)
  coords['quantile'] = np.asarray(quantiles, dtype=np.float64)
  rep = np.stack([np.broadcast_to(first + s, (366,) + first.shape)
                  for s in shifts])
  return _ds({name + '_quantile': (('quantile', 'dayofyear') + sdims, rep)},
             coords)


KW = dict(variables_3d=[], time_start='2022-01-01', time_stop='2022-01-02')
T2M = '2m_temperature'


def _gauss_pair(error):
  forecast = _shift(td.mock_forecast_data(
      variables_2d=[T2M, T2M + '_std'], lead_stop='1 day', **KW), 1.0 + error)
  truth = _shift(td.mock_truth_data(variables_2d=[T2M], **KW), 1.0)
  return forecast, truth


def test_gaussian_crps_and_variance_known_answers():
  """metrics_test.py:286-304, 348-365."""
  from weatherbench2_b200 import metrics
  forecast = _shift(td.mock_forecast_data(
      variables_3d=[], variables_2d=[T2M, T2M + '_std'],
      time_start='2022-01-01', time_stop='2022-01-02', lead_stop='1 day'), 1.0)
  truth = _shift(td.mock_truth_data(
      variables_3d=[], variables_2d=[T2M], time_start='2022-01-01',
      time_stop='2022-01-20'), 1.02)
  fds, tds = _ds(**forecast), _ds(**truth)
  res = metrics.GaussianCRPS().compute(fds, tds)
  np.testing.assert_allclose(res[T2M].values, [0.23385455, 0.23385455],
                             rtol=1e-6)
  res = metrics.GaussianVariance().compute(fds, tds)
  np.testing.assert_allclose(res[T2M].values, [1.0, 1.0], rtol=1e-6)


@pytest.mark.parametrize('error,expected_1,expected_2',
                         [(0.02, 0.04421, 0.257883), (1e6, 0.70786, 0.707861)])
def test_gaussian_brier_known_answers(error, expected_1, expected_2):
  """metrics_test.py:370-431."""
  from weatherbench2_b200 import metrics, thresholds
  forecast, truth = _gauss_pair(error)
  fds, tds = _ds(**forecast), _ds(**truth)
  thr = thresholds.GaussianQuantileThreshold(
      climatology=_clim_from_truth(truth, T2M, 'gaussian'), quantile=0.8)
  res = metrics.GaussianBrierScore([thr]).compute(fds, tds)
  assert res[T2M].dims[0] == 'quantile'
  assert res.attrs['threshold_method'] == 'GaussianQuantileThreshold'
  np.testing.assert_allclose(res[T2M].values, [[expected_1, expected_1]],
                             rtol=1e-4)
  thr = thresholds.QuantileThreshold(
      climatology=_clim_from_truth(truth, T2M, 'quantile', [0.8], [0.0]),
      quantile=0.8)
  res = metrics.GaussianBrierScore([thr]).compute(fds, tds)
  np.testing.assert_allclose(res[T2M].values, [[expected_2, expected_2]],
                             rtol=1e-4)


@pytest.mark.parametrize('error,expected', [(0.02, 0.236055), (1e6, 1.841019)])
def test_gaussian_ignorance_known_answers(error, expected):
  """metrics_test.py:436-475."""
  from weatherbench2_b200 import metrics, thresholds
  forecast, truth = _gauss_pair(error)
  thr = thresholds.GaussianQuantileThreshold(
      climatology=_clim_from_truth(truth, T2M, 'gaussian'), quantile=0.8)
  res = metrics.GaussianIgnoranceScore([thr]).compute(_ds(**forecast),
                                                      _ds(**truth))
  np.testing.assert_allclose(res[T2M].values, [[expected, expected]],
                             rtol=1e-4)


def _tercile_thresholds(truth0):
  from weatherbench2_b200 import thresholds
  clim = _clim_from_truth(truth0, T2M, 'quantile', [0.33, 0.66, 1.0],
                          [0.0, 1.0, 2.0])
  return [thresholds.QuantileThreshold(climatology=clim, quantile=q)
          for q in (0.33, 0.66, 1.0)]


@pytest.mark.parametrize('error,expected', [(0.02, 0.295746), (1e6, 0.758203)])
def test_gaussian_rps_known_answers(error, expected):
  """metrics_test.py:480-534."""
  from weatherbench2_b200 import metrics
  truth0 = td.mock_truth_data(variables_2d=[T2M], **KW)
  forecast, truth = _gauss_pair(error)
  res = metrics.GaussianRPS(_tercile_thresholds(truth0)).compute(
      _ds(**forecast), _ds(**truth))
  assert 'quantile' not in res[T2M].dims
  np.testing.assert_allclose(res[T2M].values, [expected, expected], rtol=1e-4)


def _ens_pair(error, ens_delta, truth_shift=1.0, time_stop='2022-01-02'):
  kw = dict(KW, time_stop=time_stop)
  forecast = td.mock_forecast_data(variables_2d=[T2M], ensemble_size=4,
                                   lead_stop='1 day', **kw)
  d, v = forecast['vars'][T2M]
  v = v + 1.0 + error + ens_delta * np.arange(-2, 2).reshape(
      (4,) + (1,) * (v.ndim - 1))
  forecast['vars'][T2M] = (d, v.astype(np.float32))
  truth = _shift(td.mock_truth_data(variables_2d=[T2M], **kw), truth_shift)
  return forecast, truth


@pytest.mark.parametrize('error,ens_delta,expected',
                         [(0.0, 0.1, 0.0), (0.0, 1.0, 0.25), (-10.0, 0.1, 1.0)])
def test_ensemble_brier_known_answers(error, ens_delta, expected):
  """metrics_test.py:989-1029."""
  from weatherbench2_b200 import metrics, thresholds
  forecast, truth = _ens_pair(error, ens_delta)
  thr = thresholds.GaussianQuantileThreshold(
      climatology=_clim_from_truth(truth, T2M, 'gaussian'), quantile=0.2)
  res = metrics.EnsembleBrierScore([thr]).compute(_ds(**forecast),
                                                  _ds(**truth))
  np.testing.assert_allclose(res[T2M].values, [[expected, expected]],
                             rtol=1e-4, atol=1e-12)
  assert res.attrs['ensemble_size'] == 4


@pytest.mark.parametrize('skipna', [True, False])
def test_ensemble_brier_nan_propagates_unless_skipna(skipna):
  """metrics_test.py:1031-1108."""
  from weatherbench2_b200 import metrics, thresholds
  forecast, truth = _ens_pair(0.0, 0.1, time_stop='2022-01-03')
  thr = thresholds.GaussianQuantileThreshold(
      climatology=_clim_from_truth(truth, T2M, 'gaussian'), quantile=0.2)
  d, v = forecast['vars'][T2M]
  v_nan = v.copy()
  idx = [slice(None)] * v.ndim
  idx[d.index('latitude')] = 0
  v_nan[tuple(idx)] = np.nan
  f_nan = {'vars': {T2M: (d, v_nan)}, 'coords': forecast['coords']}
  dt, vt = truth['vars'][T2M]
  vt_nan = vt.copy()
  idx = [slice(None)] * vt.ndim
  idx[dt.index('longitude')] = 0
  vt_nan[tuple(idx)] = np.nan
  t_nan = {'vars': {T2M: (dt, vt_nan)}, 'coords': truth['coords']}
  want = [[0.0, 0.0]] if skipna else [[np.nan, np.nan]]
  for f, t in ((f_nan, truth), (forecast, t_nan)):
    res = metrics.EnsembleBrierScore([thr]).compute(_ds(**f), _ds(**t),
                                                    skipna=skipna)
    np.testing.assert_allclose(res[T2M].values, want, atol=1e-12)


def test_ensemble_ignorance_and_rps_known_answers():
  """metrics_test.py:1294-1329, 1334-1390."""
  from weatherbench2_b200 import metrics, thresholds
  for error, expected in ((0.0, 0.0), (-10.0, np.inf)):
    forecast, truth = _ens_pair(error, 0.0)
    thr = thresholds.GaussianQuantileThreshold(
        climatology=_clim_from_truth(truth, T2M, 'gaussian'), quantile=0.2)
    res = metrics.EnsembleIgnoranceScore([thr]).compute(_ds(**forecast),
                                                        _ds(**truth))
    np.testing.assert_allclose(res[T2M].values, [[expected, expected]],
                               rtol=1e-4)
  truth0 = td.mock_truth_data(variables_2d=[T2M], **KW)
  for error, expected in ((0.02, 0.0), (-2.0, 2.0)):
    forecast, truth = _ens_pair(error, 0.0, truth_shift=1.5)
    res = metrics.EnsembleRPS(_tercile_thresholds(truth0)).compute(
        _ds(**forecast), _ds(**truth))
    np.testing.assert_allclose(res[T2M].values, [expected, expected],
                               rtol=1e-4, atol=1e-12)


# ---- random data against the oracle ------------------------------------------
def _random_case(ensemble_size, skipna, seed=5):
  truth, forecast = td.get_random_truth_and_forecast(
      variables=['geopotential'], ensemble_size=ensemble_size,
      lead_stop='1 day', time_stop='2019-12-03', time_resolution='12 hours',
      spatial_resolution_in_degrees=30, seed=seed)
  for ds in (truth, forecast):
    for k, (d, v) in ds['vars'].items():
      ds['vars'][k] = (d, v.astype(np.float32))
  if skipna:
    forecast = td.insert_nan(forecast, 0.05, seed=1)
    truth = td.insert_nan(truth, 0.02, seed=2)
    for ds in (truth, forecast):
      for k, (d, v) in ds['vars'].items():
        ds['vars'][k] = (d, v.astype(np.float32))
  # a climatology that depends on day of year, level and position
  rs = np.random.RandomState(seed + 7)
  tdm, _ = truth['vars']['geopotential']
  sdims = tuple(d for d in tdm if d != 'time')
  sshape = tuple(truth['coords'][d].size for d in sdims)
  coords = {k: v for k, v in truth['coords'].items() if k != 'time'}
  coords['dayofyear'] = np.arange(1, 367)
  mean = rs.normal(scale=0.3, size=(366,) + sshape).astype(np.float32)
  std = rs.uniform(0.5, 1.5, size=(366,) + sshape).astype(np.float32)
  clim = _ds({'geopotential': (('dayofyear',) + sdims, mean),
              'geopotential_std': (('dayofyear',) + sdims, std)}, coords)
  return truth, forecast, clim, mean, std, sdims


def _threshold_arrays(truth, mean, std, sdims, quantiles):
  """Oracle thresholds [(array, dims)] per quantile, dims = truth dims."""
  import pandas as pd
  tdm, _ = truth['vars']['geopotential']
  doy = pd.DatetimeIndex(truth['coords']['time']).dayofyear.values - 1
  out = []
  for q in quantiles:
    thr = orc.gaussian_quantile_threshold(mean[doy], std[doy], q)
    out.append((thr, ('time',) + sdims))
  return out


@pytest.mark.parametrize('ensemble_size,skipna', [(2, False), (5, True),
                                                  (10, False), (50, True)])
def test_ensemble_threshold_metrics_match_oracle(ensemble_size, skipna):
  from weatherbench2_b200 import metrics, regions as R, thresholds
  truth, forecast, clim, mean, std, sdims = _random_case(ensemble_size, skipna)
  quantiles = [0.1, 0.3, 0.5, 0.8, 0.95]  # 5 thresholds: passes of 4 + 1
  thrs = [thresholds.GaussianQuantileThreshold(climatology=clim, quantile=q)
          for q in quantiles]
  fds, tds = _ds(**forecast), _ds(**truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  lat, lon = truth['coords']['latitude'], truth['coords']['longitude']
  ax = fd.index('realization')
  od = tuple(d for d in fd if d != 'realization')
  preg = [None, R.SliceRegion(lat_slice=slice(-30, 60),
                              lon_slice=slice(30, 200))]
  oreg = [None, orc.SliceRegion(lat_slice=slice(-30, 60),
                                lon_slice=slice(30, 200))]
  want = {k: [] for k in ('brier', 'debiased', 'ignorance', 'rps')}
  for thr, thr_dims in _threshold_arrays(truth, mean, std, sdims, quantiles):
    # broadcast truth / thresholds against the forecast (time x lead x ...)
    ta, _, d1 = orc.align(t, tdm, np.take(f, 0, axis=ax), od)
    tha, _, d2 = orc.align(thr, thr_dims, np.take(f, 0, axis=ax), od)
    perm = [d1.index(d) for d in od]
    ta = np.transpose(ta, perm)
    tha = np.transpose(tha, [d2.index(d) for d in od])
    want['brier'].append(orc.ens_brier_pointwise(f, ta, tha, ax, False,
                                                 skipna))
    want['debiased'].append(orc.ens_brier_pointwise(f, ta, tha, ax, True,
                                                    skipna))
    want['ignorance'].append(orc.ens_ignorance_pointwise(f, ta, tha, ax,
                                                         skipna))
    want['rps'].append(orc.ens_rps_part_pointwise(f, ta, tha, ax, skipna))
  classes = {'brier': metrics.EnsembleBrierScore,
             'debiased': metrics.DebiasedEnsembleBrierScore,
             'ignorance': metrics.EnsembleIgnoranceScore,
             'rps': metrics.EnsembleRPS}
  ctx = metrics._context()  # pylint: disable=protected-access
  with metrics.batch(preg):
    before = ctx.launch_count
    for key, cls in classes.items():
      for p, o in zip(preg, oreg):
        got = cls(thrs).compute_chunk(fds, tds, region=p,
                                      skipna=skipna)['geopotential']
        per_q = [orc.spatial_average(w, od, lat, lon, o, skipna)
                 for w in want[key]]
        if key == 'rps':
          exp, ed = sum(a for a, _ in per_q), per_q[0][1]
          a, b, _ = orc.align(got.values, got.dims, exp, ed)
        else:
          assert got.dims[0] == 'quantile'
          np.testing.assert_array_equal(got.coords['quantile'].values,
                                        quantiles)
          exp = np.stack([a for a, _ in per_q])
          a, b, _ = orc.align(got.values, got.dims, exp,
                              ('quantile',) + per_q[0][1])
        finite = np.isfinite(b)
        np.testing.assert_array_equal(np.isfinite(a), finite)
        np.testing.assert_allclose(a[finite], b[finite], rtol=RTOL, atol=1e-6)
        np.testing.assert_array_equal(a[~finite], b[~finite])
    # 4 metrics x 2 regions from one pass: 2 threshold launches + finalize
    assert ctx.launch_count - before == 3


@pytest.mark.parametrize('skipna', [False, True])
def test_gaussian_metrics_match_oracle(skipna):
  from weatherbench2_b200 import metrics, regions as R, thresholds
  truth, forecast, clim, mean, std, sdims = _random_case(None, skipna, seed=9)
  rs = np.random.RandomState(3)
  fd, f = forecast['vars']['geopotential']
  s = rs.uniform(0.3, 2.0, size=f.shape).astype(np.float32)
  forecast['vars']['geopotential_std'] = (fd, s)
  tdm, t = truth['vars']['geopotential']
  fds, tds = _ds(**forecast), _ds(**truth)
  lat, lon = truth['coords']['latitude'], truth['coords']['longitude']
  region, oregion = R.ExtraTropicalRegion(), orc.ExtraTropicalRegion()
  ta, _, d1 = orc.align(t, tdm, f, fd)
  ta = np.transpose(ta, [d1.index(d) for d in fd])

  def check(got, pointwise):
    exp, ed = orc.spatial_average(pointwise, fd, lat, lon, oregion, skipna)
    a, b, _ = orc.align(got.values, got.dims, exp, ed)
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=1e-7, equal_nan=True)

  check(metrics.GaussianCRPS().compute_chunk(
      fds, tds, region=region, skipna=skipna)['geopotential'],
        orc.gaussian_crps_pointwise(f, s, ta))
  check(metrics.GaussianVariance().compute_chunk(
      fds, tds, region=region, skipna=skipna)['geopotential'],
        (s * s).astype(np.float64))
  quantiles = [0.2, 0.5, 0.9]
  thrs = [thresholds.GaussianQuantileThreshold(climatology=clim, quantile=q)
          for q in quantiles]
  fns = {metrics.GaussianBrierScore: orc.gaussian_brier_pointwise,
         metrics.GaussianIgnoranceScore: orc.gaussian_ignorance_pointwise,
         metrics.GaussianRPS: orc.gaussian_rps_part_pointwise}
  for cls, fn in fns.items():
    got = cls(thrs).compute_chunk(fds, tds, region=region,
                                  skipna=skipna)['geopotential']
    per_q = []
    for thr, thr_dims in _threshold_arrays(truth, mean, std, sdims,
                                           quantiles):
      tha, _, d2 = orc.align(thr, thr_dims, f, fd)
      tha = np.transpose(tha, [d2.index(d) for d in fd])
      per_q.append(orc.spatial_average(fn(f, s, ta, tha), fd, lat, lon,
                                       oregion, skipna))
    if cls is metrics.GaussianRPS:
      a, b, _ = orc.align(got.values, got.dims, sum(x for x, _ in per_q),
                          per_q[0][1])
    else:
      a, b, _ = orc.align(got.values, got.dims,
                          np.stack([x for x, _ in per_q]),
                          ('quantile',) + per_q[0][1])
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=1e-7, equal_nan=True)


def test_threshold_compute_matches_kernel_selection():
  """`Threshold.compute` (host, like the reference) and the kernel's offset
  tables select the same climatology cells."""
  from weatherbench2_b200 import thresholds
  truth, _, clim, mean, std, sdims = _random_case(None, False)
  tds = _ds(**truth)
  thr = thresholds.GaussianQuantileThreshold(climatology=clim, quantile=0.7)
  got = thr.compute(tds)['geopotential']
  (want, wd), = _threshold_arrays(truth, mean, std, sdims, [0.7])
  a, b, _ = orc.align(got.values, got.dims, want, wd)
  np.testing.assert_array_equal(a, b)
  assert got.values.dtype == np.float64
  with pytest.raises(KeyError):
    thresholds.QuantileThreshold(climatology=clim, quantile=0.7).compute(tds)


@pytest.mark.parametrize('ensemble_size,skipna', [(3, False), (10, True)])
def test_spatial_threshold_maps_match_oracle(ensemble_size, skipna):
  """SpatialEnsembleBrierScore & co. (metrics.py:1615-1637, 1701-1710,
  1768-1790, 1868-1891): per-time maps and the fused time mean."""
  from weatherbench2_b200 import metrics, thresholds
  truth, forecast, clim, mean, std, sdims = _random_case(ensemble_size, skipna)
  quantiles = [0.2, 0.5, 0.8, 0.9, 0.97]
  thrs = [thresholds.GaussianQuantileThreshold(climatology=clim, quantile=q)
          for q in quantiles]
  fds, tds = _ds(**forecast), _ds(**truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  ax = fd.index('realization')
  od = tuple(d for d in fd if d != 'realization')
  want = {k: [] for k in ('brier', 'debiased', 'ignorance', 'rps')}
  for thr, thr_dims in _threshold_arrays(truth, mean, std, sdims, quantiles):
    ta, _, d1 = orc.align(t, tdm, np.take(f, 0, axis=ax), od)
    tha, _, d2 = orc.align(thr, thr_dims, np.take(f, 0, axis=ax), od)
    ta = np.transpose(ta, [d1.index(d) for d in od])
    tha = np.transpose(tha, [d2.index(d) for d in od])
    want['brier'].append(orc.ens_brier_pointwise(f, ta, tha, ax, False,
                                                 skipna))
    want['debiased'].append(orc.ens_brier_pointwise(f, ta, tha, ax, True,
                                                    skipna))
    want['ignorance'].append(orc.ens_ignorance_pointwise(f, ta, tha, ax,
                                                         skipna))
    want['rps'].append(orc.ens_rps_part_pointwise(f, ta, tha, ax, skipna))
  classes = {'brier': metrics.SpatialEnsembleBrierScore,
             'debiased': metrics.SpatialDebiasedEnsembleBrierScore,
             'ignorance': metrics.SpatialEnsembleIgnoranceScore,
             'rps': metrics.SpatialEnsembleRPS}

  def compare(got, exp, exp_dims):
    a, b, _ = orc.align(np.asarray(got.values), got.dims, exp, exp_dims)
    finite = np.isfinite(b)
    np.testing.assert_array_equal(np.isfinite(a), finite)
    np.testing.assert_allclose(a[finite], b[finite], rtol=RTOL, atol=1e-6)
    np.testing.assert_array_equal(a[~finite], b[~finite])

  for key, cls in classes.items():
    got = cls(thrs).compute_chunk(fds, tds, skipna=skipna)['geopotential']
    if key == 'rps':
      exp, ed = sum(want[key]), od
    else:
      exp, ed = np.stack(want[key]), ('quantile',) + od
      assert got.dims[0] == 'quantile'
    compare(got, exp, ed)
    # time mean fused into the kernel
    res = cls(thrs).compute(fds, tds, skipna=skipna)
    assert res.attrs['ensemble_size'] == ensemble_size
    mean_exp, md = orc.time_mean(exp, ed, skipna=skipna, avg_dim='time')
    if key == 'ignorance' and not np.isfinite(mean_exp).all():
      # inf - inf never happens (scores are >= 0) but inf means stay inf
      pass
    compare(res['geopotential'], mean_exp, md)
