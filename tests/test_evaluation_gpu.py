"""GPU tests of the per-chunk dispatch loop (evaluation._metric_and_region_loop
/ evaluate_in_memory, weatherbench2/evaluation.py:388-517) against the oracle,
including BASELINE.json configs[0] (RMSE on synthetic 64x64 via
evaluate_in_memory) and the chunked == unchunked property of
weatherbench2/evaluation_test.py:110-128."""
import numpy as np
import pytest

from oracle import wb2_oracle as orc

pytestmark = pytest.mark.gpu


def _mock(nlat=64, nlon=64, ntime=6, nlead=3, levels=(500, 850), seed=0,
          variables=('geopotential', 'temperature')):
  from weatherbench2_b200 import xarray_lite as xl
  lat = np.linspace(-90, 90, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  times = (np.datetime64('2020-01-01', 'ns') +
           np.arange(ntime + nlead) * np.timedelta64(1, 'D'))
  init = times[:ntime]
  lead = np.arange(nlead) * np.timedelta64(1, 'D').astype('timedelta64[ns]')
  rs_t = np.random.RandomState(seed)        # utils.random_like recipe
  rs_f = np.random.RandomState(seed + 1)
  tdims = ('time', 'level', 'longitude', 'latitude')
  fdims = ('prediction_timedelta', 'time', 'level', 'longitude', 'latitude')
  tv = {v: rs_t.normal(size=(times.size, len(levels), nlon, nlat)).astype(
      np.float32) for v in variables}
  fv = {v: rs_f.normal(size=(nlead, ntime, len(levels), nlon, nlat)).astype(
      np.float32) for v in variables}
  coords = {'level': np.array(levels), 'latitude': lat, 'longitude': lon}
  truth = xl.Dataset({v: (tdims, a) for v, a in tv.items()},
                     dict(coords, time=times))
  forecast = xl.Dataset({v: (fdims, a) for v, a in fv.items()},
                        dict(coords, time=init, prediction_timedelta=lead))
  return forecast, truth, fv, tv, lat, lon, fdims, tdims


def test_config0_rmse_64x64_evaluate_in_memory(tmp_path):
  """BASELINE.json configs[0]."""
  from weatherbench2_b200 import config, evaluation, metrics
  forecast, truth, fv, tv, lat, lon, fdims, tdims = _mock(
      levels=(500,), variables=('geopotential',))
  data_config = config.Data(
      selection=config.Selection(variables=['geopotential'],
                                 time_slice=slice(None, None)),
      paths=config.Paths(forecast=forecast, obs=truth,
                         output_dir=str(tmp_path)), by_init=True)
  eval_configs = {'deterministic': config.Eval(
      metrics={'rmse': metrics.RMSESqrtBeforeTimeAvg(), 'mse': metrics.MSE()})}
  out = evaluation.evaluate_in_memory(data_config, eval_configs)
  res = out['deterministic']['geopotential']
  assert res.dims == ('metric', 'lead_time', 'level')
  assert (tmp_path / 'deterministic.npz').exists()
  # oracle: gather truth at valid time, RMSE per (lead, init, level), time mean
  f = fv['geopotential']
  t = tv['geopotential']
  nlead, ntime = f.shape[:2]
  tg = np.stack([np.stack([t[i + l] for i in range(ntime)])
                 for l in range(nlead)])
  want, wd = orc.rmse_sqrt_before_time_avg(f, fdims, tg, fdims, lat, lon)
  want = want.mean(axis=wd.index('time'))
  np.testing.assert_allclose(res.values[0], want, rtol=2e-6)
  want, wd = orc.mse(f, fdims, tg, fdims, lat, lon)
  np.testing.assert_allclose(res.values[1], want.mean(axis=wd.index('time')),
                             rtol=2e-6)
  saved = np.load(tmp_path / 'deterministic.npz', allow_pickle=True)
  np.testing.assert_array_equal(saved['var:geopotential'], res.values)


def test_metric_and_region_loop_many_metrics_regions():
  from weatherbench2_b200 import config, evaluation, metrics, regions as R
  from weatherbench2_b200 import xarray_lite as xl
  forecast, truth, fv, tv, lat, lon, fdims, tdims = _mock(nlat=37, nlon=72)
  forecast = evaluation.apply_time_conventions(forecast, by_init=True)
  truth_sel = evaluation.select_truth_at_valid_time(truth, forecast)
  rs = np.random.RandomState(3)
  doy = np.arange(1, 367)
  cdims = ('dayofyear', 'level', 'longitude', 'latitude')
  cv = {v: rs.normal(size=(366, 2, lon.size, lat.size)).astype(np.float32)
        for v in fv}
  clim = xl.Dataset({v: (cdims, a) for v, a in cv.items()},
                    {'dayofyear': doy, 'level': np.array([500, 850]),
                     'latitude': lat, 'longitude': lon})
  regions = {'global': R.SliceRegion(),
             'tropics': R.SliceRegion(lat_slice=slice(-20, 20)),
             'extra-tropics': R.ExtraTropicalRegion(),
             'europe': R.SliceRegion(lat_slice=slice(35, 75),
                                     lon_slice=[slice(347.5, None),
                                                slice(0, 42.5)])}
  oregions = {'global': orc.SliceRegion(),
              'tropics': orc.SliceRegion(lat_slice=slice(-20, 20)),
              'extra-tropics': orc.ExtraTropicalRegion(),
              'europe': orc.SliceRegion(lat_slice=slice(35, 75),
                                        lon_slice=[slice(347.5, None),
                                                   slice(0, 42.5)])}
  ec = config.Eval(metrics={'mse': metrics.MSE(), 'bias': metrics.Bias(),
                            'acc': metrics.ACC(climatology=clim)},
                   regions=regions)
  ctx_launches = metrics._context().launch_count
  res = evaluation._metric_and_region_loop(forecast, truth_sel, ec,
                                           skipna=False)
  # 3 metrics x 4 regions x 2 variables served by ONE K1 pass (2 kernels)
  assert metrics._context().launch_count - ctx_launches == 2
  g = res['geopotential']
  assert g.dims == ('metric', 'region', 'lead_time', 'level')
  assert list(g.coords['metric'].values) == ['mse', 'bias', 'acc']
  assert list(g.coords['region'].values) == list(regions)
  f = fv['geopotential']
  t = tv['geopotential']
  nlead, ntime = f.shape[:2]
  tg = np.stack([np.stack([t[i + l] for i in range(ntime)])
                 for l in range(nlead)])
  cg = np.stack([np.stack([cv['geopotential'][i + l] for i in range(ntime)])
                 for l in range(nlead)])  # dayofyear = 1 + day index
  for ri, (rname, oreg) in enumerate(oregions.items()):
    want, wd = orc.mse(f, fdims, tg, fdims, lat, lon, region=oreg)
    np.testing.assert_allclose(g.values[0, ri],
                               want.mean(axis=wd.index('time')), rtol=2e-6)
    want, wd = orc.bias(f, fdims, tg, fdims, lat, lon, region=oreg)
    np.testing.assert_allclose(g.values[1, ri],
                               want.mean(axis=wd.index('time')), rtol=1e-4,
                               atol=2e-6)
    want, wd = orc.acc(f, fdims, tg, fdims, cg, fdims, lat, lon, region=oreg)
    np.testing.assert_allclose(g.values[2, ri],
                               want.mean(axis=wd.index('time')), rtol=1e-4,
                               atol=2e-6)


def test_chunked_equals_unchunked():
  """Mean over init-time chunks == evaluation of the whole period
  (weatherbench2/evaluation_test.py:110-128)."""
  from weatherbench2_b200 import config, evaluation, metrics, regions as R
  forecast, truth, *_ = _mock(nlat=19, nlon=36, ntime=8)
  forecast = evaluation.apply_time_conventions(forecast, by_init=True)
  ec = config.Eval(metrics={'mse': metrics.MSE(),
                            'rmse': metrics.RMSESqrtBeforeTimeAvg()},
                   regions={'global': R.SliceRegion(),
                            'tropics': R.SliceRegion(lat_slice=slice(-20, 20))})
  whole = evaluation._metric_and_region_loop(
      forecast, evaluation.select_truth_at_valid_time(truth, forecast), ec,
      skipna=False)
  parts = []
  for i0 in range(0, 8, 2):
    fc = forecast.isel(init_time=slice(i0, i0 + 2))
    parts.append(evaluation._metric_and_region_loop(
        fc, evaluation.select_truth_at_valid_time(truth, fc), ec,
        skipna=False, compute_chunk=True))
  for name in ('geopotential', 'temperature'):
    stacked = np.concatenate(
        [p[name].transpose('metric', 'region', 'init_time', 'lead_time',
                           'level').values for p in parts], axis=2)
    np.testing.assert_allclose(stacked.mean(axis=2), whole[name].values,
                               rtol=1e-12)
