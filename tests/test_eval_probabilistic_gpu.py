"""The probabilistic eval configs of scripts/evaluate.py:496-606 through
`evaluation._metric_and_region_loop` on the GPU: 'probabilistic' (CRPS family),
'ensemble_binary' (Brier / debiased Brier / ignorance with shared thresholds)
and 'probabilistic_spatial' (map outputs), by-init layout with the lazily
gathered truth of evaluation.py:475."""
import numpy as np
import pytest

from oracle import wb2_oracle as orc

pytestmark = pytest.mark.gpu


def _mock_ens(nlat=19, nlon=36, ntime=5, nlead=2, nmember=6, seed=3):
  from weatherbench2_b200 import evaluation, xarray_lite as xl
  lat = np.linspace(-90, 90, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  times = (np.datetime64('2020-02-25', 'ns') +
           np.arange(ntime + nlead) * np.timedelta64(1, 'D'))
  lead = np.arange(nlead) * np.timedelta64(1, 'D').astype('timedelta64[ns]')
  levels = np.array([500, 850])
  rs = np.random.RandomState(seed)
  t = rs.normal(size=(times.size, levels.size, nlat, nlon)).astype(np.float32)
  f = rs.normal(size=(nmember, ntime, nlead, levels.size, nlat, nlon)).astype(
      np.float32)
  coords = {'level': levels, 'latitude': lat, 'longitude': lon}
  truth = xl.Dataset({'z': (('time', 'level', 'latitude', 'longitude'), t)},
                     dict(coords, time=times))
  forecast = xl.Dataset(
      {'z': (('realization', 'time', 'prediction_timedelta', 'level',
              'latitude', 'longitude'), f)},
      dict(coords, time=times[:ntime], prediction_timedelta=lead,
           realization=np.arange(nmember)))
  forecast = evaluation.apply_time_conventions(forecast, by_init=True)
  truth_sel = evaluation.select_truth_at_valid_time(truth, forecast)
  clim_mean = rs.normal(scale=0.2, size=(366, levels.size, nlat, nlon)).astype(
      np.float32)
  clim_std = rs.uniform(0.6, 1.4, size=clim_mean.shape).astype(np.float32)
  cdims = ('dayofyear', 'level', 'latitude', 'longitude')
  clim = xl.Dataset({'z': (cdims, clim_mean), 'z_std': (cdims, clim_std)},
                    dict(coords, dayofyear=np.arange(1, 367)))
  # dense truth at valid time for the oracle: (init, lead, level, lat, lon)
  tg = np.stack([np.stack([t[i + l] for l in range(nlead)])
                 for i in range(ntime)])
  doy = np.array([[(times[i + l].astype('datetime64[D]') -
                    times[i + l].astype('datetime64[Y]')).astype(int)
                   for l in range(nlead)] for i in range(ntime)])
  return forecast, truth_sel, clim, f, tg, clim_mean[doy], clim_std[doy], lat, lon


def test_ensemble_binary_config_with_regions():
  from weatherbench2_b200 import config, evaluation, metrics, regions as R
  from weatherbench2_b200 import thresholds
  forecast, truth, clim, f, tg, cm, cs, lat, lon = _mock_ens()
  quantiles = [0.25, 0.75, 0.9]
  thrs = [thresholds.GaussianQuantileThreshold(climatology=clim, quantile=q)
          for q in quantiles]
  regs = {'global': R.SliceRegion(),
          'tropics': R.SliceRegion(lat_slice=slice(-20, 20)),
          'extra-tropics': R.ExtraTropicalRegion()}
  oregs = {'global': None, 'tropics': orc.SliceRegion(lat_slice=slice(-20, 20)),
           'extra-tropics': orc.ExtraTropicalRegion()}
  ec = config.Eval(
      metrics={'brier_score': metrics.EnsembleBrierScore(thrs),
               'debiased_brier_score': metrics.DebiasedEnsembleBrierScore(thrs),
               'ignorance_score': metrics.EnsembleIgnoranceScore(thrs)},
      regions=regs)
  ctx = metrics._context()  # pylint: disable=protected-access
  before = ctx.launch_count
  res = evaluation._metric_and_region_loop(forecast, truth, ec, skipna=False)  # pylint: disable=protected-access
  # 3 metrics x 3 regions x 3 thresholds from one request: thresholds go 2 + 1
  # per kernel pass, + the finalize
  assert ctx.launch_count - before == 3
  z = res['z']
  assert z.dims[:3] == ('metric', 'region', 'quantile')
  assert set(z.dims[3:]) == {'lead_time', 'level'}
  pdims = ('init_time', 'lead_time', 'level', 'latitude', 'longitude')
  fdims = ('realization',) + pdims
  for qi, q in enumerate(quantiles):
    thr = orc.gaussian_quantile_threshold(cm, cs, q)
    point = {
        'brier_score': orc.ens_brier_pointwise(f, tg, thr, 0, False, False),
        'debiased_brier_score': orc.ens_brier_pointwise(f, tg, thr, 0, True,
                                                        False),
        'ignorance_score': orc.ens_ignorance_pointwise(f, tg, thr, 0, False)}
    for mi, mname in enumerate(ec.metrics):
      for ri, rname in enumerate(regs):
        avg, ad = orc.spatial_average(point[mname], pdims, lat, lon,
                                      oregs[rname], False)
        want, wd = orc.time_mean(avg, ad, avg_dim='init_time')
        got = z.isel(metric=mi, region=ri, quantile=qi)
        a, b, _ = orc.align(np.asarray(got.values), got.dims, want, wd)
        finite = np.isfinite(b)
        np.testing.assert_array_equal(np.isfinite(a), finite)
        np.testing.assert_allclose(a[finite], b[finite], rtol=1e-5, atol=1e-6)
  del fdims


def test_probabilistic_and_spatial_configs():
  from weatherbench2_b200 import config, evaluation, metrics, regions as R
  forecast, truth, _, f, tg, _, _, lat, lon = _mock_ens(nmember=5)
  fdims = ('realization', 'init_time', 'lead_time', 'level', 'latitude',
           'longitude')
  pdims = fdims[1:]
  ec = config.Eval(
      metrics={'crps': metrics.CRPS(), 'crps_spread': metrics.CRPSSpread(),
               'crps_skill': metrics.CRPSSkill(),
               'ensemble_mean_mse': metrics.EnsembleMeanMSE(),
               'debiased_ensemble_mean_mse': metrics.DebiasedEnsembleMeanMSE(),
               'ensemble_variance': metrics.EnsembleVariance()},
      regions={'global': R.SliceRegion(),
               'tropics': R.SliceRegion(lat_slice=slice(-20, 20))})
  res = evaluation._metric_and_region_loop(forecast, truth, ec, skipna=False)  # pylint: disable=protected-access
  want, wd = orc.crps(f, fdims, tg, pdims, 'realization', lat, lon,
                      region=orc.SliceRegion(lat_slice=slice(-20, 20)))
  want, wd = orc.time_mean(want, wd, avg_dim='init_time')
  got = res['z'].isel(metric=0, region=1)
  a, b, _ = orc.align(np.asarray(got.values), got.dims, want, wd)
  np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
  # map outputs: no regions, lat / lon kept, time mean fused
  ec = config.Eval(
      metrics={'crps': metrics.SpatialCRPS(),
               'crps_spread': metrics.SpatialCRPSSpread(),
               'crps_skill': metrics.SpatialCRPSSkill(),
               'ensemble_mean_mse': metrics.SpatialEnsembleMeanMSE(),
               'ensemble_variance': metrics.SpatialEnsembleVariance()})
  ctx = metrics._context()  # pylint: disable=protected-access
  before = ctx.launch_count
  res = evaluation._metric_and_region_loop(forecast, truth, ec, skipna=False)  # pylint: disable=protected-access
  z = res['z']
  assert z.dims[0] == 'metric' and {'latitude', 'longitude'} <= set(z.dims)
  assert 'init_time' not in z.dims
  maps = orc.spatial_ens_maps(f.astype(np.float64), fdims,
                              tg.astype(np.float64), pdims, 'realization',
                              False)
  for mi, key in enumerate(('crps', 'spread', 'skill', 'mse', 'variance')):
    w, wdims = maps[key]
    want, wd = orc.time_mean(w, wdims, avg_dim='init_time')
    got = z.isel(metric=mi)
    a, b, _ = orc.align(np.asarray(got.values), got.dims, want, wd)
    np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-6)
  # the five ensemble Spatial* metrics share one pass over the members
  assert ctx.launch_count - before == 1
