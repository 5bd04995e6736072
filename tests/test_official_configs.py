"""The eval configs `scripts/evaluate.py` builds for the official evaluation
(scripts/evaluate.py:336-660, docs/source/official-evaluation.md): all 13
predefined regions plus the three land regions, wind-vector errors, the
wind-speed derived variable, deterministic / spatial / temporal configs on a
deterministic forecast and the probabilistic / experimental / spatial /
histogram configs on an ensemble -- through `evaluate_in_memory`, spot-checked
against the oracle.  CPU: NumPy stand-in context; GPU: the last-collected
tests/test_zz_evaluation_cases_gpu.py."""
import numpy as np

import fake_ctx
from oracle import wb2_oracle as orc

LEVELS = (500, 850)
VARS = ('geopotential', 'u_component_of_wind', 'v_component_of_wind')
H = np.timedelta64(1, 'h').astype('timedelta64[ns]')


def predefined_regions(R, orc_mod, lsm, lat=None, lon=None):
  """scripts/evaluate.py:345-395, for the product and for the oracle (whose
  LandRegion takes the mask's coordinate labels separately)."""
  class _Oracle:
    SliceRegion = orc_mod.SliceRegion
    CombinedRegion = orc_mod.CombinedRegion

    @staticmethod
    def LandRegion(mask):  # pylint: disable=invalid-name
      return orc_mod.LandRegion(np.asarray(mask.values if hasattr(
          mask, 'values') else mask), latitude=lat, longitude=lon)

  def both(make):
    return make(R), make(_Oracle)
  out = {
      'global': both(lambda m: m.SliceRegion()),
      'tropics': both(lambda m: m.SliceRegion(lat_slice=slice(-20, 20))),
      'extra-tropics': both(lambda m: m.SliceRegion(
          lat_slice=[slice(None, -20), slice(20, None)])),
      'northern-hemisphere': both(
          lambda m: m.SliceRegion(lat_slice=slice(20, None))),
      'southern-hemisphere': both(
          lambda m: m.SliceRegion(lat_slice=slice(None, -20))),
      'europe': both(lambda m: m.SliceRegion(
          lat_slice=slice(35, 75),
          lon_slice=[slice(360 - 12.5, None), slice(0, 42.5)])),
      'north-america': both(lambda m: m.SliceRegion(
          lat_slice=slice(25, 60), lon_slice=slice(360 - 120, 360 - 75))),
      'north-atlantic': both(lambda m: m.SliceRegion(
          lat_slice=slice(25, 65), lon_slice=slice(360 - 70, 360 - 10))),
      'north-pacific': both(lambda m: m.SliceRegion(
          lat_slice=slice(25, 60), lon_slice=slice(145, 360 - 130))),
      'east-asia': both(lambda m: m.SliceRegion(
          lat_slice=slice(25, 60), lon_slice=slice(102.5, 150))),
      'ausnz': both(lambda m: m.SliceRegion(
          lat_slice=slice(-45, -12.5), lon_slice=slice(120, 175))),
      'arctic': both(lambda m: m.SliceRegion(lat_slice=slice(60, 90))),
      'antarctic': both(lambda m: m.SliceRegion(lat_slice=slice(-90, -60))),
      'global_land': both(lambda m: m.LandRegion(lsm)),
      'extra-tropics_land': both(lambda m: m.CombinedRegion([
          m.SliceRegion(lat_slice=[slice(None, -20), slice(20, None)]),
          m.LandRegion(lsm)])),
      'tropics_land': both(lambda m: m.CombinedRegion([
          m.SliceRegion(lat_slice=slice(-20, 20)), m.LandRegion(lsm)])),
  }
  return ({k: v[0] for k, v in out.items()}, {k: v[1] for k, v in out.items()})


def _data(ensemble_size=None, seed=0):
  from weatherbench2_b200 import xarray_lite as xl
  rs = np.random.RandomState(seed)
  lat = np.linspace(-90, 90, 19)
  lon = np.linspace(0, 360, 36, endpoint=False)
  ttimes = np.datetime64('2020-01-01', 'ns') + np.arange(16) * 12 * H
  init = ttimes[:5]
  lead = np.array([0, 12, 36]) * H
  tdims = ('time', 'level', 'longitude', 'latitude')
  fdims = ('time', 'prediction_timedelta', 'level', 'longitude', 'latitude')
  fshape = (init.size, lead.size, len(LEVELS), lon.size, lat.size)
  if ensemble_size:
    fdims = ('number',) + fdims
    fshape = (ensemble_size,) + fshape
  tv = {v: rs.standard_normal((ttimes.size, len(LEVELS), lon.size, lat.size)
                              ).astype(np.float32) for v in VARS}
  fv = {v: rs.standard_normal(fshape).astype(np.float32) for v in VARS}
  coords = dict(level=np.array(LEVELS), latitude=lat, longitude=lon)
  truth = xl.Dataset({v: (tdims, a) for v, a in tv.items()},
                     dict(coords, time=ttimes))
  fcoords = dict(coords, time=init, prediction_timedelta=lead)
  if ensemble_size:
    fcoords['number'] = np.arange(ensemble_size)
  forecast = xl.Dataset({v: (fdims, a) for v, a in fv.items()}, fcoords)
  cdims = ('hour', 'dayofyear', 'level', 'longitude', 'latitude')
  # (the ERA5 climatology also holds the derived wind_speed; ACC needs it)
  cv = {v: rs.standard_normal((2, 366, len(LEVELS), lon.size, lat.size)
                              ).astype(np.float32)
        for v in VARS + ('wind_speed',)}
  clim = xl.Dataset({v: (cdims, a) for v, a in cv.items()},
                    dict(coords, hour=np.array([0, 12]),
                         dayofyear=np.arange(1, 367)))
  lsm = (rs.rand(lat.size, lon.size) > 0.6).astype(np.float32)
  vt = init[:, None] + lead[None, :]
  pos = {t: i for i, t in enumerate(ttimes)}
  gather = np.array([[pos[v] for v in row] for row in vt])
  return dict(forecast=forecast, truth=truth, clim=clim, lsm=lsm, lat=lat,
              lon=lon, fv=fv, tv=tv, cv=cv, fdims=fdims, tdims=tdims,
              cdims=cdims, vt=vt, gather=gather)


def _truth_at_valid(d, name):
  return d['tv'][name][d['gather']]  # (time, lead, level, lon, lat)


ODIMS = ('time', 'lead_time', 'level', 'longitude', 'latitude')


def case_official_deterministic_configs(tmp_path, scope):
  from weatherbench2_b200 import (config, derived_variables as dvs, evaluation,
                                  metrics, regions as R)
  from weatherbench2_b200 import xarray_lite as xl
  d = _data()
  lsm_da = xl.DataArray(d['lsm'], ('latitude', 'longitude'),
                        {'latitude': d['lat'], 'longitude': d['lon']})
  regions, oregions = predefined_regions(R, orc, lsm_da, d['lat'], d['lon'])
  wind = dict(u_name='u_component_of_wind', v_name='v_component_of_wind',
              vector_name='wind_vector')
  deterministic = {
      'mse': metrics.MSE(wind_vector_mse=[metrics.WindVectorMSE(**wind)]),
      'acc': metrics.ACC(climatology=d['clim']),
      'bias': metrics.Bias(), 'mae': metrics.MAE()}
  rmse = {'rmse_sqrt_before_time_avg': metrics.RMSESqrtBeforeTimeAvg(
      wind_vector_rmse=[metrics.WindVectorRMSESqrtBeforeTimeAvg(**wind)])}
  spatial = {'bias': metrics.SpatialBias(), 'mse': metrics.SpatialMSE(),
             'mae': metrics.SpatialMAE()}
  derived = {'wind_speed': dvs.DERIVED_VARIABLE_DICT['wind_speed']}
  eval_configs = {
      'deterministic': config.Eval(metrics=deterministic, regions=regions,
                                   derived_variables=derived),
      'deterministic_spatial': config.Eval(metrics=spatial,
                                           derived_variables=derived),
      'deterministic_temporal': config.Eval(
          metrics=deterministic | rmse, regions=regions,
          derived_variables=derived, temporal_mean=False),
  }
  dc = config.Data(
      selection=config.Selection(
          variables=list(VARS) + ['wind_speed'], levels=list(LEVELS),
          time_slice=slice('2020-01-01', '2020-01-03')),
      paths=config.Paths(forecast=d['forecast'], obs=d['truth'],
                         output_dir=str(tmp_path)), by_init=True)
  with scope():
    out = evaluation.evaluate_in_memory(dc, eval_configs)
  names = list(regions)
  det = out['deterministic']
  assert set(det.keys()) == set(VARS) | {'wind_speed', 'wind_vector'}
  g = det['geopotential']
  assert g.dims == ('metric', 'region', 'lead_time', 'level')
  assert list(g.coords['region'].values) == names
  assert list(g.coords['metric'].values) == ['mse', 'acc', 'bias', 'mae']
  f, t = d['fv']['geopotential'], _truth_at_valid(d, 'geopotential')
  # MSE over 'europe' (longitude box wrapping through 0)
  want, wd = orc.mse(f, ODIMS, t, ODIMS, d['lat'], d['lon'],
                     region=oregions['europe'])
  np.testing.assert_allclose(g.values[0, names.index('europe')],
                             want.mean(axis=wd.index('time')), rtol=1e-5)
  # ACC over land
  stamps = d['vt']
  hour = ((stamps - stamps.astype('datetime64[D]')) // (12 * H)).astype(int)
  doy = (stamps.astype('datetime64[D]') -
         stamps.astype('datetime64[Y]')).astype(int)
  c = d['cv']['geopotential'][hour, doy]
  want, wd = orc.acc(f, ODIMS, t, ODIMS, c, ODIMS, d['lat'], d['lon'],
                     region=oregions['global_land'])
  np.testing.assert_allclose(g.values[1, names.index('global_land')],
                             want.mean(axis=wd.index('time')), rtol=1e-5,
                             atol=1e-7)
  # wind vector MSE, global; wind-speed bias over the tropics
  fu, fv_ = d['fv']['u_component_of_wind'], d['fv']['v_component_of_wind']
  tu = _truth_at_valid(d, 'u_component_of_wind')
  tv_ = _truth_at_valid(d, 'v_component_of_wind')
  want, wd = orc.wind_vector_mse(fu, fv_, ODIMS, tu, tv_, ODIMS, d['lat'],
                                 d['lon'])
  wvec = det['wind_vector']
  # xr.merge is an outer join: only the 'mse' entry of wind_vector has values
  assert np.isnan(wvec.values[1:]).all()
  np.testing.assert_allclose(wvec.values[0, 0],
                             want.mean(axis=wd.index('time')), rtol=1e-5)
  fs, ts = np.sqrt(fu**2 + fv_**2), np.sqrt(tu**2 + tv_**2)
  want, wd = orc.bias(fs, ODIMS, ts, ODIMS, d['lat'], d['lon'],
                      region=oregions['tropics'])
  np.testing.assert_allclose(
      det['wind_speed'].values[2, names.index('tropics')],
      want.mean(axis=wd.index('time')), rtol=1e-5, atol=1e-7)
  # temporal: the init_time axis is kept; rmse = sqrt(mse) per time
  tmp = out['deterministic_temporal']['geopotential']
  assert tmp.dims == ('metric', 'region', 'init_time', 'lead_time', 'level')
  assert tmp.sizes['init_time'] == 5
  np.testing.assert_allclose(tmp.values[4], np.sqrt(tmp.values[0]), rtol=1e-6)
  np.testing.assert_allclose(tmp.values[0].mean(axis=1), g.values[0],
                             rtol=1e-12)
  # spatial maps with the time mean fused in
  sp = out['deterministic_spatial']['geopotential']
  assert sp.dims == ('metric', 'lead_time', 'level', 'longitude', 'latitude')
  np.testing.assert_allclose(sp.values[1], ((f - t)**2).mean(axis=0),
                             rtol=1e-5, atol=1e-7)


def case_official_probabilistic_configs(tmp_path, scope):
  from weatherbench2_b200 import config, evaluation, metrics, regions as R
  from weatherbench2_b200 import xarray_lite as xl
  d = _data(ensemble_size=6, seed=1)
  lsm_da = xl.DataArray(d['lsm'], ('latitude', 'longitude'),
                        {'latitude': d['lat'], 'longitude': d['lon']})
  regions, oregions = predefined_regions(R, orc, lsm_da, d['lat'], d['lon'])
  e = dict(ensemble_dim='number')
  eval_configs = {
      'probabilistic': config.Eval(regions=regions, metrics={
          'crps': metrics.CRPS(**e), 'crps_spread': metrics.CRPSSpread(**e),
          'crps_skill': metrics.CRPSSkill(**e),
          'ensemble_mean_mse': metrics.EnsembleMeanMSE(**e),
          'debiased_ensemble_mean_mse': metrics.DebiasedEnsembleMeanMSE(**e),
          'ensemble_variance': metrics.EnsembleVariance(**e)}),
      'ensemble_forecast_vs_era_experimental_metrics': config.Eval(
          regions=regions, metrics={
              'energy_score': metrics.EnergyScore(**e),
              'energy_score_spread': metrics.EnergyScoreSpread(**e),
              'energy_score_skill': metrics.EnergyScoreSkill(**e),
              'ensemble_mean_rmse_sqrt_before_time_avg':
                  metrics.EnsembleMeanRMSESqrtBeforeTimeAvg(**e),
              'ensemble_stddev_sqrt_before_time_avg':
                  metrics.EnsembleStddevSqrtBeforeTimeAvg(**e)}),
      'probabilistic_spatial': config.Eval(metrics={
          'crps': metrics.SpatialCRPS(**e),
          'ensemble_variance': metrics.SpatialEnsembleVariance(**e)}),
      'probabilistic_spatial_histograms': config.Eval(metrics={
          'rank_histogram': metrics.RankHistogram(**e)}),
  }
  dc = config.Data(
      selection=config.Selection(variables=['geopotential'],
                                 levels=list(LEVELS),
                                 time_slice=slice(None, None)),
      paths=config.Paths(forecast=d['forecast'], obs=d['truth'],
                         output_dir=str(tmp_path)), by_init=True)
  with scope():
    out = evaluation.evaluate_in_memory(dc, eval_configs)
  names = list(regions)
  f, t = d['fv']['geopotential'], _truth_at_valid(d, 'geopotential')
  fd = ('number',) + ODIMS
  p = out['probabilistic']['geopotential']
  assert p.dims == ('metric', 'region', 'lead_time', 'level')
  assert p.attrs.get('ensemble_size', 6) == 6
  want, wd = orc.crps(f, fd, t, ODIMS, 'number', d['lat'], d['lon'],
                      region=oregions['north-america'])
  np.testing.assert_allclose(p.values[0, names.index('north-america')],
                             want.mean(axis=wd.index('time')), rtol=2e-5)
  # crps == skill - spread / 2, every region
  np.testing.assert_allclose(p.values[0], p.values[2] - 0.5 * p.values[1],
                             rtol=1e-6, atol=1e-7)
  x = out['ensemble_forecast_vs_era_experimental_metrics']['geopotential']
  want, wd = orc.energy_score(f, fd, t, ODIMS, 'number', d['lat'], d['lon'],
                              region=oregions['tropics_land'])
  np.testing.assert_allclose(x.values[0, names.index('tropics_land')],
                             want.mean(axis=wd.index('time')), rtol=2e-5)
  s = out['probabilistic_spatial']['geopotential']
  assert s.dims == ('metric', 'lead_time', 'level', 'longitude', 'latitude')
  h = out['probabilistic_spatial_histograms']['geopotential']
  assert h.dims[-1] == 'bins' and h.sizes['bins'] == 7
  np.testing.assert_allclose(h.values.sum(axis=-1), 1.0, rtol=1e-5)


def test_official_deterministic_configs(tmp_path):
  case_official_deterministic_configs(tmp_path, fake_ctx.installed)


def test_official_probabilistic_configs(tmp_path):
  case_official_probabilistic_configs(tmp_path, fake_ctx.installed)
