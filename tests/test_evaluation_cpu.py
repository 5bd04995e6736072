"""CPU tests of the Python operator layer END TO END -- evaluate_in_memory ->
_metric_and_region_loop -> metrics.batch -> _spatial.run_* -> the C-ABI call --
with `tests/fake_ctx.FakeContext` (test infrastructure, NumPy) interpreting the
raw offset tables / weight factors the operators hand to the library.  Covers
BASELINE.json configs[0] (RMSE, synthetic 64x64, evaluate_in_memory) as pure
plumbing, without a GPU; the same cases run against the CUDA kernels in
tests/test_evaluation_gpu.py."""
import numpy as np
import pytest

from oracle import wb2_oracle as orc
import fake_ctx
from test_evaluation_gpu import _mock


def case_config0_rmse_64x64_plumbing(tmp_path, scope, crps_rtol=1e-5,
                                     det_rtol=2e-6):
  from weatherbench2_b200 import config, evaluation, metrics
  forecast, truth, fv, tv, lat, lon, fdims, _ = _mock(
      levels=(500,), variables=('geopotential',))
  data_config = config.Data(
      selection=config.Selection(variables=['geopotential'],
                                 time_slice=slice(None, None)),
      paths=config.Paths(forecast=forecast, obs=truth,
                         output_dir=str(tmp_path)), by_init=True)
  eval_configs = {'deterministic': config.Eval(
      metrics={'rmse': metrics.RMSESqrtBeforeTimeAvg(), 'mse': metrics.MSE()})}
  with scope() as fake:
    out = evaluation.evaluate_in_memory(data_config, eval_configs)
  # ONE pass over the chunk serves both metrics (the reference makes two)
  if fake is not None:
    assert [c[0] for c in fake.calls] == ['det_metrics']
  res = out['deterministic']['geopotential']
  assert res.dims == ('metric', 'lead_time', 'level')
  f, t = fv['geopotential'], tv['geopotential']
  nlead, ntime = f.shape[:2]
  tg = np.stack([np.stack([t[i + l] for i in range(ntime)])
                 for l in range(nlead)])
  want, wd = orc.rmse_sqrt_before_time_avg(f, fdims, tg, fdims, lat, lon)
  np.testing.assert_allclose(res.values[0], want.mean(axis=wd.index('time')),
                             rtol=det_rtol)
  want, wd = orc.mse(f, fdims, tg, fdims, lat, lon)
  np.testing.assert_allclose(res.values[1], want.mean(axis=wd.index('time')),
                             rtol=det_rtol)


# ------------------------------------------------------------------------------
# Baseline forecasts (climatology / probabilistic climatology / persistence),
# analysis-as-truth and the config.Selection: weatherbench2/evaluation.py
# :139-365, 441-483
# ------------------------------------------------------------------------------
NS = 'datetime64[ns]'
H = np.timedelta64(1, 'h').astype('timedelta64[ns]')


def _grid(nlat=9, nlon=12):
  return np.linspace(-90, 90, nlat), np.linspace(0, 360, nlon, endpoint=False)


def _truth(start, stop, step_h, levels=(500, 850), seed=0, nlat=9, nlon=12,
           variables=('geopotential',)):
  from weatherbench2_b200 import xarray_lite as xl
  lat, lon = _grid(nlat, nlon)
  times = np.arange(np.datetime64(start, 'ns'), np.datetime64(stop, 'ns'),
                    step_h * H)
  rs = np.random.RandomState(seed)
  dims = ('time', 'level', 'longitude', 'latitude')
  arrays = {v: rs.normal(size=(times.size, len(levels), nlon, nlat)).astype(
      np.float32) for v in variables}
  ds = xl.Dataset({v: (dims, a) for v, a in arrays.items()},
                  dict(time=times, level=np.array(levels), latitude=lat,
                       longitude=lon))
  return ds, arrays, times, dims


def _forecast(start, stop, step_h, lead_h, by_init, levels=(500, 850), seed=1,
              nlat=9, nlon=12, variables=('geopotential',)):
  """WB2 on-disk conventions: dims time (= init or valid time) and
  prediction_timedelta."""
  from weatherbench2_b200 import xarray_lite as xl
  lat, lon = _grid(nlat, nlon)
  times = np.arange(np.datetime64(start, 'ns'), np.datetime64(stop, 'ns'),
                    step_h * H)
  lead = np.asarray(lead_h) * H
  rs = np.random.RandomState(seed)
  dims = ('time', 'prediction_timedelta', 'level', 'longitude', 'latitude')
  arrays = {v: rs.normal(size=(times.size, lead.size, len(levels), nlon, nlat)
                         ).astype(np.float32) for v in variables}
  ds = xl.Dataset({v: (dims, a) for v, a in arrays.items()},
                  dict(time=times, prediction_timedelta=lead,
                       level=np.array(levels), latitude=lat, longitude=lon))
  del by_init
  return ds, arrays, times, lead


def _climatology(hours=(0, 6, 12, 18), levels=(500, 850), seed=2, nlat=9,
                 nlon=12, variables=('geopotential',)):
  from weatherbench2_b200 import xarray_lite as xl
  lat, lon = _grid(nlat, nlon)
  rs = np.random.RandomState(seed)
  if hours is None:
    dims = ('dayofyear', 'level', 'longitude', 'latitude')
    shape = (366, len(levels), nlon, nlat)
    coords = dict(dayofyear=np.arange(1, 367))
  else:
    dims = ('hour', 'dayofyear', 'level', 'longitude', 'latitude')
    shape = (len(hours), 366, len(levels), nlon, nlat)
    coords = dict(hour=np.array(hours), dayofyear=np.arange(1, 367))
  arrays = {v: rs.normal(size=shape).astype(np.float32) for v in variables}
  ds = xl.Dataset({v: (dims, a) for v, a in arrays.items()},
                  dict(coords, level=np.array(levels), latitude=lat,
                       longitude=lon))
  return ds, arrays, dims


def _data_config(forecast, truth, tmp_path, by_init, climatology=None, **sel):
  from weatherbench2_b200 import config
  sel.setdefault('variables', ['geopotential'])
  sel.setdefault('time_slice', slice(None, None))
  return config.Data(
      selection=config.Selection(**sel),
      paths=config.Paths(forecast=forecast, obs=truth, output_dir=str(tmp_path),
                         climatology=climatology), by_init=by_init)


@pytest.mark.parametrize('hours', [(0, 6, 12, 18), None])
def test_climatology_forecast_is_a_pointwise_dayofyear_hour_lookup(hours):
  from weatherbench2_b200 import evaluation
  # 12-hourly inits over the 2020 -> 2021 boundary: day 366 and day 1
  fc, _, _, _ = _forecast('2020-12-29', '2021-01-02', 12, [0, 6, 18], True)
  fc = evaluation.apply_time_conventions(fc, by_init=True)
  clim, carr, cdims = _climatology(hours)
  out = evaluation.climatology_forecast(clim, fc, 'valid_time')
  da = out['geopotential']
  assert da.dims == ('init_time', 'lead_time', 'level', 'longitude',
                     'latitude')
  vt = fc['valid_time'].values
  want, _ = orc.climatology_like_forecast(
      carr['geopotential'], cdims, np.arange(1, 367),
      None if hours is None else np.array(hours), vt)
  np.testing.assert_array_equal(da.values, want)
  np.testing.assert_array_equal(out['valid_time'].values, vt)
  assert 366 in da.coords['dayofyear'].values  # 2020-12-31
  # nothing was copied: the view still addresses the climatology array
  src, maps = da.lazy_source
  assert np.shares_memory(src.values, carr['geopotential'])
  assert set(maps) == ({'dayofyear', 'hour'} if hours else {'dayofyear'})


def test_climatology_forecast_mean_suffix_and_level_subset():
  from weatherbench2_b200 import evaluation
  fc, _, _, _ = _forecast('2020-03-01', '2020-03-03', 24, [0, 24], True,
                          levels=(850,))
  fc = evaluation.apply_time_conventions(fc, by_init=True)
  clim, carr, cdims = _climatology(None, levels=(500, 700, 850))
  clim = clim.rename({'geopotential': 'geopotential_mean'})
  out = evaluation.climatology_forecast(clim, fc, 'valid_time')
  want, _ = orc.climatology_like_forecast(
      carr['geopotential'][:, 2:3], cdims, np.arange(1, 367), None,
      fc['valid_time'].values)
  np.testing.assert_array_equal(out['geopotential'].values, want)
  with pytest.raises(KeyError):
    evaluation.climatology_forecast(
        clim.rename({'geopotential_mean': 'other'}), fc, 'valid_time')


def test_persistence_forecast_by_init_and_by_valid():
  from weatherbench2_b200 import evaluation
  truth, tarr, ttimes, _ = _truth('2020-01-01', '2020-01-12', 6)
  # by-init: truth at the init time, repeated along lead_time
  fc, _, itimes, lead = _forecast('2020-01-02', '2020-01-06', 12, [0, 6, 24],
                                  True)
  fc = evaluation.apply_time_conventions(fc, by_init=True)
  out = evaluation.create_persistence_forecast(fc, truth)
  da = out['geopotential']
  assert da.dims == ('init_time', 'lead_time', 'level', 'longitude',
                     'latitude')
  want = orc.persistence_like_forecast_by_init(tarr['geopotential'], ttimes,
                                               itimes, lead.size)
  np.testing.assert_array_equal(da.values, want)
  np.testing.assert_array_equal(out['valid_time'].values,
                                fc['valid_time'].values)
  # by-valid (evaluation.py:165-193): the first max(lead) of times is dropped
  fc, _, vtimes, lead = _forecast('2020-01-02', '2020-01-06', 12, [0, 6, 24],
                                  False)
  fc = evaluation.apply_time_conventions(fc, by_init=False)
  out = evaluation.create_persistence_forecast(fc, truth)
  kept, want = orc.persistence_like_forecast_by_valid(
      tarr['geopotential'], ttimes, vtimes, lead)
  assert kept.size == vtimes.size - 2
  da = out['geopotential']
  assert da.dims == ('time', 'lead_time', 'level', 'longitude', 'latitude')
  np.testing.assert_array_equal(da.coords['time'].values, kept)
  np.testing.assert_array_equal(da.values, want)
  # an init time the observations do not hold is an error, not a silent hole
  short = truth.isel(time=slice(8, None))
  with pytest.raises(KeyError):
    evaluation.create_persistence_forecast(fc, short)


def test_probabilistic_climatology_stacks_years_as_members():
  from weatherbench2_b200 import evaluation
  truth, tarr, ttimes, _ = _truth('2019-12-20', '2021-01-10', 12, levels=(500,),
                                  nlat=5, nlon=6)
  pc = evaluation.make_probabilistic_climatology(truth, 2019, 2021, 12)
  da = pc['geopotential']
  hours, days, want = orc.probabilistic_climatology(tarr['geopotential'],
                                                    ttimes, 2019, 2021, 12)
  assert da.dims == ('hour', 'number', 'dayofyear', 'level', 'longitude',
                     'latitude')
  np.testing.assert_array_equal(da.coords['hour'].values, hours)
  np.testing.assert_array_equal(da.coords['dayofyear'].values, days)
  np.testing.assert_array_equal(da.coords['number'].values, [0, 1, 2])
  np.testing.assert_array_equal(da.values, want)
  assert days[-1] == 366
  i366 = list(days).index(366)
  assert np.isnan(da.values[:, 0, i366]).all()       # 2019: no day 366
  assert not np.isnan(da.values[:, 1, i366]).any()   # 2020: leap year
  assert da.values.dtype == np.float32


def test_label_joins_compose_with_lazy_gathers():
  """A by-valid persistence forecast covers fewer times than the truth: the
  inner join of the metric arithmetic restricts the forecast's position table
  instead of materialising the gathered forecast."""
  from weatherbench2_b200 import evaluation, xarray_lite as xl
  truth, tarr, ttimes, _ = _truth('2020-01-01', '2020-01-12', 6)
  fc, _, vtimes, lead = _forecast('2020-01-02', '2020-01-06', 12, [0, 12, 24],
                                  False)
  fc = evaluation.apply_time_conventions(fc, by_init=False)
  pf = evaluation.create_persistence_forecast(fc, truth)['geopotential']
  a, b = xl.align_inner(pf, truth['geopotential'])
  assert hasattr(a, 'lazy_source') and a._materialised is None  # pylint: disable=protected-access
  assert pf._materialised is None  # pylint: disable=protected-access
  np.testing.assert_array_equal(a.coords['time'].values,
                                b.coords['time'].values)
  kept, want = orc.persistence_like_forecast_by_valid(
      tarr['geopotential'], ttimes, vtimes, lead)
  np.testing.assert_array_equal(a.coords['time'].values, kept)
  np.testing.assert_array_equal(a.values, want)
  src, _ = a.lazy_source
  assert np.shares_memory(src.values, tarr['geopotential'])


def _rmse_time_mean(f, t, dims, lat, lon, avg):
  want, wd = orc.rmse_sqrt_before_time_avg(f, dims, t, dims, lat, lon)
  return want.mean(axis=wd.index(avg)), tuple(d for d in wd if d != avg)


def case_evaluate_climatology_and_persistence_in_memory(tmp_path, scope, crps_rtol=1e-5,
                                                        det_rtol=2e-6):
  """evaluate_in_memory with the baseline-forecast switches of config.Eval
  (evaluation.py:450-472) against the oracle on materialised arrays."""
  from weatherbench2_b200 import config, evaluation, metrics
  lat, lon = _grid()
  truth, tarr, ttimes, _ = _truth('2020-12-20', '2021-01-12', 6)
  fc, _, itimes, lead = _forecast('2020-12-28', '2021-01-03', 12, [0, 6, 30],
                                  True)
  clim, carr, cdims = _climatology()
  dc = _data_config(fc, truth, tmp_path, True, climatology=clim)
  mets = {'rmse': metrics.RMSESqrtBeforeTimeAvg()}
  vt = itimes[:, None] + lead[None, :]
  tpos = {t: i for i, t in enumerate(ttimes)}
  tg = np.stack([np.stack([tarr['geopotential'][tpos[v]] for v in row])
                 for row in vt])
  dims = ('time', 'lead_time', 'level', 'longitude', 'latitude')
  with scope():
    out = evaluation.evaluate_in_memory(dc, {
        'clim': config.Eval(metrics=mets, evaluate_climatology=True),
        'pers': config.Eval(metrics=mets, evaluate_persistence=True)})
  cf, _ = orc.climatology_like_forecast(carr['geopotential'], cdims,
                                        np.arange(1, 367),
                                        np.array([0, 6, 12, 18]), vt)
  want, wd = _rmse_time_mean(cf, tg, dims, lat, lon, 'time')
  res = out['clim']['geopotential']
  assert res.dims == ('metric',) + wd
  np.testing.assert_allclose(res.values[0], want, rtol=det_rtol)
  pf = orc.persistence_like_forecast_by_init(tarr['geopotential'], ttimes,
                                             itimes, lead.size)
  want, wd = _rmse_time_mean(pf, tg, dims, lat, lon, 'time')
  res = out['pers']['geopotential']
  assert res.dims == ('metric',) + wd
  np.testing.assert_allclose(res.values[0], want, rtol=det_rtol)
  assert res.values[0][0].max() == 0  # lead 0: persistence IS the truth


def case_evaluate_persistence_by_valid_in_memory(tmp_path, scope, crps_rtol=1e-5,
                                                 det_rtol=2e-6):
  from weatherbench2_b200 import config, evaluation, metrics
  lat, lon = _grid()
  truth, tarr, ttimes, _ = _truth('2020-01-01', '2020-01-12', 6)
  fc, _, vtimes, lead = _forecast('2020-01-02', '2020-01-08', 12, [0, 12, 24],
                                  False)
  dc = _data_config(fc, truth, tmp_path, False)
  with scope():
    out = evaluation.evaluate_in_memory(dc, {'pers': config.Eval(
        metrics={'rmse': metrics.RMSESqrtBeforeTimeAvg()},
        evaluate_persistence=True)})
  kept, pf = orc.persistence_like_forecast_by_valid(tarr['geopotential'],
                                                    ttimes, vtimes, lead)
  tpos = {t: i for i, t in enumerate(ttimes)}
  tk = np.stack([tarr['geopotential'][tpos[v]] for v in kept])
  tk = np.repeat(tk[:, None], lead.size, axis=1)
  dims = ('time', 'lead_time', 'level', 'longitude', 'latitude')
  want, wd = _rmse_time_mean(pf, tk, dims, lat, lon, 'time')
  res = out['pers']['geopotential']
  assert res.dims == ('metric',) + wd
  np.testing.assert_allclose(res.values[0], want, rtol=det_rtol)


def case_evaluate_probabilistic_climatology_crps_in_memory(tmp_path, scope, crps_rtol=1e-5,
                                                           det_rtol=2e-6):
  from weatherbench2_b200 import config, evaluation, metrics
  lat, lon = _grid(5, 6)
  truth, tarr, ttimes, _ = _truth('2018-01-01', '2021-01-10', 12, levels=(500,),
                                  nlat=5, nlon=6)
  # evaluate 2020 against the 2018 / 2019 climatological "ensemble"
  fc, _, itimes, lead = _forecast('2020-02-27', '2020-03-02', 24, [0, 12], True,
                                  levels=(500,), nlat=5, nlon=6)
  dc = _data_config(fc, truth, tmp_path, True)
  with scope() as fake:
    out = evaluation.evaluate_in_memory(dc, {'pc': config.Eval(
        metrics={'crps': metrics.CRPS(ensemble_dim='number')},
        evaluate_probabilistic_climatology=True,
        probabilistic_climatology_start_year=2018,
        probabilistic_climatology_end_year=2019,
        probabilistic_climatology_hour_interval=12)})
  if fake is not None:
    assert fake.calls and fake.calls[0][0] == 'ens_metrics'
  hours, days, pc = orc.probabilistic_climatology(tarr['geopotential'], ttimes,
                                                  2018, 2019, 12)
  vt = itimes[:, None] + lead[None, :]
  cf, cd = orc.climatology_like_forecast(
      np.moveaxis(pc, 0, 1), ('number', 'hour', 'dayofyear', 'level',
                              'longitude', 'latitude'), days, hours, vt)
  assert cd[:3] == ('vt0', 'vt1', 'number')
  # 2020-02-29 is day 60: both non-leap members still have a day 60
  assert not np.isnan(cf).any()
  tpos = {t: i for i, t in enumerate(ttimes)}
  tg = np.stack([np.stack([tarr['geopotential'][tpos[v]] for v in row])
                 for row in vt])
  fd = ('time', 'lead_time', 'number', 'level', 'longitude', 'latitude')
  td = ('time', 'lead_time', 'level', 'longitude', 'latitude')
  want, wd = orc.crps(cf, fd, tg, td, 'number', lat, lon)
  want = want.mean(axis=wd.index('time'))
  res = out['pc']['geopotential']
  np.testing.assert_allclose(res.values[0], want, rtol=crps_rtol)


def case_against_analysis_by_valid_and_by_init(tmp_path, scope, crps_rtol=1e-5,
                                               det_rtol=2e-6):
  from weatherbench2_b200 import config, evaluation, metrics
  lat, lon = _grid()
  truth, _, _, _ = _truth('2020-01-01', '2020-01-12', 6)
  mets = {'rmse': metrics.RMSESqrtBeforeTimeAvg()}
  dims = ('time', 'lead_time', 'level', 'longitude', 'latitude')
  # by-valid: truth = the forecast's own lead-0 fields at the same valid time
  fc, farr, _, lead = _forecast('2020-01-02', '2020-01-05', 12, [0, 12, 24],
                                False)
  f = farr['geopotential']
  with scope():
    out = evaluation.evaluate_in_memory(
        _data_config(fc, truth, tmp_path, False),
        {'an': config.Eval(metrics=mets, against_analysis=True)})
  t = np.repeat(f[:, :1], lead.size, axis=1)
  want, wd = _rmse_time_mean(f, t, dims, lat, lon, 'time')
  np.testing.assert_allclose(out['an']['geopotential'].values[0], want,
                             rtol=det_rtol)
  # by-init (evaluation.py:259-293): inits every 12 h, leads every 6 h ->
  # every second lead has an analysis; the last inits lack one -> error unless
  # the evaluated init times stop early enough
  fc, farr, itimes, lead = _forecast('2020-01-02', '2020-01-06', 12,
                                     [0, 6, 12, 18, 24], True)
  f = farr['geopotential']
  with scope():
    with pytest.raises(AssertionError, match='Analysis does not extend'):
      evaluation.evaluate_in_memory(
          _data_config(fc, truth, tmp_path, True),
          {'an': config.Eval(metrics=mets, against_analysis=True)})
    out = evaluation.evaluate_in_memory(
        _data_config(fc, truth, tmp_path, True,
                     time_slice=slice('2020-01-02', '2020-01-04')),
        {'an': config.Eval(metrics=mets, against_analysis=True)})
  n = int((itimes < np.datetime64('2020-01-05')).sum())
  assert n == 6
  fs = f[:n, ::2]                                   # leads 0, 12, 24 h
  ts = np.stack([np.stack([f[i + k, 0] for k in range(3)]) for i in range(n)])
  want, wd = _rmse_time_mean(fs, ts, dims, lat, lon, 'time')
  res = out['an']['geopotential']
  assert res.sizes['lead_time'] == 3
  np.testing.assert_allclose(res.values[0], want, rtol=det_rtol)


def case_selection_box_levels_suffixes_and_step_thinning(tmp_path, scope, crps_rtol=1e-5,
                                                         det_rtol=2e-6):
  from weatherbench2_b200 import config, evaluation, metrics
  from weatherbench2_b200 import xarray_lite as xl
  lat, lon = _grid(13, 24)
  truth, tarr, ttimes, _ = _truth('2020-01-01', '2020-01-08', 6, nlat=13,
                                  nlon=24, levels=(500, 700, 850))
  # forecast stored with pressure-level suffixes, by valid time, every 12 h
  fc, farr, vtimes, lead = _forecast('2020-01-02', '2020-01-06', 12, [0, 12],
                                     False, nlat=13, nlon=24,
                                     levels=(500, 700, 850))
  f = farr['geopotential']
  sdims = ('time', 'prediction_timedelta', 'longitude', 'latitude')
  flat = xl.Dataset(
      {f'geopotential_{lev}': (sdims, f[:, :, i])
       for i, lev in [(2, 850), (0, 500), (1, 700)]},
      dict(time=vtimes, prediction_timedelta=lead, latitude=lat, longitude=lon))
  dc = _data_config(flat, truth, tmp_path, False, levels=[500, 850],
                    lat_slice=slice(-45, 60), lon_slice=slice(30, 200),
                    time_slice=slice('2020-01-02', '2020-01-04'))
  dc.pressure_level_suffixes = True
  with scope():
    out = evaluation.evaluate_in_memory(dc, {'det': config.Eval(
        metrics={'mse': metrics.MSE()})})
  # oracle: same box / levels / times; truth thinned 6 h -> 12 h by the
  # by-valid step rule, then joined on the forecast's times by label
  li = (lat >= -45) & (lat <= 60)
  lo = (lon >= 30) & (lon <= 200)
  tsel = vtimes < np.datetime64('2020-01-05')
  fs = f[tsel][:, :, [0, 2]][:, :, :, lo][..., li]
  tpos = {t: i for i, t in enumerate(ttimes)}
  ts = np.stack([tarr['geopotential'][tpos[v]] for v in vtimes[tsel]])
  ts = ts[:, [0, 2]][:, :, lo][..., li]
  ts = np.repeat(ts[:, None], lead.size, axis=1)
  dims = ('time', 'lead_time', 'level', 'longitude', 'latitude')
  want, wd = orc.mse(fs, dims, ts, dims, lat[li], lon[lo])
  res = out['det']['geopotential']
  assert res.sizes == {'metric': 1, 'lead_time': 2, 'level': 2}
  np.testing.assert_allclose(
      res.transpose('metric', *[d for d in wd if d != 'time']).values[0],
      want.mean(axis=wd.index('time')), rtol=det_rtol)
  # time steps that are not multiples of each other are refused
  odd = truth.isel(time=np.array([0, 1, 2, 4, 8]))
  with scope(), pytest.raises(ValueError, match='unique'):
    evaluation.evaluate_in_memory(
        _data_config(fc, odd, tmp_path, False),
        {'det': config.Eval(metrics={'mse': metrics.MSE()})})
  # a variable that neither dataset holds is a KeyError, as with xarray
  with scope(), pytest.raises(KeyError):
    evaluation.evaluate_in_memory(
        _data_config(fc, truth, tmp_path, False, variables=['nope']),
        {'det': config.Eval(metrics={'mse': metrics.MSE()})})


def consistency_setup(tmp_path, by_init=True):
  from weatherbench2_b200 import config, metrics, regions as R
  truth, _, _, _ = _truth('2019-12-20', '2020-02-10', 6)
  # (by-valid persistence needs every valid time - lead in the time-sliced,
  # step-thinned truth: evaluation.py:165-193)
  fc, _, _, _ = _forecast('2020-01-01', '2020-01-08', 12,
                             [0, 6, 30] if by_init else [0, 12, 36], by_init)
  clim, _, _ = _climatology()
  regions = {'global': R.SliceRegion(),
             'tropics': R.SliceRegion(lat_slice=slice(-20, 20)),
             'extra-tropics': R.ExtraTropicalRegion()}
  eval_configs = {
      'forecast_vs_era': config.Eval(metrics={
          'rmse': metrics.RMSESqrtBeforeTimeAvg(),
          'acc': metrics.ACC(climatology=clim)}),
      'forecast_vs_era_by_region': config.Eval(
          metrics={'rmse': metrics.RMSESqrtBeforeTimeAvg()}, regions=regions),
      'forecast_vs_era_temporal': config.Eval(
          metrics={'rmse': metrics.RMSESqrtBeforeTimeAvg()},
          temporal_mean=False),
      'climatology_vs_era': config.Eval(
          metrics={'mse': metrics.MSE()}, evaluate_climatology=True),
      'persistence_vs_era': config.Eval(
          metrics={'mse': metrics.MSE()}, evaluate_persistence=True),
  }
  dc = _data_config(fc, truth, tmp_path, by_init, climatology=clim,
                       time_slice=slice('2020-01-02', '2020-01-06'))
  return dc, eval_configs


def case_in_memory_and_chunked_consistency(tmp_path, scope, crps_rtol=1e-5,
                                           det_rtol=2e-6):
  """evaluate_with_beam (chunks of 2 init times through the feeder, slab cache
  and [sum, count] accumulator; one process) == evaluate_in_memory, the
  reference's own consistency test (evaluation_test.py:30-128).  The
  two-rank version runs under gloo in tests/test_distributed_cpu.py."""
  del crps_rtol, det_rtol
  from weatherbench2_b200 import evaluation
  for by_init in (True, False):
    chunk_dim = 'init_time' if by_init else 'time'
    dc, eval_configs = consistency_setup(tmp_path / f'mem{by_init}', by_init)
    with scope():
      mem = evaluation.evaluate_in_memory(dc, eval_configs)
    dc, eval_configs = consistency_setup(tmp_path / f'chunked{by_init}',
                                         by_init)
    with scope():
      chunked = evaluation.evaluate_with_beam(
          dc, eval_configs, input_chunks={chunk_dim: 2}, runner='DirectRunner')
    for name in eval_configs:
      got = chunked[name]['geopotential']
      want = mem[name]['geopotential']
      assert got.dims == want.dims and got.shape == want.shape, name
      np.testing.assert_allclose(got.values, want.values, rtol=1e-12,
                                 atol=1e-15, err_msg=name)


CASES = [case_config0_rmse_64x64_plumbing,
         case_evaluate_climatology_and_persistence_in_memory,
         case_evaluate_persistence_by_valid_in_memory,
         case_evaluate_probabilistic_climatology_crps_in_memory,
         case_against_analysis_by_valid_and_by_init,
         case_selection_box_levels_suffixes_and_step_thinning,
         case_in_memory_and_chunked_consistency]


@pytest.mark.parametrize('case', CASES, ids=lambda c: c.__name__[5:])
def test_in_memory_evaluation_host_logic(case, tmp_path):
  """The cases above with the NumPy stand-in context (no GPU)."""
  case(tmp_path, fake_ctx.installed)
