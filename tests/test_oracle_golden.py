"""Pins oracle/wb2_oracle.py against the known-answer values the reference's
OWN tests hold for the hot path.  Each test cites
the reference test (paths relative to /root/reference/weatherbench2).  CPU only.
"""
import numpy as np
import pytest
from scipy import stats

from oracle import wb2_oracle as orc
import wb2_testdata as td

LATLON = dict()


def _ll(ds):
  return ds['coords']['latitude'], ds['coords']['longitude']


# --- metrics_test.py:63-82 ---------------------------------------------------
def test_get_lat_weights():
  w = orc.get_lat_weights(np.array([-75, -45, -15, 15, 45, 75]))
  assert abs(float(w.mean()) - 1.0) < 1e-12
  expected = 3 * np.array([1 - np.sqrt(3) / 2, (np.sqrt(3) - 1) / 2, 1 / 2,
                           1 / 2, (np.sqrt(3) - 1) / 2, 1 - np.sqrt(3) / 2])
  np.testing.assert_allclose(w, expected, rtol=1e-5)


def test_lat_weights_721_survey_values():
  # SURVEY.md A.1 (validated numerically against the cited lines)
  w = orc.get_lat_weights(np.linspace(-90, 90, 721))
  np.testing.assert_allclose(w.sum(), 721.0, rtol=1e-12)
  np.testing.assert_allclose(w[0], 8.579e-4, rtol=1e-3)
  np.testing.assert_allclose(w[1], 6.863e-3, rtol=1e-3)
  np.testing.assert_allclose(w[360], 1.57298, rtol=1e-5)


def test_lat_weights_not_increasing_raises():
  with pytest.raises(ValueError):
    orc.get_lat_weights(np.array([10.0, 0.0, -10.0]))


# --- metrics_test.py:84-131 --------------------------------------------------
def test_wind_vector_rmse():
  kw = dict(variables_3d=['u_component_of_wind', 'v_component_of_wind'],
            variables_2d=[], time_start='2022-01-01', time_stop='2022-01-02')
  forecast = td.mock_forecast_data(lead_stop='0 day', **kw)
  truth = td.mock_truth_data(**kw)
  lat, lon = _ll(truth)
  fdims, fu = forecast['vars']['u_component_of_wind']
  _, fv = forecast['vars']['v_component_of_wind']
  tdims, tu = truth['vars']['u_component_of_wind']
  _, tv = truth['vars']['v_component_of_wind']

  def lvl(x, dims, vals):
    shape = [1] * x.ndim
    shape[dims.index('level')] = 3
    return x + np.array(vals, dtype=float).reshape(shape)

  fu = lvl(fu, fdims, [0, 3, np.nan])
  fv = lvl(fv, fdims, [0, -4, 1])
  tu = lvl(tu, tdims, [0, -3, np.nan])
  tv = lvl(tv, tdims, [0, 4, 1])
  r, d = orc.wind_vector_mse(fu, fv, fdims, tu, tv, tdims, lat, lon)
  r = np.sqrt(r)
  r, d = orc.time_mean(r, d)
  np.testing.assert_allclose(r.squeeze(), np.array([0, 10, np.nan]))


# --- metrics_test.py:133-152 -------------------------------------------------
@pytest.mark.parametrize('invalid_value', [np.inf, np.nan])
def test_rmse_over_invalid_region(invalid_value):
  lat = np.array([-45.0, 0.0, 45.0])
  lon = np.array([0.0])
  truth = np.array([0.0, invalid_value, 0.0]).reshape(1, 1, 3)
  dims = ('time', 'longitude', 'latitude')
  forecast = truth + 1
  r, d = orc.rmse_sqrt_before_time_avg(forecast, dims, truth, dims, lat, lon)
  r, _ = orc.time_mean(r, d)
  assert np.isnan(r)
  r, d = orc.rmse_sqrt_before_time_avg(
      forecast, dims, truth, dims, lat, lon, region=orc.ExtraTropicalRegion())
  r, _ = orc.time_mean(r, d)
  np.testing.assert_allclose(r, 1.0)


# --- regions_test.py:27-49 ---------------------------------------------------
def test_land_region():
  kw = dict(variables_3d=[], variables_2d=['2m_temperature'],
            time_start='2020-01-01', time_stop='2020-01-03')
  truth = td.mock_truth_data(**kw)
  lat, lon = _ll(truth)
  dims, t = truth['vars']['2m_temperature']
  f = t.copy()
  # forecast wrong over the sea (lsm == 0), perfect over land
  lsm = np.zeros((lat.size, lon.size))
  lsm[2:5, 3:9] = 1.0
  ilat, ilon = dims.index('latitude'), dims.index('longitude')
  sea = np.transpose(1 - lsm) if ilon < ilat else (1 - lsm)
  f = f + sea
  r, d = orc.rmse_sqrt_before_time_avg(
      f, dims, t, dims, lat, lon, region=orc.LandRegion(lsm))
  np.testing.assert_allclose(r, 0.0)
  r, d = orc.rmse_sqrt_before_time_avg(f, dims, t, dims, lat, lon)
  assert (r > 0).all()


# --- metrics_test.py:173-187 -------------------------------------------------
@pytest.mark.parametrize('shape,axis', [((4, 5, 6), 0), ((4, 8, 6), 1),
                                        ((4, 2, 6), 2), ((4, 5, 7), -1),
                                        ((1, 5), 0), ((1, 5), 1)])
def test_rankdata_vs_scipy(shape, axis):
  x = np.random.RandomState(1729 + axis + np.prod(shape)).rand(*shape)
  np.testing.assert_array_equal(
      orc.rankdata(x, axis), stats.rankdata(x, method='ordinal', axis=axis))


# --- metrics_test.py:192-208 -------------------------------------------------
@pytest.mark.parametrize('ensemble_size', [2, 3, 5])
def test_crps_vs_brute_force(ensemble_size):
  truth, forecast = td.get_random_truth_and_forecast(
      ensemble_size=ensemble_size)
  lat, lon = _ll(truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  expected, ed = orc.crps_brute_force(f, fd, t, tdm, 'realization', lat, lon,
                                      skipna=False)
  got, gd = orc.crps(f, fd, t, tdm, 'realization', lat, lon)
  a, b, _ = orc.align(expected['score'], ed, got, gd)
  np.testing.assert_allclose(a, b, rtol=1e-5)


# --- metrics_test.py:210-230 -------------------------------------------------
def test_crps_ensemble_size_1_gives_mae():
  truth, forecast = td.get_random_truth_and_forecast(ensemble_size=1)
  lat, lon = _ll(truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  f0 = f[0]
  ta, fa, d = orc.align(t, tdm, f0, fd[1:])
  expected, ed = orc.spatial_average(np.abs(ta - fa), d, lat, lon)
  skill, sd = orc.crps_skill(f, fd, t, tdm, 'realization', lat, lon)
  a, b, _ = orc.align(skill, sd, expected, ed)
  np.testing.assert_allclose(a, b)
  spread, _ = orc.crps_spread(f, fd, 'realization', lat, lon)
  np.testing.assert_array_equal(spread, 0)
  score, scd = orc.crps(f, fd, t, tdm, 'realization', lat, lon)
  a, b, _ = orc.align(score, scd, expected, ed)
  np.testing.assert_allclose(a, b)


# --- metrics_test.py:232-267 -------------------------------------------------
@pytest.mark.parametrize('skipna', [True, False])
def test_nan_forecasts_result_in_nan_crps(skipna):
  truth, forecast = td.get_random_truth_and_forecast(
      variables=['geopotential', 'temperature'], ensemble_size=7)
  lat, lon = _ll(truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  f = f.copy()
  f[(0,) * f.ndim] = np.nan
  got, gd = orc.crps(f, fd, t, tdm, 'realization', lat, lon, skipna=skipna)
  # xarray puts truth's dims first for abs(truth - forecast): metrics.py:824
  assert gd == ('time', 'level', 'prediction_timedelta')
  sv = got.copy()
  assert np.isnan(sv[0, 0, 0]) == (not skipna)
  sv[0, 0, 0] = 0
  assert np.all(np.isfinite(sv))
  _, f2 = forecast['vars']['temperature']
  _, t2 = truth['vars']['temperature']
  got2, _ = orc.crps(f2, fd, t2, tdm, 'realization', lat, lon, skipna=skipna)
  assert np.all(np.isfinite(got2))
  expected, ed = orc.crps_brute_force(f, fd, t, tdm, 'realization', lat, lon,
                                      skipna=skipna)
  a, b, _ = orc.align(expected['score'], ed, got, gd)
  np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-4)


# --- metrics_test.py:269-281 -------------------------------------------------
def test_crps_repeated_forecasts_are_okay():
  truth, forecast = td.get_random_truth_and_forecast(ensemble_size=7)
  lat, lon = _ll(truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  assert fd.index('realization') == 0
  f = f.copy()
  f[0] = f[1]
  got, gd = orc.crps(f, fd, t, tdm, 'realization', lat, lon)
  expected, ed = orc.crps_brute_force(f, fd, t, tdm, 'realization', lat, lon,
                                      skipna=False)
  a, b, _ = orc.align(expected['score'], ed, got, gd)
  np.testing.assert_allclose(a, b, rtol=1e-5)


# --- metrics_test.py:782-851 -------------------------------------------------
@pytest.mark.parametrize('ensemble_size', [1, 2, 3, 100])
def test_ensemble_mean_rmse_and_stddev(ensemble_size):
  truth, forecast = td.get_random_truth_and_forecast(
      ensemble_size=ensemble_size)
  lat, lon = _ll(truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  rmse, rd = orc.ensemble_mean_rmse_sqrt_before_time_avg(
      f, fd, t, tdm, 'realization', lat, lon)
  std, sd = orc.ensemble_stddev_sqrt_before_time_avg(
      f, fd, 'realization', lat, lon)
  assert set(rd) == {'prediction_timedelta', 'time', 'level'}
  assert set(sd) == {'prediction_timedelta', 'time', 'level'}
  if ensemble_size == 1:
    np.testing.assert_array_equal(std, 0)
    return
  n_indep = rmse.size
  atol = 4 * (1 / np.sqrt(n_indep) + 1 / ensemble_size)
  np.testing.assert_allclose(rmse.mean(), std.mean(), atol=atol)


def test_effect_of_large_bias_on_rmse():
  truth, forecast = td.get_random_truth_and_forecast(ensemble_size=10)
  lat, lon = _ll(truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  rmse, _ = orc.ensemble_mean_rmse_sqrt_before_time_avg(
      f, fd, t + 1000, tdm, 'realization', lat, lon)
  np.testing.assert_allclose(1000, rmse.mean(), rtol=1e-3)


def test_perfect_prediction_zero_rmse():
  truth, _ = td.get_random_truth_and_forecast(ensemble_size=10)
  lat, lon = _ll(truth)
  tdm, t = truth['vars']['geopotential']
  rmse, _ = orc.ensemble_mean_rmse_sqrt_before_time_avg(
      t[None], ('realization',) + tdm, t, tdm, 'realization', lat, lon)
  np.testing.assert_allclose(rmse, 0)


# --- metrics_test.py:854-893 -------------------------------------------------
def test_debiased_ensemble_mean_mse_versus_large_ensemble():
  truth, forecast = td.get_random_truth_and_forecast(
      ensemble_size=1000, spatial_resolution_in_degrees=20)
  lat, lon = _ll(truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  small = f[:2]
  mse_large, _ = orc.ensemble_mean_mse(f, fd, t, tdm, 'realization', lat, lon)
  mse_small, _ = orc.ensemble_mean_mse(small, fd, t, tdm, 'realization', lat,
                                       lon)
  mse_deb, _ = orc.debiased_ensemble_mean_mse(small, fd, t, tdm,
                                              'realization', lat, lon)
  var_large, _ = orc.ensemble_variance(f, fd, 'realization', lat, lon)
  np.testing.assert_allclose((mse_small - mse_large).mean(),
                             var_large.max() / 2, rtol=0.05)
  stderr = np.sqrt(var_large.max() / t.size)
  np.testing.assert_allclose(mse_large.mean(), mse_deb.mean(),
                             atol=4 * stderr)


# --- metrics_test.py:923-984 -------------------------------------------------
@pytest.mark.parametrize('ensemble_size', [1, 2, 3])
def test_energy_score_on_random_dataset(ensemble_size):
  truth, forecast = td.get_random_truth_and_forecast(
      ensemble_size=ensemble_size)
  lat, lon = _ll(truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  args = (f, fd, t, tdm, 'realization', lat, lon)
  score, scd = orc.energy_score(*args)
  spread, spd = orc.energy_score_spread(f, fd, 'realization', lat, lon)
  skill, skd = orc.energy_score_skill(*args)
  for d in (scd, spd, skd):
    assert set(d) == {'prediction_timedelta', 'time', 'level'}
  if ensemble_size == 1:
    np.testing.assert_array_equal(spread, 0)
    np.testing.assert_allclose(score, skill)
    return
  n = score.size
  np.testing.assert_allclose(spread.mean(), skill.mean(),
                             atol=4 * score.std() / np.sqrt(n))
  np.testing.assert_allclose(score, skill - 0.5 * spread)


def test_energy_score_effect_of_bias():
  truth, forecast = td.get_random_truth_and_forecast(ensemble_size=10)
  lat, lon = _ll(truth)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  score, _ = orc.energy_score(f, fd, t + 1000, tdm, 'realization', lat, lon)
  spread, _ = orc.energy_score_spread(f, fd, 'realization', lat, lon)
  np.testing.assert_allclose(1000, score.mean(), rtol=1e-3)
  np.testing.assert_allclose(spread.mean(), np.sqrt(2), rtol=0.05)


# --- regridding_test.py:252-271 ----------------------------------------------
def test_conservative_latitude_weights():
  expected = np.array([
      [1 - np.sqrt(3) / 2, (np.sqrt(3) - 1) / 2, 1 / 2, 0, 0, 0],
      [0, 0, 0, 1 / 2, (np.sqrt(3) - 1) / 2, 1 - np.sqrt(3) / 2]])
  actual = orc.conservative_latitude_weights(
      np.array([-75, -45, -15, 15, 45, 75]), np.array([-45, 45]), True, True)
  np.testing.assert_almost_equal(expected, actual)


# --- regridding_test.py:273-283 ----------------------------------------------
@pytest.mark.parametrize('x,y,expected', [(1, 0, 1), (-1, 0, -1), (5, 0, 5),
                                          (6, 0, -4), (1, 9, 11), (5, 9, 5)])
def test_align_phase_with(x, y, expected):
  assert orc.align_phase_with(x, y, period=10) == expected


# --- regridding_test.py:285-311 ----------------------------------------------
def test_conservative_longitude_weights():
  expected = np.array([[4, 1, 0, 0, 0, 1], [0, 3, 3, 0, 0, 0],
                       [0, 0, 1, 4, 1, 0], [0, 0, 0, 0, 3, 3]]) / 6
  actual = orc.conservative_longitude_weights(
      np.array([0, 60, 120, 180, 240, 300]), np.array([0, 90, 180, 270]),
      True, True)
  np.testing.assert_allclose(expected, actual, atol=1e-5)
  actual = orc.conservative_longitude_weights(
      np.array([90, 180, 270, 360]), np.array([-270, -180, -90, 0]), True,
      True)
  np.testing.assert_allclose(np.eye(4), actual, atol=1e-5)


# --- regridding_test.py:313-330 ----------------------------------------------
def test_conservative_regridding_extrapolation():
  src = orc.Grid(longitudes=np.array([1, 3, 5]), latitudes=np.array([1, 3]),
                 includes_poles=False, periodic=False)
  tgt = orc.Grid(longitudes=np.array([0, 2, 4]), latitudes=np.array([0, 2]),
                 includes_poles=False, periodic=False)
  field = np.array([[1, 1], [2, 2], [3, 3]])
  actual = orc.conservative_regrid(field, src, tgt)
  expected = np.array([[np.nan, np.nan], [np.nan, 1.5], [np.nan, 2.5]])
  np.testing.assert_allclose(actual, expected, atol=1e-6)


# --- regridding_test.py:332-412 ----------------------------------------------
@pytest.mark.parametrize('sp,tp,sper,tper,expect_nans', [
    (True, True, True, True, False), (False, False, True, True, True),
    (False, True, True, True, True), (True, False, True, True, False),
    (True, True, False, False, True)])
def test_conservative_regridder_has_expected_nans(sp, tp, sper, tper,
                                                  expect_nans):
  def lats(poles, n):
    return np.linspace(-90, 90, n) if poles else np.linspace(-80, 80, n)

  def lons(periodic, n):
    return (np.linspace(0, 360, n, endpoint=False) if periodic
            else np.linspace(0, 180, n))

  src = orc.Grid(longitudes=lons(sper, 20), latitudes=lats(sp, 10),
                 includes_poles=sp, periodic=sper)
  tgt = orc.Grid(longitudes=lons(tper, 15), latitudes=lats(tp, 8),
                 includes_poles=tp, periodic=tper)
  actual = orc.conservative_regrid(np.ones(src.shape), src, tgt)
  assert np.isnan(actual).any() == expect_nans
  np.testing.assert_allclose(actual[~np.isnan(actual)], 1.0, atol=1e-6)


# --- regridding_test.py:428-449 ----------------------------------------------
def test_regridding_shape():
  src = orc.Grid(longitudes=np.linspace(0, 360, 128, endpoint=False),
                 latitudes=np.linspace(-90, 90, 65))
  tgt = orc.Grid(longitudes=np.linspace(0, 360, 100, endpoint=False),
                 latitudes=np.linspace(-90, 90, 50))
  assert orc.conservative_regrid(np.zeros(src.shape), src, tgt).shape == (
      100, 50)
  assert orc.conservative_regrid(np.zeros((2,) + src.shape), src,
                                 tgt).shape == (2, 100, 50)


# --- regridding_test.py:465-493 ----------------------------------------------
def test_regridding_nans():
  src = orc.Grid(longitudes=np.linspace(0, 360, 512, endpoint=False),
                 latitudes=np.linspace(-90, 90, 256))
  tgt = orc.Grid(longitudes=np.linspace(0, 360, 360, endpoint=False),
                 latitudes=np.linspace(-90, 90, 181))
  slat = np.deg2rad(src.latitudes)
  slon = np.deg2rad(src.longitudes)
  in_valid = (slat[None, :] ** 2 + (slon[:, None] - np.pi) ** 2
              < (np.pi / 2) ** 2)
  inputs = np.where(in_valid, 1.0, np.nan)
  out = orc.conservative_regrid(inputs, src, tgt)
  out_valid = ~np.isnan(out)
  np.testing.assert_allclose(out_valid.mean(), in_valid.mean(), atol=0.01)
  np.testing.assert_allclose(out[out_valid], 1.0, rtol=1e-6)


def test_regrid_quarter_degree_taps():
  # 0.25 -> 1.5 degree has exactly 7 lon taps [1/12, 1/6 x5, 1/12]
  wlon = orc.conservative_longitude_weights(
      np.arange(1440) * 0.25, np.arange(240) * 1.5, True, True,
      dtype=np.float64)
  assert ((wlon > 0).sum(axis=1) == 7).all()
  row = wlon[0]
  nz = np.nonzero(row)[0]
  assert set(nz) == {1437, 1438, 1439, 0, 1, 2, 3}
  np.testing.assert_allclose(np.sort(row[nz]), [1 / 12] * 2 + [1 / 6] * 5)
  wlat = orc.conservative_latitude_weights(
      np.linspace(-90, 90, 721), np.linspace(-90, 90, 121), True, True,
      dtype=np.float64)
  cnt = (wlat > 0).sum(axis=1)
  assert cnt[0] == 4 and cnt[-1] == 4 and (cnt[1:-1] == 7).all()
  np.testing.assert_allclose(wlat.sum(axis=1), 1.0)


# --- derived_variables_test.py:246-288 ---------------------------------------
def _random_weather(res=30, seed=802701, **kw):
  args = dict(variables_3d=['geopotential'], variables_2d=[],
              time_start='2019-12-01', time_stop='2019-12-02',
              spatial_resolution_in_degrees=res)
  args.update(kw)
  return td.random_like(td.mock_forecast_data(**args), seed=seed + 1)


def test_spectrum_shape_and_coords():
  ds = _random_weather()
  lat, lon = _ll(ds)
  dims, x = ds['vars']['geopotential']
  s, sd, freq, wl = orc.zonal_energy_spectrum(x, dims, lat, lon)
  assert sd[-1] == 'zonal_wavenumber'
  assert s.shape[-1] == lon.size // 2 + 1
  assert freq.shape == (lon.size // 2 + 1, lat.size)
  assert (np.diff(freq[:, 1:-1], axis=0) > 0).all()
  np.testing.assert_array_equal(freq[0], 0)
  np.testing.assert_array_equal(wl, 1 / freq)


# --- derived_variables_test.py:290-321 ---------------------------------------
@pytest.mark.parametrize('latitude', [0, 30, 60])
def test_longitudinal_wave_detected(latitude):
  ds = _random_weather(res=10)
  lat, lon = _ll(ds)
  dims, x = ds['vars']['geopotential']
  i = int(np.argmin(np.abs(lat - latitude)))
  x = np.take(x, [i], axis=dims.index('latitude'))
  lat1 = lat[i:i + 1]
  wavelength_lon = 100
  shape = [1] * x.ndim
  shape[dims.index('longitude')] = lon.size
  x = x + 10 * np.cos(2 * np.pi * lon / wavelength_lon).reshape(shape)
  s, sd, freq, _ = orc.zonal_energy_spectrum(x, dims, lat1, lon)
  wavelength_m = (wavelength_lon / 360) * 2 * np.pi * orc.EARTH_RADIUS_M * (
      np.cos(np.deg2rad(latitude)))
  k_expected = int(np.argmin(np.abs(freq[:, 0] - 1 / wavelength_m)))
  assert (np.argmax(s, axis=-1) == k_expected).all()


# --- derived_variables_test.py:409-435 (Parseval) ----------------------------
@pytest.mark.parametrize('add_constant', [False, True])
def test_parsevals_relation(add_constant):
  res = 5
  ds = _random_weather(res=res)
  lat, lon = _ll(ds)
  dims, x = ds['vars']['geopotential']
  x = 0 * x
  level = ds['coords']['level']

  def bc(v, name):
    shape = [1] * x.ndim
    shape[dims.index(name)] = v.size
    return v.reshape(shape)

  n_signals = 100
  for wl in np.linspace(50, 100, num=n_signals):
    x = x + (np.cos(2 * np.pi * bc(lon, 'longitude') / wl) * np.exp(-wl / 100)
             * np.sin(bc(level, 'level') / 500)
             * np.cos(bc(lat, 'latitude') / 100)) / n_signals
  sel = (lat >= -30) & (lat <= 30)
  x = np.compress(sel, x, axis=dims.index('latitude'))
  lat = lat[sel]
  x = x + (50 if add_constant else 0) * np.abs(x).mean()
  spacing = orc.lon_spacing_m(lat, lon)
  shape = [1] * x.ndim
  shape[dims.index('latitude')] = lat.size
  energy = (spacing.reshape(shape) * x ** 2).sum(axis=dims.index('longitude'))
  s, sd, _, _ = orc.zonal_energy_spectrum(x, dims, lat, lon)
  np.testing.assert_allclose(s.sum(axis=-1), energy, rtol=2e-3)


def test_spectrum_nyquist_doubling_identity():
  # sum_k S_k / C = mean(f^2) + |F_{L/2}|^2 for even L
  rs = np.random.RandomState(0)
  lat = np.array([0.0])
  lon = np.arange(16) * 22.5
  x = rs.normal(size=(1, 16))
  s, _, _, _ = orc.zonal_energy_spectrum(x, ('latitude', 'longitude'), lat,
                                         lon)
  c = orc.circumference(lat)[0]
  fk = np.fft.rfft(x[0], norm='forward')
  np.testing.assert_allclose(s.sum() / c, np.mean(x ** 2) + abs(fk[-1]) ** 2)


# --- label-inclusive slice rule (regions.py:79-95; pandas slice_indexer) -----
def test_slice_region_label_rule_matches_pandas():
  import pandas as pd
  lat = np.linspace(-90, 90, 37)
  lon = np.linspace(0, 360, 72, endpoint=False)
  for coord, slices in [(lat, [slice(-20, 20), slice(None, -20.0),
                               slice(20, None), slice(-22.5, 17.5),
                               slice(None, None)]),
                        (lon, [slice(347.5, None), slice(0, 42.5),
                               slice(240, 290)])]:
    idx = pd.Index(coord)
    for s in slices:
      exp = np.arange(coord.size)[idx.slice_indexer(s.start, s.stop)]
      np.testing.assert_array_equal(orc._label_slice_indices(coord, s), exp)


def test_slice_region_mean_subset():
  rs = np.random.RandomState(3)
  lat = np.linspace(-90, 90, 19)
  lon = np.linspace(0, 360, 36, endpoint=False)
  x = rs.normal(size=(4, 19, 36))
  dims = ('time', 'latitude', 'longitude')
  region = orc.SliceRegion(lat_slice=slice(-20, 20),
                           lon_slice=[slice(335, None), slice(0, 45)])
  r, _ = orc.spatial_average(x, dims, lat, lon, region)
  ilat = np.where((lat >= -20) & (lat <= 20))[0]
  ilon = np.concatenate([np.where(lon >= 335)[0], np.where(lon <= 45)[0]])
  w = orc.get_lat_weights(lat)[ilat]
  sub = x[:, ilat][:, :, ilon]
  exp = (sub * w[None, :, None]).sum(axis=(1, 2)) / (w.sum() * ilon.size)
  np.testing.assert_allclose(r, exp, rtol=1e-12)


# ---- Gaussian / threshold metrics: the reference's known answers -------------
# weatherbench2/metrics_test.py:284-304, 368-532, 987-1030, 1292-1390 (mock data
# are constant fields, so the point-wise score IS the expected average).
def test_gaussian_crps_known_answer():
  got = orc.gaussian_crps_pointwise(np.float32(1.0), np.float32(1.0),
                                    np.float32(1.02))
  np.testing.assert_allclose(got, 0.23385455, rtol=1e-6)


@pytest.mark.parametrize('error,expected_1,expected_2',
                         [(0.02, 0.04421, 0.257883), (1e6, 0.70786, 0.707861)])
def test_gaussian_brier_known_answers(error, expected_1, expected_2):
  f = s = np.float32(1.0 + error)
  t = np.float32(1.0)
  thr = orc.gaussian_quantile_threshold(np.float32(1.0), np.float32(1.0), 0.8)
  np.testing.assert_allclose(orc.gaussian_brier_pointwise(f, s, t, thr),
                             expected_1, rtol=1e-4)
  np.testing.assert_allclose(
      orc.gaussian_brier_pointwise(f, s, t, np.float32(1.0)), expected_2,
      rtol=1e-4)


@pytest.mark.parametrize('error,expected', [(0.02, 0.236055), (1e6, 1.841019)])
def test_gaussian_ignorance_known_answers(error, expected):
  f = s = np.float32(1.0 + error)
  thr = orc.gaussian_quantile_threshold(np.float32(1.0), np.float32(1.0), 0.8)
  np.testing.assert_allclose(
      orc.gaussian_ignorance_pointwise(f, s, np.float32(1.0), thr), expected,
      rtol=1e-4)


@pytest.mark.parametrize('error,expected', [(0.02, 0.295746), (1e6, 0.758203)])
def test_gaussian_rps_known_answers(error, expected):
  f = s = np.float32(1.0 + error)
  got = sum(orc.gaussian_rps_part_pointwise(f, s, np.float32(1.0),
                                            np.float32(q))
            for q in (0.0, 1.0, 2.0))
  np.testing.assert_allclose(got, expected, rtol=1e-4)


@pytest.mark.parametrize('error,ens_delta,expected',
                         [(0.0, 0.1, 0.0), (0.0, 1.0, 0.25), (-10.0, 0.1, 1.0)])
def test_ensemble_brier_known_answers(error, ens_delta, expected):
  x = (1.0 + error + ens_delta * np.arange(-2, 2)).astype(np.float32)
  thr = orc.gaussian_quantile_threshold(np.float32(1.0), np.float32(1.0), 0.2)
  got = orc.ens_brier_pointwise(x, np.float32(1.0), thr, 0, False, False)
  np.testing.assert_allclose(got, expected, rtol=1e-4, atol=1e-12)


def test_ensemble_ignorance_and_rps_known_answers():
  thr = orc.gaussian_quantile_threshold(np.float32(1.0), np.float32(1.0), 0.2)
  x = np.full(4, 1.0, np.float32)
  assert orc.ens_ignorance_pointwise(x, np.float32(1.0), thr, 0, False) == 0
  assert np.isinf(orc.ens_ignorance_pointwise(x - 10, np.float32(1.0), thr, 0,
                                              False))
  for error, expected in ((0.02, 0.0), (-2.0, 2.0)):
    x = np.full(4, 1.0 + error, np.float32)
    got = sum(orc.ens_rps_part_pointwise(x, np.float32(1.5), np.float32(q), 0,
                                         False) for q in (0.0, 1.0, 2.0))
    assert got == expected


def test_debiased_brier_integrates_to_crps():
  """metrics_test.py:1207-1289: the integral over thresholds of the debiased
  Brier score of a 2-member ensemble equals its (fair) CRPS."""
  rs = np.random.RandomState(0)
  x = rs.normal(size=(2, 400))
  t = rs.normal(size=400)
  grid = np.linspace(-6, 6, 4001)
  bs = np.stack([orc.ens_brier_pointwise(x, t, np.full(400, g), 0, True, False)
                 for g in grid])
  integral = np.trapezoid(bs, grid, axis=0).mean()
  skill = np.abs(x - t).mean()
  spread = np.abs(x[0] - x[1]).mean()
  np.testing.assert_allclose(integral, skill - 0.5 * spread, rtol=5e-3)


# ---- nearest / bilinear regridders: regridding_test.py:495-591 ----------------
@pytest.mark.parametrize('periodic,expected', [
    (True, [[0.5], [1.5], [2.5], [1.5]]), (False, [[0.5], [1.5], [2.5], [np.nan]])])
def test_bilinear_regridder_longitude_periodicity(periodic, expected):
  src = orc.Grid(longitudes=np.array([0.0, 90.0, 180.0, 270.0]),
                 latitudes=np.array([0]), includes_poles=True, periodic=periodic)
  tgt = orc.Grid(longitudes=np.array([45.0, 135.0, 225.0, 315.0]),
                 latitudes=np.array([0]), includes_poles=True, periodic=periodic)
  got = orc.bilinear_regrid(np.array([[0.0], [1.0], [2.0], [3.0]]), src, tgt)
  np.testing.assert_allclose(got, expected, atol=1e-6)


@pytest.mark.parametrize('poles,slat,tlat,vals,expected', [
    (True, [-90.0, -30.0, 30.0, 90.0], [-60.0, 0.0, 60.0], [0.0, 1.0, 2.0, 3.0],
     [[0.5, 1.5, 2.5]]),
    (True, [-60.0, 0.0, 60.0], [-90.0, -30.0, 30.0, 90.0], [0.0, 1.0, 2.0],
     [[0.0, 0.5, 1.5, 2.0]]),
    (False, [-60.0, -20.0, 20.0, 60.0], [-70.0, 0.0, 70.0],
     [0.0, 1.0, 2.0, 3.0], [[np.nan, 1.5, np.nan]])])
def test_bilinear_regridder_latitude_poles(poles, slat, tlat, vals, expected):
  src = orc.Grid(longitudes=np.array([0.0]), latitudes=np.array(slat),
                 includes_poles=poles, periodic=True)
  tgt = orc.Grid(longitudes=np.array([0.0]), latitudes=np.array(tlat),
                 includes_poles=poles, periodic=True)
  got = orc.bilinear_regrid(np.array(vals)[np.newaxis, :], src, tgt)
  np.testing.assert_allclose(got, expected, atol=1e-6)


def test_nearest_regridder_exact():
  src = orc.Grid(longitudes=np.array([0, 90, 180, 270]),
                 latitudes=np.array([-30, 0, 30]))
  tgt = orc.Grid(longitudes=np.array([0, 180]), latitudes=np.array([-30, 0, 30]))
  field = np.array([[0, 1, 2], [4, 5, 6], [7, 8, 9], [10, 11, 12]])
  got = orc.nearest_regrid(field, src, tgt)
  np.testing.assert_allclose(got, [[0, 1, 2], [7, 8, 9]], atol=1e-6)


def test_seeps_known_answers():
  """metrics_test.py:1392-1440: perfect forecast -> 0; forecast light while the
  observation is dry -> 0.5 / p1 = 1.25 at p1 = 0.4."""
  t = np.zeros((3, 4), np.float32)
  wet = np.ones((3, 4), np.float32)
  p1 = np.full((3, 4), 0.4, np.float32)
  np.testing.assert_allclose(orc.seeps_pointwise(t, t, wet, wet, p1), 0,
                             atol=1e-4)
  np.testing.assert_allclose(orc.seeps_pointwise(t + 0.5, t, wet, wet, p1),
                             1.25, atol=1e-4)


# ---- interpolate_spectral_frequencies (derived_variables_test.py:432-528) -----
def _multispectral(nlat_half=30, nlon=360, seed=0):
  rs = np.random.RandomState(seed)
  lat = np.arange(-nlat_half, nlat_half + 1, 5.0)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  x = rs.standard_normal((2, lat.size, nlon))
  for k in (3, 7, 20):
    x += np.cos(2 * np.pi * k * lon / 360)[None, None, :] * (1 + k)
  return x.astype(np.float32), lat, lon


def test_interpolate_spectral_frequencies_reference_properties():
  """The two properties the reference's tests pin: with the default
  frequencies the latitude = 0 row (the narrowest range) is unchanged
  (derived_variables_test.py:454-462); with the frequencies of latitude 5 the
  latitude-5 row is unchanged (:503-511).  (The "nearby rows barely change"
  checks of the reference hold for its smooth test spectrum only.)"""
  x, lat, lon = _multispectral()
  dims = ('time', 'latitude', 'longitude')
  spec, sd, freq, _ = orc.zonal_energy_spectrum(x, dims, lat, lon)
  assert sd == ('time', 'latitude', 'zonal_wavenumber')
  out, fr = orc.interpolate_spectral_frequencies(spec, freq)
  i0 = int(np.where(lat == 0)[0][0])
  np.testing.assert_allclose(fr, freq[:, i0])
  np.testing.assert_allclose(out[:, i0], spec[:, i0], rtol=1e-9)
  i5 = int(np.where(lat == 5)[0][0])
  out, fr = orc.interpolate_spectral_frequencies(spec, freq, freq[3:8, i5])
  np.testing.assert_allclose(out[:, i5], spec[:, i5, 3:8], rtol=1e-9)
  assert out.shape == (2, lat.size, 5)
