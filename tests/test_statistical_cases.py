"""The statistical property tests of the reference's metrics_test.py that no
known-answer vector covers, at operator level: GaussianCRPS against the CRPS of
a large Gaussian ensemble (metrics_test.py:306-343) and the shape of rank
histograms of well / badly calibrated ensembles (metrics_test.py:546-610).
Here on the NumPy stand-in context; tests/test_zz_evaluation_cases_gpu.py runs
the same cases on the CUDA kernels."""
import numpy as np
import pytest

import fake_ctx
import wb2_testdata as td


def _ds(data):
  from weatherbench2_b200 import xarray_lite as xl
  return xl.Dataset({k: (d, v) for k, (d, v) in data['vars'].items()},
                    data['coords'])


def case_gaussian_crps_converges(scope, exact_rtol=1e-5):
  """The closed-form CRPS of N(mu, sigma) equals the ensemble CRPS of many
  draws from it.  1000 members (the reference draws 5000; K2 holds up to 1551)
  -> 3 % instead of the reference's 2 %."""
  from weatherbench2_b200 import metrics
  kw = dict(variables_3d=[], time_start='2022-01-01', time_stop='2022-01-02',
            lead_stop='1 day')
  name = '2m_temperature'
  gauss = td.mock_forecast_data(variables_2d=[name, name + '_std'], **kw)
  ens = td.mock_forecast_data(variables_2d=[name], ensemble_size=1000, **kw)
  truth = td.mock_truth_data(variables_3d=[], variables_2d=[name],
                             time_start='2022-01-01', time_stop='2022-01-20')
  d, v = gauss['vars'][name]
  gauss['vars'][name] = (d, v + np.float32(0.1))
  d, v = gauss['vars'][name + '_std']
  gauss['vars'][name + '_std'] = (d, v + np.float32(1.0))
  d, v = ens['vars'][name]
  rs = np.random.RandomState(0)
  ens['vars'][name] = (d, (v + rs.standard_normal(v.shape) + 0.1).astype(
      np.float32))
  # mock data are by-valid: the forecast's `time` is joined with the truth's
  with scope():
    a = metrics.GaussianCRPS().compute(_ds(gauss), _ds(truth))[name].values
    b = metrics.CRPS().compute(_ds(ens), _ds(truth))[name].values
  assert a.shape == b.shape == (2,)
  np.testing.assert_allclose(a, b, rtol=3e-2)
  # N(0.1, 1) against truth 0: sigma (z (2 Phi(z) - 1) + 2 phi(z) - 1/sqrt(pi))
  from scipy import stats
  z = -0.1
  exact = z * (2 * stats.norm.cdf(z) - 1) + 2 * stats.norm.pdf(z) - np.pi**-0.5
  np.testing.assert_allclose(a, exact, rtol=exact_rtol)


def _increasing(x):
  assert (np.diff(x) > 0).all(), x


def _decreasing(x):
  assert (np.diff(x) < 0).all(), x


def case_rank_histogram_calibration(scope, ensemble_size, num_bins=None):
  from weatherbench2_b200 import metrics
  num_bins = ensemble_size + 1 if num_bins is None else num_bins
  truth, forecast = td.get_random_truth_and_forecast(
      ensemble_size=ensemble_size, time_start='2019-12-01',
      time_stop='2019-12-10', levels=(0, 1, 2, 3, 4))
  fd, f = forecast['vars']['geopotential']
  f = f.astype(np.float32)
  lev = [slice(None)] * f.ndim

  def at(level):
    idx = list(lev)
    idx[fd.index('level')] = level
    return tuple(idx)

  f[at(1)] *= 0.1   # under-dispersed
  f[at(2)] *= 10    # over-dispersed
  f[at(3)] -= 1     # skewed left / right
  f[at(4)] += 1
  forecast['vars']['geopotential'] = (fd, f)
  td_, t = truth['vars']['geopotential']
  truth['vars']['geopotential'] = (td_, t.astype(np.float32))
  with scope():
    one_hot = metrics.RankHistogram(
        ensemble_dim='realization', num_bins=num_bins).compute_chunk(
            _ds(forecast), _ds(truth))['geopotential']
  want = {d: n for d, n in zip(fd, f.shape) if d != 'realization'}
  want['bins'] = num_bins
  assert one_hot.sizes == want
  avg = ('prediction_timedelta', 'time', 'latitude', 'longitude')
  sample = np.prod([one_hot.sizes[d] for d in avg])
  rtol = 5 * np.sqrt((num_bins - 1) / sample)  # 5 standard errors
  hist = one_hot.mean(avg).transpose('level', 'bins').values
  np.testing.assert_allclose(hist.sum(axis=1), 1.0, rtol=1e-6)
  np.testing.assert_allclose(hist[0], 1 / num_bins, rtol=rtol)
  if num_bins > 2:
    _decreasing(hist[1][:num_bins // 2 + 1])   # convex
    _increasing(hist[1][num_bins // 2:])
    _increasing(hist[2][:num_bins // 2 + 1])   # concave
    _decreasing(hist[2][num_bins // 2:])
  _increasing(hist[3])
  _decreasing(hist[4])


RANK_HIST_CASES = [(1, None), (10, None), (2, None), (9, 5)]


def test_gaussian_crps_converges():
  case_gaussian_crps_converges(fake_ctx.installed)


@pytest.mark.parametrize('ensemble_size,num_bins', RANK_HIST_CASES)
def test_rank_histogram_calibration(ensemble_size, num_bins):
  case_rank_histogram_calibration(fake_ctx.installed, ensemble_size, num_bins)


# ------------------------------------------------------------------------------
# regridding_test.py:72-249 (test_coarse_grid_interpolates): a smooth periodic
# field regridded to half the resolution stays between the min and max of the
# 2 x 2 source cells nearest to each target node -- for all three regridders,
# both longitude schemes (incl. negative longitudes), with / without pole
# nodes and custom latitudes.  Checks orientation and index alignment.
# ------------------------------------------------------------------------------
S0, C0 = 'START_AT_ZERO', 'CENTER_AT_ZERO'
WP, NP_, CU = ('EQUIANGULAR_WITH_POLES', 'EQUIANGULAR_WITHOUT_POLES', 'CUSTOM')
COARSE_GRID_CASES = [
    # regridder, source lat spacing, source lon scheme, target lat, target lon
    ('ConservativeRegridder', WP, C0, WP, C0),
    ('ConservativeRegridder', WP, S0, WP, S0),
    ('ConservativeRegridder', NP_, S0, NP_, S0),
    ('ConservativeRegridder', WP, S0, WP, C0),
    ('ConservativeRegridder', WP, C0, WP, S0),
    ('ConservativeRegridder', NP_, C0, WP, S0),
    ('BilinearRegridder', WP, S0, WP, C0),
    ('BilinearRegridder', WP, C0, WP, S0),
    ('ConservativeRegridder', WP, S0, WP, C0),
    ('NearestRegridder', NP_, C0, WP, S0),
    ('ConservativeRegridder', CU, S0, CU, S0),
    ('ConservativeRegridder', NP_, S0, CU, S0),
]


def case_coarse_grid_interpolates(scope, regridder, source_lat, source_lon,
                                  target_lat, target_lon):
  from weatherbench2_b200 import regridding as rg, xarray_lite as xl
  n_lats, n_lons, reduce_factor = 32, 64, 2

  def lats(spacing, n):
    if spacing == CU:
      return np.linspace(-87.5, 87.5, n)
    return rg.latitude_values(rg.LatitudeSpacing[spacing], n)

  lat_s = lats(source_lat, n_lats)
  lon_s = rg.longitude_values(rg.LongitudeScheme[source_lon], n_lons)
  theta = 2 * np.pi * lon_s / 360
  phi = 2 * np.pi * (lat_s - 90) / 180
  x = np.sin(phi)[:, None] * (np.cos(theta) ** 2 + np.sin(theta))[None, :]
  lat_t = lats(target_lat, n_lats // reduce_factor)
  lon_t = rg.longitude_values(rg.LongitudeScheme[target_lon],
                              n_lons // reduce_factor)
  # source cells grouped under the target node they are nearest to
  rolled = x if source_lon == target_lon else np.roll(x, n_lons // 2, axis=1)
  blocks = rolled.reshape(n_lats // 2, 2, n_lons // 2, 2)
  lower, upper = blocks.min(axis=(1, 3)), blocks.max(axis=(1, 3))
  source = xl.Dataset(
      {'X': (('time', 'latitude', 'longitude'), x[None])},
      dict(time=np.array(['2000-01-01T00'], dtype='datetime64[ns]'),
           latitude=lat_s, longitude=lon_s))
  grid = lambda lo, la: rg.Grid(longitudes=lo, latitudes=la,  # noqa: E731
                                includes_poles=True, periodic=True)
  with scope():
    out = getattr(rg, regridder)(grid(lon_s, lat_s), grid(lon_t, lat_t)
                                 ).regrid_dataset(source)['X']
  assert out.dims == ('time', 'latitude', 'longitude')
  got = np.asarray(out.values)[0]
  assert got.shape == lower.shape and np.isfinite(got).all()
  np.testing.assert_array_equal(out.coords['longitude'].values, lon_t)
  np.testing.assert_array_equal(out.coords['latitude'].values, lat_t)
  if target_lat == CU:
    min_frac = 0.97
  elif source_lon == target_lon:
    min_frac = 0.99
  else:
    min_frac = 0.5
  assert (lower <= got).mean() >= min_frac
  assert (got <= upper).mean() >= min_frac


@pytest.mark.parametrize('case', COARSE_GRID_CASES,
                         ids=lambda c: '-'.join(c))
def test_coarse_grid_interpolates(case):
  case_coarse_grid_interpolates(fake_ctx.installed, *case)
