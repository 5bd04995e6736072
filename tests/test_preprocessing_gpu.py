"""SURVEY.md section 8 f3 on the device: the ensemble mean of
scripts/compute_ensemble_mean.py, interpolate_spectral_frequencies
(weatherbench2/derived_variables.py:629-682) and the fused latitude-mean
spectrum operator, against the oracle."""
import numpy as np
import pytest

from oracle import wb2_oracle as orc
import wb2_testdata as td

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('skipna', [False, True])
@pytest.mark.parametrize('shape', [(10, 3, 2, 19, 36), (50, 1, 33, 64),
                                   (7, 5, 9, 11)])
def test_ensemble_mean_matches_numpy_mean(skipna, shape):
  from weatherbench2_b200 import preprocessing as pp, xarray_lite as xl
  rs = np.random.RandomState(sum(shape))
  x = (280 + 5 * rs.standard_normal(shape)).astype(np.float32)
  x[rs.rand(*shape) < 0.05] = np.nan
  x[:, 0, ..., 1, 2] = np.nan  # a cell without any valid member
  dims = ('realization', 'time', 'level', 'latitude', 'longitude')[:1] + (
      ('time', 'level', 'latitude', 'longitude')[-(len(shape) - 1):])
  coords = {d: np.arange(n) for d, n in zip(dims, shape)}
  want = orc.ensemble_mean(x, 0, skipna)
  for data in td.host_and_device(x):
    ds = xl.Dataset({'t': (dims, data), 'orog': (dims[-2:], x[0, ..., :, :].reshape(
        (-1,) + shape[-2:])[0])}, coords)
    out = pp.compute_ensemble_mean(ds, skipna=skipna)
    assert out['t'].dims == dims[1:]
    assert 'realization' not in out['t'].coords
    got = out['t'].values
    np.testing.assert_allclose(got, want, rtol=2e-6, equal_nan=True)
    assert out['orog'].dims == dims[-2:]  # no ensemble dim: passes through
  # member axis not leading
  ds = xl.Dataset({'t': (dims[1:] + dims[:1], np.moveaxis(x, 0, -1))}, coords)
  got = pp.compute_ensemble_mean(ds, skipna=skipna)['t'].values
  np.testing.assert_allclose(got, want, rtol=2e-6, equal_nan=True)


def test_ensemble_mean_feeds_k1_on_the_device():
  """EnsembleMeanRMSE two ways: K2's fused statistic and K1 on the device-side
  ensemble mean (metrics.py:1293-1333)."""
  import torch
  from weatherbench2_b200 import metrics, preprocessing as pp, xarray_lite as xl
  rs = np.random.RandomState(1)
  m, nlat, nlon = 8, 37, 72
  lat = np.linspace(-90, 90, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  x = rs.standard_normal((m, 3, nlat, nlon)).astype(np.float32)
  t = rs.standard_normal((3, nlat, nlon)).astype(np.float32)
  dims = ('time', 'latitude', 'longitude')
  coords = {'time': np.arange(3), 'latitude': lat, 'longitude': lon}
  xd = xl.Dataset({'z': (('realization',) + dims, torch.from_numpy(x).cuda())},
                  dict(coords, realization=np.arange(m)))
  td = xl.Dataset({'z': (dims, torch.from_numpy(t).cuda())}, coords)
  mean = pp.compute_ensemble_mean(xd)
  assert mean['z'].data.is_cuda
  a = metrics.RMSESqrtBeforeTimeAvg().compute_chunk(mean, td)['z'].values
  b = metrics.EnsembleMeanRMSESqrtBeforeTimeAvg().compute_chunk(xd, td)['z'].values
  np.testing.assert_allclose(a, b, rtol=1e-5)
  want, _ = orc.ensemble_mean_rmse_sqrt_before_time_avg(
      x, ('realization',) + dims, t, dims, 'realization', lat, lon)
  np.testing.assert_allclose(a, want, rtol=1e-5)


@pytest.mark.parametrize('nlon,nlat', [(72, 19), (240, 33)])
def test_interpolate_spectral_frequencies_matches_scipy_interp1d(nlon, nlat):
  from weatherbench2_b200 import derived_variables as dvs, xarray_lite as xl
  rs = np.random.RandomState(nlon)
  lat = np.linspace(-80, 80, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  x = rs.standard_normal((2, nlat, nlon)).astype(np.float32)
  ds = xl.Dataset({'u': (('time', 'latitude', 'longitude'), x)},
                  {'time': np.arange(2), 'latitude': lat, 'longitude': lon})
  spec = dvs.ZonalEnergySpectrum('u').compute(ds)
  assert spec.dims == ('time', 'latitude', 'zonal_wavenumber')
  for freqs in (None, np.linspace(0, spec.coords['frequency'].values.max(), 17)):
    got = dvs.interpolate_spectral_frequencies(spec, 'zonal_wavenumber', freqs)
    want, wf = orc.interpolate_spectral_frequencies(
        spec.values, spec.coords['frequency'].values, freqs)
    assert got.dims == ('time', 'latitude', 'frequency')
    np.testing.assert_allclose(got.coords['frequency'].values, wf)
    np.testing.assert_array_equal(np.isnan(got.values), np.isnan(want))
    np.testing.assert_allclose(got.values, want, rtol=2e-6, equal_nan=True)
    np.testing.assert_array_equal(got.coords['wavelength'].values, 1 / wf)
  assert np.isnan(got.values).any() and np.isfinite(got.values).any()


@pytest.mark.parametrize('nlon,nlat', [(1440, 41), (240, 121), (72, 19)])
def test_latitude_mean_spectrum_operator(nlon, nlat):
  """ZonalEnergySpectrum.compute_latitude_mean == the weighted latitude mean
  of ZonalEnergySpectrum.compute (both vs the oracle), NumPy and CUDA inputs,
  with a time mean and a latitude band."""
  from weatherbench2_b200 import derived_variables as dvs, xarray_lite as xl
  rs = np.random.RandomState(nlat)
  lat = np.linspace(-90, 90, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  x = (rs.standard_normal((4, 3, nlat, nlon)) + 0.3).astype(np.float32)
  dims = ('time', 'level', 'latitude', 'longitude')
  coords = {'time': np.arange(4), 'level': np.array([200, 500, 850]),
            'latitude': lat, 'longitude': lon}
  op = dvs.ZonalEnergySpectrum('u')
  for data in td.host_and_device(x):
    ds = xl.Dataset({'u': (dims, data)}, coords)
    got = op.compute_latitude_mean(ds)
    want, wd = orc.zonal_energy_spectrum_latitude_mean(x, dims, lat, lon)
    assert got.dims == wd == ('time', 'level', 'zonal_wavenumber')
    tot = want.sum(axis=-1, keepdims=True)
    assert np.max(np.abs(got.values - want) / tot) < 1e-5
    got = op.compute_latitude_mean(ds, time_mean_dim='time',
                                   lat_slice=slice(-60, -30))
    want, _ = orc.zonal_energy_spectrum_latitude_mean(
        x, dims, lat, lon, lat_slice=slice(-60, -30))
    want = want.mean(axis=0)
    assert got.dims == ('level', 'zonal_wavenumber')
    tot = want.sum(axis=-1, keepdims=True)
    assert np.max(np.abs(got.values - want) / tot) < 1e-5
