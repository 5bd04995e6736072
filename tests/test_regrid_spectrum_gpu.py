"""GPU parity tests for K5 (conservative regridding) and K4 (zonal energy
spectrum) against the oracle and the reference tests' known answers
(weatherbench2/regridding_test.py:313-493, derived_variables_test.py:246-435).

Tolerances.  Regridding is float32 in the reference too: 1e-5 relative, NaN
pattern exact.  Spectrum: per-bin relative error is not achievable in float32
even for pocketfft, so bins are compared with an absolute
tolerance of 1e-5 x the row's total power, and total power to 1e-5 relative.
"""
import numpy as np
import pytest

from oracle import wb2_oracle as orc

pytestmark = pytest.mark.gpu


def _grids(sp=True, tp=True, sper=True, tper=True, ns=(20, 10), nt=(15, 8)):
  from weatherbench2_b200 import regridding as rg

  def lats(poles, n):
    return np.linspace(-90, 90, n) if poles else np.linspace(-80, 80, n)

  def lons(periodic, n):
    return (np.linspace(0, 360, n, endpoint=False) if periodic
            else np.linspace(0, 180, n))

  src = dict(longitudes=lons(sper, ns[0]), latitudes=lats(sp, ns[1]),
             includes_poles=sp, periodic=sper)
  tgt = dict(longitudes=lons(tper, nt[0]), latitudes=lats(tp, nt[1]),
             includes_poles=tp, periodic=tper)
  return (rg.ConservativeRegridder(rg.Grid(**src), rg.Grid(**tgt)),
          orc.Grid(**src), orc.Grid(**tgt))


def test_regridding_extrapolation_known_answer():
  """regridding_test.py:313-330."""
  from weatherbench2_b200 import regridding as rg
  kw = dict(includes_poles=False, periodic=False)
  r = rg.ConservativeRegridder(
      rg.Grid(longitudes=np.array([1, 3, 5]), latitudes=np.array([1, 3]),
              **kw),
      rg.Grid(longitudes=np.array([0, 2, 4]), latitudes=np.array([0, 2]),
              **kw))
  actual = r.regrid_array(np.array([[1, 1], [2, 2], [3, 3]]))
  expected = np.array([[np.nan, np.nan], [np.nan, 1.5], [np.nan, 2.5]])
  np.testing.assert_allclose(actual, expected, atol=1e-6)


@pytest.mark.parametrize('sp,tp,sper,tper,expect_nans', [
    (True, True, True, True, False), (False, False, True, True, True),
    (False, True, True, True, True), (True, False, True, True, False),
    (True, True, False, False, True)])
def test_expected_nans_and_oracle(sp, tp, sper, tper, expect_nans):
  """regridding_test.py:332-412 + value parity on random data."""
  r, osrc, otgt = _grids(sp, tp, sper, tper)
  actual = r.regrid_array(np.ones(osrc.shape))
  assert np.isnan(actual).any() == expect_nans
  np.testing.assert_allclose(actual[~np.isnan(actual)], 1.0, atol=1e-6)
  rs = np.random.RandomState(0)
  x = rs.normal(size=(3, 2) + osrc.shape).astype(np.float32)
  x[rs.rand(*x.shape) < 0.05] = np.nan
  got = r.regrid_array(x)
  want = orc.conservative_regrid(x, osrc, otgt)
  assert got.shape == (3, 2) + otgt.shape and got.dtype == np.float32
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_regridding_nans_disc():
  """regridding_test.py:465-493."""
  r, osrc, otgt = _grids(ns=(512, 256), nt=(360, 181))
  slat = np.deg2rad(osrc.latitudes)
  slon = np.deg2rad(osrc.longitudes)
  in_valid = (slat[None, :] ** 2 + (slon[:, None] - np.pi) ** 2
              < (np.pi / 2) ** 2)
  out = r.regrid_array(np.where(in_valid, 1.0, np.nan))
  out_valid = ~np.isnan(out)
  np.testing.assert_allclose(out_valid.mean(), in_valid.mean(), atol=0.01)
  np.testing.assert_allclose(out[out_valid], 1.0, rtol=1e-6)
  want = orc.conservative_regrid(np.where(in_valid, 1.0, np.nan), osrc, otgt)
  np.testing.assert_array_equal(np.isnan(out), np.isnan(want))


def test_quarter_degree_to_1p5_headline_shape():
  """BASELINE.json configs[3] shape: 721x1440 -> 121x240, with a NaN disc."""
  import torch
  from weatherbench2_b200 import regridding as rg
  src = dict(longitudes=np.arange(1440) * 0.25,
             latitudes=np.linspace(-90, 90, 721))
  tgt = dict(longitudes=np.arange(240) * 1.5,
             latitudes=np.linspace(-90, 90, 121))
  r = rg.ConservativeRegridder(rg.Grid.from_degrees(src['longitudes'],
                                                    src['latitudes']),
                               rg.Grid.from_degrees(tgt['longitudes'],
                                                    tgt['latitudes']))
  rs = np.random.RandomState(1)
  x = rs.standard_normal((3, 1440, 721)).astype(np.float32)
  x[1, 300:500, 100:300] = np.nan
  got = r.regrid_array(x)
  want = orc.conservative_regrid(x, orc.Grid(**src), orc.Grid(**tgt))
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6)
  # device-resident input -> device-resident output, same numbers
  gd = r.regrid_array(torch.from_numpy(x).cuda())
  assert gd.is_cuda
  np.testing.assert_array_equal(gd.cpu().numpy(), got)
  # conservation: area-weighted global mean is preserved
  w_src = orc.get_lat_weights(src['latitudes'])
  w_tgt = orc.get_lat_weights(tgt['latitudes'])
  np.testing.assert_allclose((got[0] * w_tgt).mean(), (x[0] * w_src).mean(),
                             atol=2e-6)


def test_regrid_dataset_flips_latitude_and_keeps_dims():
  from weatherbench2_b200 import regridding as rg, xarray_lite as xl
  lon = np.linspace(0, 360, 36, endpoint=False)
  lat = np.linspace(-90, 90, 19)
  r = rg.ConservativeRegridder(
      rg.Grid.from_degrees(lon, lat),
      rg.Grid.from_degrees(np.linspace(0, 360, 12, endpoint=False),
                           np.linspace(-90, 90, 7)))
  rs = np.random.RandomState(2)
  x = rs.normal(size=(4, 19, 36)).astype(np.float32)
  ds = xl.Dataset({'t': (('time', 'latitude', 'longitude'), x[:, ::-1])},
                  {'time': np.arange(4), 'latitude': lat[::-1],
                   'longitude': lon})
  out = r.regrid_dataset(ds)
  assert out['t'].dims == ('time', 'latitude', 'longitude')
  assert out['t'].shape == (4, 7, 12)
  np.testing.assert_array_equal(out['latitude'].values,
                                np.linspace(-90, 90, 7))
  want = orc.conservative_regrid(
      np.transpose(x, (0, 2, 1)),
      orc.Grid(longitudes=lon, latitudes=lat),
      orc.Grid(longitudes=np.linspace(0, 360, 12, endpoint=False),
               latitudes=np.linspace(-90, 90, 7)))
  np.testing.assert_allclose(out['t'].values, np.transpose(want, (0, 2, 1)),
                             rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------- spectrum ----
def _spectrum_case(nlon, nlat=9, outer=(2, 3), seed=0, offset=0.0):
  rs = np.random.RandomState(seed)
  lat = np.linspace(-80, 80, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  x = (rs.standard_normal(outer + (nlat, nlon)) + offset).astype(np.float32)
  dims = ('time', 'level', 'latitude', 'longitude')
  return x, dims, lat, lon


def _check_spectrum(got, want):
  power = want.sum(axis=-1, keepdims=True)
  np.testing.assert_allclose(got.sum(axis=-1), want.sum(axis=-1), rtol=1e-5)
  assert np.max(np.abs(got - want) / power) < 1e-5


@pytest.mark.parametrize('nlon', [2, 4, 12, 18, 36, 72, 64, 240, 360, 1440])
def test_spectrum_matches_oracle(nlon):
  from weatherbench2_b200 import derived_variables as dvs, xarray_lite as xl
  x, dims, lat, lon = _spectrum_case(nlon, offset=3.0 if nlon == 72 else 0.0)
  if nlon == 2:
    lon = np.array([0.0, 180.0])
  ds = xl.Dataset({'u': (dims, x)},
                  {'time': np.arange(2), 'level': np.arange(3),
                   'latitude': lat, 'longitude': lon})
  got = dvs.ZonalEnergySpectrum('u').compute(ds)
  want, wd, freq, wl = orc.zonal_energy_spectrum(x, dims, lat, lon)
  assert got.dims == wd
  assert got.shape[-1] == nlon // 2 + 1
  _check_spectrum(got.values.astype(np.float64), want)
  assert got.coords['frequency'].dims == ('zonal_wavenumber', 'latitude')
  np.testing.assert_allclose(got.coords['frequency'].values, freq)
  np.testing.assert_array_equal(got.coords['wavelength'].values,
                                1 / got.coords['frequency'].values)
  assert got.coords['frequency'].attrs['units'] == '1 / m'


@pytest.mark.parametrize('nlon', [64, 240, 360, 512, 720, 1440])
def test_spectrum_fixed_plan_vs_generic_kernel(monkeypatch, nlon):
  """Sizes with a compile-time radix plan (spectrum_fixed_kernel) against the
  generic runtime-plan kernel and the oracle; 23 latitude rows so that the
  last CTA owns a ragged row block."""
  from weatherbench2_b200 import derived_variables as dvs, xarray_lite as xl
  x, dims, lat, lon = _spectrum_case(nlon, nlat=23, outer=(2, 2), seed=nlon)
  ds = xl.Dataset({'u': (dims, x)},
                  {'time': np.arange(2), 'level': np.arange(2),
                   'latitude': lat, 'longitude': lon})
  want, _, _, _ = orc.zonal_energy_spectrum(x, dims, lat, lon)
  res = {}
  for path in ('fixed', 'generic'):
    monkeypatch.setenv('WB2_SPECTRUM_PATH', path)
    res[path] = dvs.ZonalEnergySpectrum('u').compute(ds).values.astype(
        np.float64)
    _check_spectrum(res[path], want)
  power = want.sum(axis=-1, keepdims=True)
  assert np.max(np.abs(res['fixed'] - res['generic']) / power) < 2e-6


def test_spectrum_longitude_first_layout_and_peak():
  """Mock layout (..., longitude, latitude) and the spectral-peak test of
  derived_variables_test.py:290-321."""
  from weatherbench2_b200 import derived_variables as dvs, xarray_lite as xl
  nlon, nlat = 36, 19
  lat = np.linspace(-90, 90, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  rs = np.random.RandomState(4)
  x = rs.standard_normal((2, nlon, nlat)).astype(np.float32)
  x += (10 * np.cos(2 * np.pi * lon / 100))[None, :, None].astype(np.float32)
  dims = ('time', 'longitude', 'latitude')
  ds = xl.Dataset({'z': (dims, x)}, {'time': np.arange(2), 'latitude': lat,
                                     'longitude': lon})
  got = dvs.ZonalEnergySpectrum('z').compute(ds)
  want, wd, freq, _ = orc.zonal_energy_spectrum(x, dims, lat, lon)
  assert got.dims == wd == ('time', 'latitude', 'zonal_wavenumber')
  _check_spectrum(got.values.astype(np.float64)[:, 1:-1], want[:, 1:-1])
  for ilat in (9, 12, 15):  # 0, 30, 60 degrees
    wavelength_m = (100 / 360) * 2 * np.pi * orc.EARTH_RADIUS_M * np.cos(
        np.deg2rad(lat[ilat]))
    k_expected = int(np.argmin(np.abs(freq[:, ilat] - 1 / wavelength_m)))
    assert (np.argmax(got.values[:, ilat], axis=-1) == k_expected).all()


def test_spectrum_parseval_and_time_sum():
  """Parseval (derived_variables_test.py:409-435) and the device-side time sum
  (scripts/compute_zonal_energy_spectrum.py:234 = sum / count)."""
  from weatherbench2_b200 import derived_variables as dvs, xarray_lite as xl
  x, dims, lat, lon = _spectrum_case(72, nlat=13, outer=(5, 2), seed=3,
                                     offset=2.0)
  ds = xl.Dataset({'u': (dims, x)},
                  {'time': np.arange(5), 'level': np.arange(2),
                   'latitude': lat, 'longitude': lon})
  dv = dvs.ZonalEnergySpectrum('u')
  spec = dv.compute(ds)
  spacing = dv.lon_spacing_m(ds).values
  energy = (spacing[None, None, :, None] * x.astype(np.float64) ** 2).sum(-1)
  # White noise has Nyquist content, which the reference counts twice
  # (derived_variables.py:600): sum_k S_k = (C/L) sum f^2 + C |F_{L/2}|^2.
  nyq = np.abs(np.fft.rfft(x.astype(np.float64), axis=-1,
                           norm='forward')[..., -1]) ** 2
  circ = dv._circumference(lat)[None, None, :]
  np.testing.assert_allclose(spec.values.sum(-1), energy + circ * nyq,
                             rtol=1e-5)
  summed = dv.compute(ds, time_sum_dim='time')
  assert summed.dims == ('level', 'latitude', 'zonal_wavenumber')
  np.testing.assert_allclose(summed.values, spec.values.sum(axis=0),
                             rtol=2e-6, atol=1e-3 * spec.values.max() * 1e-3)


def test_spectrum_unsupported_sizes_raise():
  from weatherbench2_b200 import _lib, derived_variables as dvs
  from weatherbench2_b200 import xarray_lite as xl
  for nlon in (14, 9):
    x, dims, lat, lon = _spectrum_case(nlon)
    ds = xl.Dataset({'u': (dims, x)}, {'time': np.arange(2),
                                       'level': np.arange(3), 'latitude': lat,
                                       'longitude': lon})
    with pytest.raises(_lib.Wb2Error):
      dvs.ZonalEnergySpectrum('u').compute(ds)
  x, dims, lat, lon = _spectrum_case(12)
  lon = lon.copy()
  lon[3] += 1.0
  ds = xl.Dataset({'u': (dims, x)}, {'time': np.arange(2),
                                     'level': np.arange(3), 'latitude': lat,
                                     'longitude': lon})
  with pytest.raises(ValueError):
    dvs.ZonalEnergySpectrum('u').compute(ds)
