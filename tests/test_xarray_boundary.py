"""The real-xarray drop-in boundary (xarray_lite.from_xarray / to_xarray),
driven with a stand-in `xarray` package (tests/fake_xarray) because xarray is
not installable here.  CPU part: lossless round trip.  GPU part: the metric
classes called with xr.Dataset arguments return xr.Dataset results equal to the
oracle (the reference's calling convention, weatherbench2/metrics.py:88-115)."""
import importlib
import os
import sys

import numpy as np
import pytest

from oracle import wb2_oracle as orc

FAKE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fake_xarray')


@pytest.fixture
def xr(monkeypatch):
  monkeypatch.syspath_prepend(FAKE)
  sys.modules.pop('xarray', None)
  mod = importlib.import_module('xarray')
  assert mod.__version__ == '0.0-fake'
  yield mod
  sys.modules.pop('xarray', None)


def _case():
  rs = np.random.RandomState(0)
  lat = np.linspace(-90, 90, 19)
  lon = np.linspace(0, 360, 36, endpoint=False)
  dims = ('time', 'level', 'latitude', 'longitude')
  f = rs.normal(size=(2, 3, 19, 36)).astype(np.float32)
  t = rs.normal(size=(2, 3, 19, 36)).astype(np.float32)
  coords = {'time': np.array(['2020-01-01', '2020-01-02'], 'datetime64[ns]'),
            'level': np.array([500, 700, 850]), 'latitude': lat,
            'longitude': lon}
  return f, t, dims, coords, lat, lon


def test_round_trip_through_the_lite_container(xr):
  from weatherbench2_b200 import xarray_lite as xl
  f, _, dims, coords, _, _ = _case()
  ds = xr.Dataset({'z': (dims, f, {'units': 'm'})}, coords=coords,
                  attrs={'title': 'x'})
  assert xl.have_xarray() and xl.is_native_xarray(ds)
  lite = xl.from_xarray(ds)
  assert isinstance(lite, xl.Dataset) and lite['z'].dims == dims
  assert lite['z'].values is f or np.shares_memory(lite['z'].values, f)
  assert lite['z'].attrs == {'units': 'm'} and lite.attrs == {'title': 'x'}
  np.testing.assert_array_equal(lite['latitude'].values, coords['latitude'])
  back = xl.to_xarray(lite)
  assert type(back).__module__ == 'xarray' and isinstance(back, xr.Dataset)
  assert back['z'].dims == dims
  np.testing.assert_array_equal(back['z'].values, f)
  np.testing.assert_array_equal(back.coords['time'].values, coords['time'])
  da = xl.to_xarray(lite['z'])
  assert isinstance(da, xr.DataArray) and da.name == 'z'
  assert xl.from_xarray(da).dims == dims
  assert xl.from_xarray(lite) is lite  # lite containers pass through


@pytest.mark.gpu
def test_metric_classes_take_and_return_xarray(xr):
  from weatherbench2_b200 import metrics, regions as R
  f, t, dims, coords, lat, lon = _case()
  fds = xr.Dataset({'z': (dims, f)}, coords=coords)
  tds = xr.Dataset({'z': (dims, t)}, coords=coords)
  got = metrics.MSE().compute_chunk(fds, tds)
  assert isinstance(got, xr.Dataset)
  assert got['z'].dims == ('time', 'level')
  want, _ = orc.mse(f, dims, t, dims, lat, lon)
  np.testing.assert_allclose(got['z'].values, want, rtol=1e-5)
  got = metrics.RMSESqrtBeforeTimeAvg().compute(
      fds, tds, region=R.SliceRegion(lat_slice=slice(-20, 20)))
  assert isinstance(got, xr.Dataset) and got['z'].dims == ('level',)
  want, wd = orc.rmse_sqrt_before_time_avg(
      f, dims, t, dims, lat, lon,
      region=orc.SliceRegion(lat_slice=slice(-20, 20)))
  np.testing.assert_allclose(got['z'].values, want.mean(axis=0), rtol=1e-5)
  # ensemble metric + the ensemble_size attr (metrics.py:598-607)
  rs = np.random.RandomState(1)
  x = rs.normal(size=(5,) + f.shape).astype(np.float32)
  xds = xr.Dataset({'z': (('realization',) + dims, x)},
                   coords=dict(coords, realization=np.arange(5)))
  got = metrics.CRPS().compute(xds, tds)
  assert isinstance(got, xr.Dataset)
  want, wd = orc.crps(x, ('realization',) + dims, t, dims, 'realization', lat,
                      lon)
  np.testing.assert_allclose(got['z'].values, want.mean(axis=0), rtol=1e-5,
                             atol=1e-6)
  assert got.attrs['ensemble_size'] == 5


@pytest.mark.gpu
def test_spectrum_and_regridder_take_xarray(xr):
  from weatherbench2_b200 import derived_variables as dvs, regridding as rg
  rs = np.random.RandomState(2)
  lat = np.linspace(-90, 90, 19)
  lon = np.linspace(0, 360, 36, endpoint=False)
  x = rs.normal(size=(2, 19, 36)).astype(np.float32)
  ds = xr.Dataset({'u': (('time', 'latitude', 'longitude'), x)},
                  coords={'time': np.arange(2), 'latitude': lat,
                          'longitude': lon})
  got = dvs.ZonalEnergySpectrum('u').compute(ds)
  assert isinstance(got, xr.DataArray)
  want, wd, _, _ = orc.zonal_energy_spectrum(
      x, ('time', 'latitude', 'longitude'), lat, lon)
  assert got.dims == wd
  power = want.sum(axis=-1, keepdims=True)
  assert np.max(np.abs(got.values - want) / power) < 1e-5
  # regridder on the reference's (lon, lat) layout
  src = rg.Grid.from_degrees(lon, lat)
  tgt = rg.Grid.from_degrees(np.linspace(0, 360, 12, endpoint=False),
                             np.linspace(-90, 90, 7))
  y = rs.normal(size=(2, 36, 19)).astype(np.float32)
  dsy = xr.Dataset({'u': (('time', 'longitude', 'latitude'), y)},
                   coords={'time': np.arange(2), 'latitude': lat,
                           'longitude': lon})
  out = rg.ConservativeRegridder(src, tgt).regrid_dataset(dsy)
  assert isinstance(out, xr.Dataset)
  want = orc.conservative_regrid(
      y, orc.Grid(lon, lat),
      orc.Grid(np.linspace(0, 360, 12, endpoint=False),
               np.linspace(-90, 90, 7)))
  np.testing.assert_allclose(out['u'].values, want, rtol=1e-5, atol=1e-6)
