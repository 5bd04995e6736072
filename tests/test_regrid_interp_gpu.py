"""GPU parity tests for K8 (nearest / bilinear regridders) against the oracle
and the reference's known answers (weatherbench2/regridding_test.py:416-591)."""
import numpy as np
import pytest

from oracle import wb2_oracle as orc

pytestmark = pytest.mark.gpu


def _grids(regridding, lon_s, lat_s, lon_t, lat_t, **kw):
  return (regridding.Grid(longitudes=np.asarray(lon_s), latitudes=np.asarray(lat_s), **kw),
          regridding.Grid(longitudes=np.asarray(lon_t), latitudes=np.asarray(lat_t), **kw),
          orc.Grid(np.asarray(lon_s), np.asarray(lat_s), **kw),
          orc.Grid(np.asarray(lon_t), np.asarray(lat_t), **kw))


@pytest.mark.parametrize('periodic,expected', [
    (True, [[0.5], [1.5], [2.5], [1.5]]), (False, [[0.5], [1.5], [2.5], [np.nan]])])
def test_bilinear_longitude_periodicity_known_answers(periodic, expected):
  from weatherbench2_b200 import regridding
  src, tgt, *_ = _grids(regridding, [0.0, 90.0, 180.0, 270.0], [0],
                        [45.0, 135.0, 225.0, 315.0], [0], includes_poles=True,
                        periodic=periodic)
  got = regridding.BilinearRegridder(src, tgt).regrid_array(
      np.array([[0.0], [1.0], [2.0], [3.0]]))
  np.testing.assert_allclose(got, expected, atol=1e-6)


@pytest.mark.parametrize('poles,slat,tlat,vals,expected', [
    (True, [-90.0, -30.0, 30.0, 90.0], [-60.0, 0.0, 60.0], [0.0, 1.0, 2.0, 3.0],
     [[0.5, 1.5, 2.5]]),
    (True, [-60.0, 0.0, 60.0], [-90.0, -30.0, 30.0, 90.0], [0.0, 1.0, 2.0],
     [[0.0, 0.5, 1.5, 2.0]]),
    (False, [-60.0, -20.0, 20.0, 60.0], [-70.0, 0.0, 70.0],
     [0.0, 1.0, 2.0, 3.0], [[np.nan, 1.5, np.nan]])])
def test_bilinear_latitude_poles_known_answers(poles, slat, tlat, vals,
                                               expected):
  from weatherbench2_b200 import regridding
  src, tgt, *_ = _grids(regridding, [0.0], slat, [0.0], tlat,
                        includes_poles=poles, periodic=True)
  got = regridding.BilinearRegridder(src, tgt).regrid_array(
      np.array(vals)[np.newaxis, :])
  np.testing.assert_allclose(got, expected, atol=1e-6)


def test_nearest_exact_known_answer():
  from weatherbench2_b200 import regridding
  src, tgt, *_ = _grids(regridding, [0, 90, 180, 270], [-30, 0, 30], [0, 180],
                        [-30, 0, 30], includes_poles=True, periodic=True)
  field = np.array([[0, 1, 2], [4, 5, 6], [7, 8, 9], [10, 11, 12]])
  got = regridding.NearestRegridder(src, tgt).regrid_array(field)
  np.testing.assert_allclose(got, [[0, 1, 2], [7, 8, 9]], atol=1e-6)


@pytest.mark.parametrize('periodic,poles', [(True, True), (False, False),
                                            (True, False)])
def test_regridders_match_oracle_on_random_fields(periodic, poles):
  from weatherbench2_b200 import regridding
  rs = np.random.RandomState(4)
  lon_s = np.arange(0, 360, 7.5)
  lat_s = np.linspace(-90, 90, 25) if poles else np.linspace(-86.25, 86.25, 24)
  lon_t = np.arange(-180, 180, 20.0) + (0 if periodic else 181)
  lat_t = np.linspace(-90, 90, 10)
  src, tgt, osrc, otgt = _grids(regridding, lon_s, lat_s, lon_t, lat_t,
                                includes_poles=poles, periodic=periodic)
  field = rs.normal(size=(3, 2, lon_s.size, lat_s.size)).astype(np.float32)
  field[0, 0, 5, 7] = np.nan
  got = regridding.BilinearRegridder(src, tgt).regrid_array(field)
  want = orc.bilinear_regrid(field, osrc, otgt)
  assert got.shape == want.shape and got.dtype == np.float32
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6, equal_nan=True)
  got = regridding.NearestRegridder(src, tgt).regrid_array(field)
  want = orc.nearest_regrid(field, osrc, otgt)
  np.testing.assert_array_equal(got, want)  # a pure gather: bit-exact


def test_regrid_dataset_and_device_tensors():
  """regrid_dataset keeps dims / coords (regridding_test.py:416-463) and CUDA
  tensors stay on the device."""
  import torch
  from weatherbench2_b200 import regridding, xarray_lite as xl
  lon_s, lat_s = np.arange(0, 360, 5.625), np.linspace(-87.1875, 87.1875, 32)
  lon_t, lat_t = np.arange(0, 360, 11.25), np.linspace(-84.375, 84.375, 16)
  src, tgt, osrc, otgt = _grids(regridding, lon_s, lat_s, lon_t, lat_t,
                                includes_poles=False, periodic=True)
  rs = np.random.RandomState(0)
  x = rs.normal(size=(2, lat_s.size, lon_s.size)).astype(np.float32)
  ds = xl.Dataset({'x': (('time', 'latitude', 'longitude'), x)},
                  {'time': np.arange(2), 'latitude': lat_s, 'longitude': lon_s})
  for cls, fn in ((regridding.BilinearRegridder, orc.bilinear_regrid),
                  (regridding.NearestRegridder, orc.nearest_regrid)):
    out = cls(src, tgt).regrid_dataset(ds)['x']
    assert out.dims == ('time', 'latitude', 'longitude')
    assert out.shape == (2, 16, 32)
    want = fn(np.swapaxes(x, -1, -2), osrc, otgt)
    np.testing.assert_allclose(out.values, np.swapaxes(want, -1, -2),
                               rtol=2e-5, atol=2e-6)
    dev = torch.from_numpy(np.swapaxes(x, -1, -2).copy()).cuda()
    res = cls(src, tgt).regrid_array(dev)
    assert res.is_cuda
    np.testing.assert_allclose(res.cpu().numpy(), want, rtol=2e-5, atol=2e-6)


def test_bad_indices_are_rejected():
  from weatherbench2_b200 import _lib
  ctx = _lib.default_context(0)
  buf = ctx.malloc(1024)
  with pytest.raises(_lib.Wb2Error, match='out of range'):
    ctx.regrid_gather(buf, buf, 1, 16, 4, 16, np.array([0, 3, 16, 2]))
  ctx.free(buf)
