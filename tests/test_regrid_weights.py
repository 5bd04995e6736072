"""CPU-only: the regridder's weight matrices (built in the product with the
reference's formulas) against the reference's known answers
(weatherbench2/regridding_test.py:252-311) and against the oracle; CSR
conversion round trip."""
import numpy as np
import pytest

from oracle import wb2_oracle as orc
from weatherbench2_b200 import _lib
from weatherbench2_b200 import regridding as rg


def test_latitude_weights_known_answer():
  expected = np.array([
      [1 - np.sqrt(3) / 2, (np.sqrt(3) - 1) / 2, 1 / 2, 0, 0, 0],
      [0, 0, 0, 1 / 2, (np.sqrt(3) - 1) / 2, 1 - np.sqrt(3) / 2]])
  actual = rg._conservative_latitude_weights(
      np.array([-75, -45, -15, 15, 45, 75]), np.array([-45, 45]), True, True)
  np.testing.assert_almost_equal(expected, actual)


def test_longitude_weights_known_answers():
  expected = np.array([[4, 1, 0, 0, 0, 1], [0, 3, 3, 0, 0, 0],
                       [0, 0, 1, 4, 1, 0], [0, 0, 0, 0, 3, 3]]) / 6
  actual = rg._conservative_longitude_weights(
      np.array([0, 60, 120, 180, 240, 300]), np.array([0, 90, 180, 270]),
      True, True)
  np.testing.assert_allclose(expected, actual, atol=1e-5)
  actual = rg._conservative_longitude_weights(
      np.array([90, 180, 270, 360]), np.array([-270, -180, -90, 0]), True,
      True)
  np.testing.assert_allclose(np.eye(4), actual, atol=1e-5)
  with pytest.raises(ValueError):
    rg._conservative_longitude_weights(np.arange(8) * 45.0,
                                       np.array([0.0, 180.0]), True, True)


@pytest.mark.parametrize('sp,tp,sper,tper', [
    (True, True, True, True), (False, False, True, True),
    (False, True, True, True), (True, False, True, True),
    (True, True, False, False)])
def test_weights_equal_oracle_bitwise(sp, tp, sper, tper):
  def lats(poles, n):
    return np.linspace(-90, 90, n) if poles else np.linspace(-80, 80, n)

  def lons(periodic, n):
    return (np.linspace(0, 360, n, endpoint=False) if periodic
            else np.linspace(0, 180, n))

  a = rg._conservative_latitude_weights(lats(sp, 10), lats(tp, 8), sp, tp)
  b = orc.conservative_latitude_weights(lats(sp, 10), lats(tp, 8), sp, tp)
  np.testing.assert_array_equal(a, b)
  a = rg._conservative_longitude_weights(lons(sper, 20), lons(tper, 15), sper,
                                         tper)
  b = orc.conservative_longitude_weights(lons(sper, 20), lons(tper, 15), sper,
                                         tper)
  np.testing.assert_array_equal(a, b)


def test_quarter_degree_weights_and_csr():
  wlon = rg._conservative_longitude_weights(
      np.arange(1440) * 0.25, np.arange(240) * 1.5, True, True)
  wlat = rg._conservative_latitude_weights(
      np.linspace(-90, 90, 721), np.linspace(-90, 90, 121), True, True)
  np.testing.assert_array_equal(wlon, orc.conservative_longitude_weights(
      np.arange(1440) * 0.25, np.arange(240) * 1.5, True, True))
  np.testing.assert_array_equal(wlat, orc.conservative_latitude_weights(
      np.linspace(-90, 90, 721), np.linspace(-90, 90, 121), True, True))
  for dense in (wlon, wlat):
    csr = _lib.CsrSpec(dense)
    back = np.zeros_like(dense)
    for i in range(csr.n_tgt):
      s, e = csr.row_ptr[i], csr.row_ptr[i + 1]
      back[i, csr.col_idx[s:e]] = csr.val[s:e]
    np.testing.assert_array_equal(back, dense)
    assert not csr.nan_row.any()
    assert (np.diff(csr.row_ptr) <= 8).all()  # banded: <= 8 taps per target


def test_csr_marks_uncovered_rows():
  w = rg._conservative_latitude_weights(np.linspace(-80, 80, 10),
                                        np.linspace(-90, 90, 8), False, True)
  csr = _lib.CsrSpec(w)
  np.testing.assert_array_equal(csr.nan_row.astype(bool),
                                np.isnan(w).any(axis=1))
  assert csr.nan_row.any()


def test_grid_validation_and_hash():
  with pytest.raises(ValueError):
    rg.Grid.from_degrees(np.arange(4) * 90.0, np.array([90.0, 0.0, -90.0]))
  g1 = rg.Grid.from_degrees(np.arange(4) * 90.0, np.array([-90.0, 0.0, 90.0]))
  g2 = rg.Grid.from_degrees(np.arange(4) * 90.0, np.array([-90.0, 0.0, 90.0]))
  assert g1 == g2 and hash(g1) == hash(g2) and g1.shape == (4, 3)


def test_grid_node_generators_and_coverage_check():
  """regridding.py:43-114: the helpers scripts/regrid.py builds target grids
  with."""
  from weatherbench2_b200 import regridding as rg
  lat = rg.latitude_values(rg.LatitudeSpacing.EQUIANGULAR_WITH_POLES, 121)
  np.testing.assert_allclose(lat, np.linspace(-90, 90, 121))
  lat = rg.latitude_values(rg.LatitudeSpacing.EQUIANGULAR_WITHOUT_POLES, 32)
  np.testing.assert_allclose(lat, -90 + (np.arange(32) + 0.5) * 180 / 32)
  with pytest.raises(ValueError, match='Unhandled'):
    rg.latitude_values(rg.LatitudeSpacing.CUSTOM, 10)
  lon = rg.longitude_values(rg.LongitudeScheme.START_AT_ZERO, 240)
  np.testing.assert_allclose(lon, np.arange(240) * 1.5)
  lon = rg.longitude_values(rg.LongitudeScheme.CENTER_AT_ZERO, 64)
  np.testing.assert_allclose(lon, -180 + (np.arange(64) + 0.5) * 360 / 64)
  assert lon[0] == -lon[-1]
  # the generated axes feed Grid / the weight builders directly
  grid = rg.Grid(longitudes=rg.longitude_values(
      rg.LongitudeScheme.START_AT_ZERO, 8), latitudes=rg.latitude_values(
          rg.LatitudeSpacing.EQUIANGULAR_WITH_POLES, 5), periodic=True,
                 includes_poles=True)
  assert grid.shape == (8, 5)
  ok_lat = np.linspace(-90, 90, 19)
  rg._check_global_coverage(np.arange(0, 360.5, 10.0), ok_lat, 0.5)
  rg._check_global_coverage(np.arange(-180, 180.5, 10.0), ok_lat, 0.5)
  with pytest.raises(ValueError, match='min latitude'):
    rg._check_global_coverage(np.arange(0, 361, 10.0), ok_lat[1:], 0.5)
  with pytest.raises(ValueError, match='max latitude'):
    rg._check_global_coverage(np.arange(0, 361, 10.0), ok_lat[:-1], 0.5)
  with pytest.raises(ValueError, match='min longitude'):
    rg._check_global_coverage(np.arange(10, 361, 10.0), ok_lat, 0.5)
  with pytest.raises(ValueError, match='max longitude'):
    rg._check_global_coverage(np.arange(0, 350, 10.0), ok_lat, 0.5)
