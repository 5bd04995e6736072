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
