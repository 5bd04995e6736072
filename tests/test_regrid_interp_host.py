"""Host logic of the nearest / bilinear regridders (no GPU): the tap tables the
kernel consumes reproduce np.interp (= jnp.interp) on the reference's own test
cases (weatherbench2/regridding_test.py:495-591) and on random grids."""
import numpy as np
import pytest


def _apply(taps, fp):
  i0, i1, t = taps
  fp = np.asarray(fp, dtype=np.float64)
  out = fp[np.maximum(i0, 0)] + t.astype(np.float64) * (
      fp[np.maximum(i1, 0)] - fp[np.maximum(i0, 0)])
  return np.where(i0 < 0, np.nan, out)


@pytest.mark.parametrize('clamp', [True, False])
def test_taps_match_np_interp_non_periodic(clamp):
  from weatherbench2_b200 import regridding  # noqa: F401  (defines Regridder first)
  from weatherbench2_b200 import _regrid_interp as ri
  rs = np.random.RandomState(0)
  xp = np.sort(rs.uniform(-80, 80, 17))
  x = np.concatenate([rs.uniform(-95, 95, 40), xp[:3], [xp[0], xp[-1]]])
  fp = rs.normal(size=xp.size)
  kw = {} if clamp else dict(left=np.nan, right=np.nan)
  want = np.interp(x.astype(np.float32), xp.astype(np.float32), fp, **kw)
  got = _apply(ri.interp_taps(x, xp, clamp=clamp), fp)
  np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6, equal_nan=True)


def test_taps_match_np_interp_periodic():
  from weatherbench2_b200 import regridding  # noqa: F401  (defines Regridder first)
  from weatherbench2_b200 import _regrid_interp as ri
  rs = np.random.RandomState(1)
  for xp in (np.arange(0, 360, 22.5), np.arange(-180, 180, 30.0),
             np.array([0.0, 90.0, 180.0, 270.0])):
    x = np.concatenate([rs.uniform(-360, 720, 50), xp, [359.99, 0.0]])
    fp = rs.normal(size=xp.size)
    want = np.interp(x, xp, fp, period=360)
    got = _apply(ri.interp_taps(x, xp, clamp=True, period=360), fp)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5)


def test_reference_known_answers_through_the_taps():
  from weatherbench2_b200 import regridding  # noqa: F401  (defines Regridder first)
  from weatherbench2_b200 import _regrid_interp as ri
  taps = ri.interp_taps([45.0, 135.0, 225.0, 315.0], [0.0, 90.0, 180.0, 270.0],
                        clamp=True, period=360)
  np.testing.assert_allclose(_apply(taps, [0.0, 1.0, 2.0, 3.0]),
                             [0.5, 1.5, 2.5, 1.5], atol=1e-6)
  taps = ri.interp_taps([45.0, 135.0, 225.0, 315.0], [0.0, 90.0, 180.0, 270.0],
                        clamp=False)
  np.testing.assert_allclose(_apply(taps, [0.0, 1.0, 2.0, 3.0]),
                             [0.5, 1.5, 2.5, np.nan], atol=1e-6)
  taps = ri.interp_taps([-90.0, -30.0, 30.0, 90.0], [-60.0, 0.0, 60.0],
                        clamp=True)
  np.testing.assert_allclose(_apply(taps, [0.0, 1.0, 2.0]),
                             [0.0, 0.5, 1.5, 2.0], atol=1e-6)


def test_nearest_indices_match_the_oracle():
  from oracle import wb2_oracle as orc
  from weatherbench2_b200 import regridding
  lon_s, lat_s = np.arange(0, 360, 10.0), np.linspace(-90, 90, 19)
  lon_t, lat_t = np.arange(2.5, 360, 15.0), np.linspace(-85, 85, 9)
  got = regridding.nearest_neighbor_indices(
      regridding.Grid.from_degrees(lon_s, lat_s),
      regridding.Grid.from_degrees(lon_t, lat_t))
  want = orc.nearest_neighbor_indices(orc.Grid(lon_s, lat_s),
                                      orc.Grid(lon_t, lat_t))
  np.testing.assert_array_equal(got, want)
