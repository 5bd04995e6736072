"""GPU parity tests for K10 (RankHistogram, weatherbench2/metrics.py:1894-2042)
mirroring weatherbench2/metrics_test.py:536-668."""
import numpy as np
import pytest

from oracle import wb2_oracle as orc
import wb2_testdata as td

pytestmark = pytest.mark.gpu


def _ds(vars, coords):  # pylint: disable=redefined-builtin
  from weatherbench2_b200 import xarray_lite as xl
  return xl.Dataset({k: (d, v) for k, (d, v) in vars.items()}, coords)


def _pair(ensemble_size, **kw):
  truth, forecast = td.get_random_truth_and_forecast(
      variables=['geopotential'], ensemble_size=ensemble_size,
      lead_stop='1 day', time_stop='2019-12-03', time_resolution='12 hours',
      spatial_resolution_in_degrees=30, **kw)
  for ds in (truth, forecast):
    for k, (d, v) in ds['vars'].items():
      ds['vars'][k] = (d, v.astype(np.float32))
  return truth, forecast


@pytest.mark.parametrize('ensemble_size,num_bins', [(1, None), (4, None),
                                                    (9, 5), (50, None),
                                                    (51, 4)])
def test_rank_one_hot_matches_oracle_without_ties(ensemble_size, num_bins):
  """Continuous data has no ties: identical to the reference with or without
  random tie-breaking; NaN values rank last (metrics.py:1911)."""
  from weatherbench2_b200 import metrics
  truth, forecast = _pair(ensemble_size)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  f = f.copy()
  t = t.copy()
  f.flat[3::37] = np.nan
  t.flat[5::29] = np.nan
  forecast['vars']['geopotential'] = (fd, f)
  truth['vars']['geopotential'] = (tdm, t)
  fds, tds = _ds(**forecast), _ds(**truth)
  want, wd = orc.rank_histogram_one_hot(f, fd, t, tdm, 'realization', num_bins)
  for ties in (True, False):
    metric = metrics.RankHistogram(num_bins=num_bins,
                                   break_ties_randomly=ties, seed=1)
    got = metric.compute_chunk(fds, tds)['geopotential']
    assert got.dims[-1] == 'bins'
    a, b, _ = orc.align(np.asarray(got.values), got.dims, want, wd)
    np.testing.assert_array_equal(a, b)
    # time mean of the one-hots = the histogram (EnsembleMetric.compute)
    res = metric.compute(fds, tds)
    assert res.attrs['ensemble_size'] == ensemble_size
    hist, hd = orc.time_mean(want, wd, avg_dim='time')
    a, b, _ = orc.align(np.asarray(res['geopotential'].values),
                        res['geopotential'].dims, hist, hd)
    np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(a.sum(axis=-1), 1.0, rtol=1e-6)


@pytest.mark.parametrize('ensemble_size', [1, 4, 7])
def test_repeated_entries_get_random_bin(ensemble_size):
  """metrics_test.py:612-649: truth and all members identical -> every bin is
  equally likely; without random tie-breaking the truth always ranks first."""
  from weatherbench2_b200 import metrics
  truth, forecast = _pair(ensemble_size, )
  for ds in (truth, forecast):
    for k, (d, v) in ds['vars'].items():
      ds['vars'][k] = (d, np.zeros_like(v))
  fds, tds = _ds(**forecast), _ds(**truth)
  num_bins = ensemble_size + 1
  got = metrics.RankHistogram(seed=802701).compute_chunk(fds, tds)[
      'geopotential']
  v = np.asarray(got.values)
  sample_size = v.size // num_bins
  hist = v.reshape(-1, num_bins).mean(axis=0)
  rtol = 5 * (num_bins - 1) / np.sqrt(sample_size)  # >= 5 standard errors
  np.testing.assert_allclose(hist, 1 / num_bins, rtol=rtol)
  # reproducible for a given seed, different for another one
  again = metrics.RankHistogram(seed=802701).compute_chunk(fds, tds)[
      'geopotential']
  np.testing.assert_array_equal(np.asarray(again.values), v)
  other = metrics.RankHistogram(seed=7).compute_chunk(fds, tds)['geopotential']
  assert (np.asarray(other.values) != v).any()
  fixed = metrics.RankHistogram(break_ties_randomly=False).compute_chunk(
      fds, tds)['geopotential']
  assert (np.asarray(fixed.values)[..., 0] == 1).all()


def test_partial_ties_stay_within_the_tied_bins():
  """Truth equal to 2 of 5 members: the rank is uniform on {below, ..., below+2}."""
  from weatherbench2_b200 import metrics
  truth, forecast = _pair(5)
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  f = np.zeros_like(f)
  ax = fd.index('realization')
  vals = np.array([-2.0, -1.0, 0.5, 0.5, 3.0], np.float32)
  f += vals.reshape([5 if i == ax else 1 for i in range(f.ndim)])
  t = np.full_like(t, 0.5)
  forecast['vars']['geopotential'] = (fd, f)
  truth['vars']['geopotential'] = (tdm, t)
  got = metrics.RankHistogram(seed=3).compute_chunk(
      _ds(**forecast), _ds(**truth))['geopotential']
  hist = np.asarray(got.values).reshape(-1, 6).mean(axis=0)
  assert hist[[0, 1, 5]].sum() == 0
  np.testing.assert_allclose(hist[2:5], 1 / 3, rtol=0.2)


def test_bad_num_bins_raises():
  from weatherbench2_b200 import metrics
  truth, forecast = _pair(4)
  with pytest.raises(ValueError, match='Cannot bin'):
    metrics.RankHistogram(num_bins=3).compute_chunk(_ds(**forecast),
                                                    _ds(**truth))
