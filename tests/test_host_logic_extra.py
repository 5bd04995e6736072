"""More host-side logic that needs no GPU: the per-context call lock, threshold
spec validation, SEEPS climatology handling, rank-histogram binning."""
import threading
import time

import numpy as np
import pytest

import test_threshold_metrics_gpu as thr_helpers
import wb2_testdata as td


def test_locked_lib_serialises_calls():
  from weatherbench2_b200 import _lib

  class FakeLib:
    def __init__(self):
      self.inside = 0
      self.max_inside = 0

    def wb2_work(self, x):
      self.inside += 1
      self.max_inside = max(self.max_inside, self.inside)
      time.sleep(0.002)
      self.inside -= 1
      return x + 1

  fake = FakeLib()
  locked = _lib._LockedLib(fake, threading.RLock())  # pylint: disable=protected-access
  out = []
  threads = [threading.Thread(target=lambda i=i: out.append(locked.wb2_work(i)))
             for i in range(16)]
  for t in threads:
    t.start()
  for t in threads:
    t.join()
  assert sorted(out) == list(range(1, 17))
  assert fake.max_inside == 1
  with pytest.raises(AttributeError):
    locked.wb2_missing  # pylint: disable=pointless-statement


def test_threshold_spec_validation():
  from weatherbench2_b200 import _spatial as sp, _thresholded, thresholds
  truth, _, clim, *_ = thr_helpers._random_case(None, False)  # pylint: disable=protected-access
  tds = thr_helpers._ds(**truth)  # pylint: disable=protected-access
  t_da = tds['geopotential']
  layout = sp.prepare_operand(t_da, None, np.float32).layout
  g1 = thresholds.GaussianQuantileThreshold(clim, 0.2)
  g2 = thresholds.GaussianQuantileThreshold(clim, 0.8)
  kind, m_op, s_op, z = _thresholded._threshold_spec(  # pylint: disable=protected-access
      [g1, g2], tds, 'geopotential', t_da, layout)
  assert kind == 'gaussian' and len(z) == 2 and z[0] < 0 < z[1]
  assert m_op.nrow == s_op.nrow
  with pytest.raises(ValueError, match='at least one'):
    _thresholded._threshold_spec([], tds, 'geopotential', t_da, layout)  # pylint: disable=protected-access
  other = thr_helpers._ds(  # pylint: disable=protected-access
      {k: (clim[k].dims, np.asarray(clim[k].values)) for k in clim.keys()},
      {k: np.asarray(c.values) for k, c in clim.coords.items()})
  g3 = thresholds.GaussianQuantileThreshold(other, 0.5)
  with pytest.raises(ValueError, match='share'):
    _thresholded._threshold_spec([g1, g3], tds, 'geopotential', t_da, layout)  # pylint: disable=protected-access
  # the metric classes keep the reference's positional signature
  from weatherbench2_b200 import metrics
  m = metrics.EnsembleBrierScore([g1, g2])
  assert m.thresholds == [g1, g2] and m.ensemble_dim == 'realization'
  assert metrics.GaussianRPS([g1]).thresholds == [g1]
  assert metrics.EnsembleBrierScore([g1], 'number').ensemble_dim == 'number'


def test_seeps_climatology_handling():
  import test_seeps_gpu as seeps_helpers
  from weatherbench2_b200 import _seeps, evaluation, metrics
  fds, tds, _, truth = seeps_helpers._pair(nday=4)  # pylint: disable=protected-access
  rs = np.random.RandomState(0)
  clim, frac, _ = seeps_helpers._climatology(truth, 0, 0, rs)  # pylint: disable=protected-access
  s = metrics.SpatialSEEPS(climatology=clim)
  np.testing.assert_allclose(np.asarray(s.p1.values), frac.mean(axis=(0, 1)),
                             rtol=1e-6)
  assert set(s.p1.dims) == {'latitude', 'longitude'}
  tsel = evaluation.select_truth_at_valid_time(tds, fds)
  wet = clim[seeps_helpers.NAME + '_seeps_threshold']
  maps = _seeps._threshold_maps(wet, fds[seeps_helpers.NAME])  # pylint: disable=protected-access
  assert set(maps) == {'dayofyear', 'hour'}
  dims, pos = maps['dayofyear']
  assert dims == ('init_time', 'lead_time')
  np.testing.assert_array_equal(pos[:, 0], np.arange(4))  # 1..4 January
  np.testing.assert_array_equal(maps['hour'][1], 0)
  tmaps = _seeps._threshold_maps(wet, tsel[seeps_helpers.NAME])  # pylint: disable=protected-access
  np.testing.assert_array_equal(tmaps['dayofyear'][1], pos)
  with pytest.raises(KeyError):
    _seeps._threshold_maps(wet.isel(hour=0), fds[seeps_helpers.NAME])  # pylint: disable=protected-access


def test_rank_histogram_binning_rules():
  from weatherbench2_b200 import metrics
  assert metrics.RankHistogram()._num_bins_actual(9) == 10  # pylint: disable=protected-access
  assert metrics.RankHistogram(num_bins=5)._num_bins_actual(9) == 5  # pylint: disable=protected-access
  with pytest.raises(ValueError, match='Cannot bin'):
    metrics.RankHistogram(num_bins=4)._num_bins_actual(9)  # pylint: disable=protected-access
  rh = metrics.RankHistogram(ensemble_dim='number', seed=3)
  assert rh.ensemble_dim == 'number' and rh._seed == 3  # pylint: disable=protected-access
  truth, forecast = td.get_random_truth_and_forecast(ensemble_size=None)
  with pytest.raises(ValueError):  # no ensemble dimension (metrics.py:574-577)
    rh.compute_chunk(thr_helpers._ds(**forecast), thr_helpers._ds(**truth))  # pylint: disable=protected-access


def test_time_mean_accumulator_with_map_and_quantile_outputs():
  """Chunked time means of map-output / threshold results (extra `quantile`,
  `bins`, latitude / longitude dims) equal the unchunked mean, NaN-aware."""
  from weatherbench2_b200 import distributed, xarray_lite as xl
  rs = np.random.RandomState(0)
  dims = ('quantile', 'init_time', 'lead_time', 'latitude', 'longitude')
  full = rs.normal(size=(3, 6, 2, 5, 8))
  full[rs.uniform(size=full.shape) < 0.1] = np.nan
  full[:, :, 0, 0, 0] = np.nan  # a cell that is NaN in every chunk
  coords = {'quantile': np.array([0.1, 0.5, 0.9]), 'init_time': np.arange(6),
            'lead_time': np.arange(2), 'latitude': np.linspace(-60, 60, 5),
            'longitude': np.arange(8) * 45.0}
  for skipna in (False, True):
    acc = distributed.TimeMeanAccumulator('init_time', skipna)
    for i0 in range(0, 6, 2):
      chunk = {k: (v[i0:i0 + 2] if k == 'init_time' else v)
               for k, v in coords.items()}
      acc.add(xl.Dataset({'z': (dims, full[:, i0:i0 + 2])}, chunk))
    got = acc.finish()['z']
    assert got.dims == ('quantile', 'lead_time', 'latitude', 'longitude')
    import warnings
    with warnings.catch_warnings():
      warnings.simplefilter('ignore', RuntimeWarning)
      want = np.nanmean(full, axis=1) if skipna else full.mean(axis=1)
    np.testing.assert_allclose(got.values, want, rtol=1e-12, equal_nan=True)
    assert np.isnan(got.values[:, 0, 0, 0]).all()


def test_upload_gathered_packs_only_referenced_slabs_and_keeps_members_plain():
  """_spatial._upload_gathered with a fake context: a (member, time) operand
  whose time axis is a label gather of a long record is packed to
  members x DISTINCT times slabs; the member axis stays evenly strided (what
  split_member_dim needs); every addressed slab holds the right data."""
  from weatherbench2_b200 import _spatial as sp, xarray_lite as xl

  class FakeCtx:
    def __init__(self):
      self.blobs = {}
    def to_device(self, a):
      a = np.ascontiguousarray(a)
      self.blobs[a.ctypes.data] = a  # "device" memory = this host copy
      return a.ctypes.data

  rs = np.random.RandomState(0)
  nm, nt, nlat, nlon = 3, 50, 5, 8
  x = rs.standard_normal((nm, nt, nlat, nlon)).astype(np.float32)
  da = xl.DataArray(x, ('realization', 'time', 'latitude', 'longitude'),
                    {'realization': np.arange(nm), 'time': np.arange(nt),
                     'latitude': np.linspace(-90, 90, nlat),
                     'longitude': np.arange(nlon) * 45.0})
  pos = np.array([[7, 9, 11], [9, 11, 13]])  # (init, lead) -> time, overlapping
  op = sp.gather_operand(sp.prepare_operand(da),
                         {'time': (('init', 'lead'), pos)})
  ctx, staged = FakeCtx(), []
  dev = sp._upload_gathered(ctx, op, staged)  # pylint: disable=protected-access
  blob = ctx.blobs[dev.addr]
  assert blob.shape[:2] == (nm, 4)  # 4 distinct times: 7, 9, 11, 13
  dims, shape = dev.outer_dims, dev.outer_shape
  table = sp.offset_table(dev, dims, shape).reshape(shape)
  flat = blob.reshape(-1)
  for m in range(nm):
    for i in range(2):
      for l in range(3):
        idx = {'realization': m, 'init': i, 'lead': l}
        off = table[tuple(idx[d] for d in dims)]
        np.testing.assert_array_equal(
            flat[off:off + nlat * nlon].reshape(nlat, nlon), x[m, pos[i, l]])
  rest, m_, stride = sp.split_member_dim(dev, 'realization')
  assert m_ == nm and stride == table[tuple(
      1 if d == 'realization' else 0 for d in dims)]
  assert 'realization' not in rest.outer_dims


@pytest.mark.parametrize('hist,expected,desired', [
    # weatherbench2/metrics_test.py:701-779 (known answers)
    ([0.2, 0.1, 0.7], [0.1, 1.0], [1 / 3, 1.0]),
    ([0.2, 0.0, 0.1, 0.1, 0.6], [0.1, 0.2, 1.0], [1 / 5, 3 / 5, 1]),
    ([0.1, 0.1, 0.5, 0.3], [0.6, 1.0], [1 / 2, 1.0]),
    ([0.1, 0.1, 0.3, 0.2, 0.0, 0.3], [0.5, 0.6, 1.0], [1 / 3, 2 / 3, 1]),
])
def test_central_reliability_known_answers(hist, expected, desired):
  from weatherbench2_b200 import metrics
  from weatherbench2_b200 import xarray_lite as xl
  ds = xl.Dataset({'temperature': (('bins',), np.array(hist))},
                  {'bins': np.arange(len(hist))})
  rel = metrics.central_reliability(ds)
  da = rel['temperature']
  assert da.dims == ('desired_prob',)
  np.testing.assert_allclose(da.values, expected, rtol=1e-12)
  np.testing.assert_allclose(da.coords['desired_prob'].values, desired,
                             rtol=1e-12)
  np.testing.assert_array_equal(da.coords['prob_index'].values,
                                np.arange(len(expected)))


@pytest.mark.parametrize('n_bins', [3, 4, 10, 11])
def test_central_reliability_of_a_calibrated_histogram(n_bins):
  """metrics_test.py:666-699: expected == desired probabilities; extra
  dimensions ride along; fewer than 3 bins is an error (:655-664)."""
  from weatherbench2_b200 import metrics
  from weatherbench2_b200 import xarray_lite as xl
  h = np.ones((2, n_bins)) / n_bins
  ds = xl.Dataset({'temperature': (('level', 'bins'), h)},
                  {'bins': np.arange(n_bins), 'level': [500, 850]})
  rel = metrics.central_reliability(ds)['temperature']
  assert rel.sizes['desired_prob'] == n_bins // 2 + n_bins % 2
  unnorm = np.ones(n_bins // 2)
  if n_bins % 2:
    unnorm = np.concatenate(([0.5], unnorm))
  want = np.cumsum(unnorm) / unnorm.sum()
  got = rel.transpose('level', 'desired_prob').values
  np.testing.assert_allclose(got, np.stack([want, want]), rtol=1e-12)
  np.testing.assert_allclose(rel.coords['desired_prob'].values, want,
                             rtol=1e-12)
  with pytest.raises(ValueError, match='Too few bins'):
    metrics.central_reliability(xl.Dataset(
        {'temperature': (('bins',), np.ones(2) / 2)}, {'bins': np.arange(2)}))


def test_land_mask_device_copies_are_cached_per_context():
  """The cached LandRegion mask is a DEVICE pointer: a second context (another
  device, or a context created after the first was closed) must upload its
  own copy, never reuse the first one's pointer."""
  import fake_ctx
  from weatherbench2_b200 import _spatial as sp, regions as R
  lat = np.linspace(-90, 90, 7)
  lon = np.linspace(0, 360, 12, endpoint=False)
  lsm = (np.random.RandomState(0).rand(7, 12) > 0.5).astype(float)
  shared = {}  # what metrics.py passes to ask for caching
  a, b = fake_ctx.FakeContext(), fake_ctx.FakeContext()
  ptrs = []
  for ctx in (a, b, a):
    (_, spec), = sp.build_weights(ctx, lat, lon, [R.LandRegion(lsm)],
                                  'lat_lon', 12, shared)
    assert spec.cell_w_dev in ctx._bufs  # owned by THIS context
    ptrs.append(spec.cell_w_dev)
  assert ptrs[0] != ptrs[1] and ptrs[0] == ptrs[2]  # one upload per context
  assert a.h2d_bytes == b.h2d_bytes == lsm.size * 4
  assert not shared  # the caller's dict is only a request flag


def test_land_mask_with_missing_values_is_refused():
  """xarray's `weighted()` raises on NaN weights (the mask is a factor of the
  weights, regions.py:138 / metrics.py:161); with a threshold the comparison
  turns NaN into 0 first (regions.py:136-137)."""
  from weatherbench2_b200 import regions as R
  lat = np.linspace(-90, 90, 5)
  lon = np.linspace(0, 360, 8, endpoint=False)
  lsm = np.ones((5, 8))
  lsm[2, 3] = np.nan
  with pytest.raises(ValueError, match='cannot contain missing values'):
    R.LandRegion(lsm).factors(lat, lon)
  cell = R.LandRegion(lsm, threshold=0.5).factors(lat, lon).cell
  assert cell[2, 3] == 0 and cell.sum() == 39
