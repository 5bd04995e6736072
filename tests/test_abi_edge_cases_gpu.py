"""C-ABI edge cases on the GPU: empty launches, bad arguments (error codes and
messages instead of crashes), degenerate shapes, ragged sizes, all-NaN and
zero-weight inputs."""
import ctypes as C

import numpy as np
import pytest

from oracle import wb2_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
  from weatherbench2_b200 import _lib
  return _lib.default_context(0)


def _spec(ctx, nlat, nlon, regions=(None,), layout='lat_lon'):
  from weatherbench2_b200 import _spatial as sp
  lat = np.linspace(-90, 90, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  (_, spec), = sp.build_weights(ctx, lat, lon, list(regions), layout,
                                nlon if layout == 'lat_lon' else nlat)
  return spec, lat, lon


def test_empty_launch_is_a_noop(ctx):
  from weatherbench2_b200 import _lib
  spec, _, _ = _spec(ctx, 8, 16)
  empty = np.zeros(0, dtype=np.int64)
  out = ctx.malloc(64)
  before = ctx.launch_count
  ctx.det_metrics(out, out, None, _lib.F32, empty, empty, None, spec, False,
                  out)
  ctx.ens_metrics(out, out, _lib.F32, 3, 128, empty, empty, spec, False, out)
  ctx.energy_score(out, out, _lib.F32, 3, 128, empty, empty, spec, out)
  assert ctx.launch_count == before
  ctx.free(out)


def test_bad_arguments_return_errors(ctx):
  from weatherbench2_b200 import _lib
  spec, _, _ = _spec(ctx, 8, 16)
  off = np.zeros(1, dtype=np.int64)
  buf = ctx.malloc(8 * 16 * 4 * 4)
  with pytest.raises(_lib.Wb2Error, match='dtype'):
    ctx.det_metrics(buf, buf, None, 7, off, off, None, spec, False, buf)
  with pytest.raises(_lib.Wb2Error, match='NULL'):
    ctx.det_metrics(0, buf, None, _lib.F32, off, off, None, spec, False, buf)
  with pytest.raises(_lib.Wb2Error, match='at most'):
    ctx.ens_metrics(buf, buf, _lib.F32, 4000, 0, off, off, spec, False, buf)
  with pytest.raises(_lib.Wb2Error, match='64'):
    ctx.energy_score(buf, buf, _lib.F32, 65, 128, off, off, spec, buf)
  with pytest.raises(_lib.Wb2Error, match='F32'):
    ctx.ens_metrics(buf, buf, _lib.F64, 4, 128, off, off, spec, False, buf)
  # corrupt weight descriptors are rejected before any launch
  w = spec.as_struct()
  w.nregion = 40
  rc = ctx.lib.wb2_det_metrics(ctx.handle, buf, buf, None, _lib.F32, 1,
                               off.ctypes.data_as(C.POINTER(C.c_int64)),
                               off.ctypes.data_as(C.POINTER(C.c_int64)), None,
                               C.byref(w), 0, buf)
  assert rc == -1 and b'nregion' in ctx.lib.wb2_last_error()
  bad = spec.seg_start.copy()
  bad[-1] = 15
  w = spec.as_struct()
  w.seg_start = bad.ctypes.data_as(C.POINTER(C.c_int32))
  rc = ctx.lib.wb2_det_metrics(ctx.handle, buf, buf, None, _lib.F32, 1,
                               off.ctypes.data_as(C.POINTER(C.c_int64)),
                               off.ctypes.data_as(C.POINTER(C.c_int64)), None,
                               C.byref(w), 0, buf)
  assert rc == -1 and b'seg_start' in ctx.lib.wb2_last_error()
  # spectrum: odd / unfactorable longitude counts
  scale = np.ones(4)
  for ncol, code in ((9, -4), (14, -4), (1, -1)):
    rc = ctx.lib.wb2_zonal_spectrum(
        ctx.handle, buf, 1, 4, ncol,
        scale.ctypes.data_as(C.POINTER(C.c_double)), buf, 0, 1)
    assert rc == code, (ncol, rc, ctx.lib.wb2_last_error())
  # the library is still usable afterwards
  ctx.synchronize()
  ctx.free(buf)


@pytest.mark.parametrize('nlat,nlon', [(1, 4), (2, 1), (3, 5), (5, 33),
                                       (33, 128), (64, 132)])
def test_degenerate_and_ragged_grids(ctx, nlat, nlon):
  """1 x N, N x 1 and sizes that are not multiples of the vector width / warp
  size, against the oracle (uses the LDG path; 33 x 128 and 64 x 132 are
  TMA-eligible)."""
  from weatherbench2_b200 import metrics, xarray_lite as xl
  rs = np.random.RandomState(nlat * 100 + nlon)
  lat = np.linspace(-60, 60, nlat) if nlat > 1 else np.array([10.0])
  lon = np.linspace(0, 360, nlon, endpoint=False)
  dims = ('time', 'latitude', 'longitude')
  f = rs.normal(size=(3, nlat, nlon)).astype(np.float32)
  t = rs.normal(size=(3, nlat, nlon)).astype(np.float32)
  coords = {'time': np.arange(3), 'latitude': lat, 'longitude': lon}
  got = metrics.MSE().compute_chunk(xl.Dataset({'a': (dims, f)}, coords),
                                    xl.Dataset({'a': (dims, t)}, coords))['a']
  want, _ = orc.mse(f, dims, t, dims, lat, lon)
  np.testing.assert_allclose(got.values, want, rtol=2e-6)
  x = rs.normal(size=(4, 3, nlat, nlon)).astype(np.float32)
  got = metrics.CRPS().compute_chunk(
      xl.Dataset({'a': (('realization',) + dims, x)},
                 dict(coords, realization=np.arange(4))),
      xl.Dataset({'a': (dims, t)}, coords))['a']
  want, wd = orc.crps(x, ('realization',) + dims, t, dims, 'realization', lat,
                      lon)
  a, b, _ = orc.align(got.values, got.dims, want, wd)
  np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)


def test_all_nan_and_zero_weight_regions():
  """All-NaN field -> NaN (skipna False) and NaN (skipna True: 0 / 0 weights);
  a region with zero total weight -> NaN, like xarray's weighted mean."""
  from weatherbench2_b200 import metrics, regions as R, xarray_lite as xl
  lat = np.linspace(-90, 90, 13)
  lon = np.linspace(0, 360, 24, endpoint=False)
  dims = ('time', 'latitude', 'longitude')
  f = np.full((2, 13, 24), np.nan, np.float32)
  f[1] = 1.0
  t = np.zeros((2, 13, 24), np.float32)
  coords = {'time': np.arange(2), 'latitude': lat, 'longitude': lon}
  fds, tds = xl.Dataset({'a': (dims, f)}, coords), xl.Dataset({'a': (dims, t)},
                                                              coords)
  for skipna in (False, True):
    got = metrics.MSE().compute_chunk(fds, tds, skipna=skipna)['a'].values
    want, _ = orc.mse(f, dims, t, dims, lat, lon, skipna=skipna)
    np.testing.assert_allclose(got, want, equal_nan=True)
    assert np.isnan(got[0]) and got[1] == pytest.approx(1.0, rel=1e-6)
  # a latitude box that selects no row: weights sum to zero -> NaN
  empty = R.SliceRegion(lat_slice=slice(91, 95))
  got = metrics.MSE().compute_chunk(fds, tds, region=empty)['a'].values
  assert np.isnan(got).all()
  lsm = np.zeros((13, 24))
  got = metrics.MAE().compute_chunk(fds, tds, region=R.LandRegion(lsm))['a']
  assert np.isnan(got.values).all()


def test_threads_share_the_default_context():
  """Chunks evaluated from several threads against the process-wide context
  (Beam DirectRunner style, weatherbench2/evaluation.py:697, 733) give the
  same results as sequential calls."""
  import concurrent.futures
  from weatherbench2_b200 import metrics, regions as R, xarray_lite as xl
  lat = np.linspace(-90, 90, 19)
  lon = np.linspace(0, 360, 36, endpoint=False)
  dims = ('time', 'level', 'latitude', 'longitude')
  coords = {'time': np.arange(4), 'level': np.arange(3), 'latitude': lat,
            'longitude': lon}
  rs = np.random.RandomState(0)
  chunks = []
  for _ in range(16):
    f = rs.normal(size=(4, 3, 19, 36)).astype(np.float32)
    t = rs.normal(size=(4, 3, 19, 36)).astype(np.float32)
    x = rs.normal(size=(5, 4, 3, 19, 36)).astype(np.float32)
    chunks.append((xl.Dataset({'a': (dims, f)}, coords),
                   xl.Dataset({'a': (dims, t)}, coords),
                   xl.Dataset({'a': (('realization',) + dims, x)},
                              dict(coords, realization=np.arange(5)))))
  region = R.SliceRegion(lat_slice=slice(-30, 60))

  def work(chunk):
    f, t, x = chunk
    return (metrics.MSE().compute_chunk(f, t, region=region)['a'].values,
            metrics.CRPS().compute_chunk(x, t)['a'].values,
            np.asarray(metrics.SpatialMAE().compute(f, t)['a'].values))

  want = [work(c) for c in chunks]
  with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
    got = list(pool.map(work, chunks))
  for g, w in zip(got, want):
    for a, b in zip(g, w):
      np.testing.assert_array_equal(a, b)


def test_many_fields_and_more_than_32_regions():
  """7 800 fields in one launch (the size of a 10-init chunk at configs[1]) on a
  small grid, and 40 regions (split into launches of <= 32)."""
  from weatherbench2_b200 import metrics, regions as R, xarray_lite as xl
  rs = np.random.RandomState(5)
  lat = np.linspace(-90, 90, 9)
  lon = np.linspace(0, 360, 16, endpoint=False)
  dims = ('init_time', 'lead_time', 'level', 'latitude', 'longitude')
  f = rs.normal(size=(100, 6, 13, 9, 16)).astype(np.float32)
  t = rs.normal(size=(100, 6, 13, 9, 16)).astype(np.float32)
  coords = {'init_time': np.arange(100), 'lead_time': np.arange(6),
            'level': np.arange(13), 'latitude': lat, 'longitude': lon}
  fds, tds = xl.Dataset({'a': (dims, f)}, coords), xl.Dataset({'a': (dims, t)},
                                                              coords)
  got = metrics.MSE().compute_chunk(fds, tds)['a']
  want, _ = orc.mse(f, dims, t, dims, lat, lon)
  assert got.shape == (100, 6, 13)
  np.testing.assert_allclose(got.values, want, rtol=2e-6)
  regs = [R.SliceRegion(lat_slice=slice(-90 + 2 * i, 90 - 2 * i))
          for i in range(40)]
  oregs = [orc.SliceRegion(lat_slice=slice(-90 + 2 * i, 90 - 2 * i))
           for i in range(40)]
  with metrics.batch(regs):
    for r, o in list(zip(regs, oregs))[::7]:
      got = metrics.MAE().compute_chunk(fds, tds, region=r)['a']
      want, _ = orc.mae(f, dims, t, dims, lat, lon, region=o)
      np.testing.assert_allclose(got.values, want, rtol=2e-6)
