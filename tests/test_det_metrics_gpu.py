"""GPU parity tests for K1 (deterministic metrics) against the oracle.

All calls go through the C ABI (ctypes -> libwb2b200.so).  Tolerance: the
north-star bound is 1e-5 relative in fp32; these tests use 2e-6 where the
statistic is a sum of same-sign terms and an absolute floor where the statistic
is a signed sum that cancels (Bias).
"""
import numpy as np
import pytest

from oracle import wb2_oracle as orc
import wb2_testdata as td

pytestmark = pytest.mark.gpu

RTOL = 2e-6


@pytest.fixture(scope='module')
def ctx():
  from weatherbench2_b200 import _lib
  return _lib.default_context(0)


def _ds(vars_, coords):
  from weatherbench2_b200 import xarray_lite as xl
  return xl.Dataset({k: (d, v) for k, (d, v) in vars_.items()}, coords)


def _grid(nlat, nlon):
  return np.linspace(-90, 90, nlat), np.linspace(0, 360, nlon, endpoint=False)


def _regions_pair():
  """(oracle regions, product regions) with matching order."""
  from weatherbench2_b200 import regions as R
  o = [None,
       orc.SliceRegion(lat_slice=slice(-20, 20)),
       orc.ExtraTropicalRegion(),
       orc.SliceRegion(lat_slice=slice(20, 90), lon_slice=slice(0, 180)),
       orc.SliceRegion(lat_slice=slice(35, 75),
                       lon_slice=[slice(347.5, None), slice(0, 42.5)]),
       orc.SliceRegion(lat_slice=[slice(None, -60), slice(60, None)])]
  p = [None,
       R.SliceRegion(lat_slice=slice(-20, 20)),
       R.ExtraTropicalRegion(),
       R.SliceRegion(lat_slice=slice(20, 90), lon_slice=slice(0, 180)),
       R.SliceRegion(lat_slice=slice(35, 75),
                     lon_slice=[slice(347.5, None), slice(0, 42.5)]),
       R.SliceRegion(lat_slice=[slice(None, -60), slice(60, None)])]
  return o, p


# ------------------------------------------------------------------------------
# raw C-ABI level: device buffers, offset tables, all statistics at once
# ------------------------------------------------------------------------------
@pytest.mark.parametrize('layout', ['lat_lon', 'lon_lat'])
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('skipna', [False, True])
def test_raw_abi_all_stats(ctx, layout, dtype, skipna):
  from weatherbench2_b200 import _lib, _spatial as sp
  rs = np.random.RandomState(11)
  nlat, nlon, nb = 61, 120, 5
  lat, lon = _grid(nlat, nlon)
  shape = (nb, nlat, nlon) if layout == 'lat_lon' else (nb, nlon, nlat)
  dims = (('b', 'latitude', 'longitude') if layout == 'lat_lon'
          else ('b', 'longitude', 'latitude'))
  f = rs.normal(size=shape).astype(dtype)
  t = rs.normal(size=shape).astype(dtype)
  c = (0.3 * rs.normal(size=shape[1:])).astype(dtype)
  if skipna:
    f[rs.rand(*shape) < 0.01] = np.nan
    t[rs.rand(*shape) < 0.01] = np.nan
    c[rs.rand(*c.shape) < 0.01] = np.nan
  oreg, preg = _regions_pair()
  slab = nlat * nlon
  df, dt_, dc = ctx.to_device(f), ctx.to_device(t), ctx.to_device(c)
  base = min(df, dt_, dc)
  es = np.dtype(dtype).itemsize
  off_f = np.arange(nb, dtype=np.int64) * slab + (df - base) // es
  off_t = np.arange(nb, dtype=np.int64) * slab + (dt_ - base) // es
  off_c = np.zeros(nb, dtype=np.int64) + (dc - base) // es
  (ids, spec), = sp.build_weights(ctx, lat, lon, preg, layout,
                                  shape[-1])
  assert ids == list(range(len(preg)))
  out = ctx.malloc(nb * len(preg) * _lib.DET_NSTAT * 8)
  ctx.det_metrics(base, base, base, _lib.F32 if dtype == np.float32 else
                  _lib.F64, off_f, off_t, off_c, spec, skipna, out)
  st = ctx.from_device(out, (nb, len(preg), _lib.DET_NSTAT), np.float64)
  for p in (df, dt_, dc, out):
    ctx.free(p)
  cd = dims[1:]
  for ri, region in enumerate(oreg):
    kw = dict(lat=lat, lon=lon, region=region, skipna=skipna)
    want_mse, _ = orc.mse(f, dims, t, dims, **kw)
    want_mae, _ = orc.mae(f, dims, t, dims, **kw)
    want_bias, _ = orc.bias(f, dims, t, dims, **kw)
    want_acc, _ = orc.acc(f, dims, t, dims, c, cd, **kw)
    s = st[:, ri]
    with np.errstate(invalid='ignore', divide='ignore'):
      np.testing.assert_allclose(s[:, 0] / s[:, 6], want_mse, rtol=RTOL)
      np.testing.assert_allclose(s[:, 1] / s[:, 6], want_mae, rtol=RTOL)
      np.testing.assert_allclose(s[:, 2] / s[:, 6], want_bias, rtol=1e-4,
                                 atol=2e-6)
      acc = (s[:, 3] / s[:, 7]) / np.sqrt((s[:, 4] / s[:, 8]) *
                                          (s[:, 5] / s[:, 9]))
      np.testing.assert_allclose(acc, want_acc, rtol=1e-4, atol=2e-6)


def test_raw_abi_headline_shape_matches_oracle(ctx):
  """721 x 1440 x 13 levels, one variable, f32, global region."""
  from weatherbench2_b200 import _lib, _spatial as sp
  rs = np.random.RandomState(5)
  nlat, nlon, nlev = 721, 1440, 13
  lat, lon = _grid(nlat, nlon)
  f = rs.standard_normal((nlev, nlat, nlon)).astype(np.float32)
  t = rs.standard_normal((nlev, nlat, nlon)).astype(np.float32)
  c = rs.standard_normal((nlev, nlat, nlon)).astype(np.float32)
  dims = ('level', 'latitude', 'longitude')
  df, dt_, dc = ctx.to_device(f), ctx.to_device(t), ctx.to_device(c)
  base = min(df, dt_, dc)
  slab = nlat * nlon
  offs = [np.arange(nlev, dtype=np.int64) * slab + (p - base) // 4
          for p in (df, dt_, dc)]
  (_, spec), = sp.build_weights(ctx, lat, lon, [None], 'lat_lon', nlon)
  out = ctx.malloc(nlev * _lib.DET_NSTAT * 8)
  ctx.det_metrics(base, base, base, _lib.F32, offs[0], offs[1], offs[2], spec,
                  False, out)
  st = ctx.from_device(out, (nlev, _lib.DET_NSTAT), np.float64)
  # bit-stable run to run (fixed-order reduction)
  ctx.det_metrics(base, base, base, _lib.F32, offs[0], offs[1], offs[2], spec,
                  False, out)
  st2 = ctx.from_device(out, (nlev, _lib.DET_NSTAT), np.float64)
  np.testing.assert_array_equal(st, st2)
  for p in (df, dt_, dc, out):
    ctx.free(p)
  want_mse, _ = orc.mse(f, dims, t, dims, lat, lon)
  want_acc, _ = orc.acc(f, dims, t, dims, c, dims, lat, lon)
  np.testing.assert_allclose(st[:, 0] / st[:, 6], want_mse, rtol=RTOL)
  acc = (st[:, 3] / st[:, 7]) / np.sqrt((st[:, 4] / st[:, 8]) *
                                        (st[:, 5] / st[:, 9]))
  np.testing.assert_allclose(acc, want_acc, rtol=1e-4, atol=1e-6)
  np.testing.assert_allclose(st[:, 6], 721.0 * 1440, rtol=1e-9)


@pytest.mark.parametrize('skipna', [False, True])
@pytest.mark.parametrize('clim', [False, True])
def test_tma_and_ldg_paths_agree(ctx, monkeypatch, skipna, clim):
  """The TMA-staged persistent kernel (det_tma.cu) and the LDG kernel
  (det_metrics.cu) produce the same sums; both match the oracle.  Shapes are
  chosen so that CTAs own partial fields and partial rows (ragged tiles)."""
  from weatherbench2_b200 import _lib, _spatial as sp, regions as R
  rs = np.random.RandomState(21)
  nlat, nlon, nb = 181, 360, 7
  lat, lon = _grid(nlat, nlon)
  dims = ('b', 'latitude', 'longitude')
  f = rs.standard_normal((nb, nlat, nlon)).astype(np.float32)
  t = rs.standard_normal((nb, nlat, nlon)).astype(np.float32)
  c = rs.standard_normal((nb, nlat, nlon)).astype(np.float32)
  if skipna:
    f[rs.rand(*f.shape) < 0.003] = np.nan
    t[rs.rand(*f.shape) < 0.003] = np.nan
    c[rs.rand(*f.shape) < 0.003] = np.nan
  preg = [None, R.SliceRegion(lat_slice=slice(-20, 20)),
          R.ExtraTropicalRegion()]
  oreg = [None, orc.SliceRegion(lat_slice=slice(-20, 20)),
          orc.ExtraTropicalRegion()]
  df, dt_, dc = ctx.to_device(f), ctx.to_device(t), ctx.to_device(c)
  base = min(df, dt_, dc)
  slab = nlat * nlon
  offs = [np.arange(nb, dtype=np.int64) * slab + (p - base) // 4
          for p in (df, dt_, dc)]
  (_, spec), = sp.build_weights(ctx, lat, lon, preg, 'lat_lon', nlon)
  assert spec.nseg == 1
  out = ctx.malloc(nb * 3 * _lib.DET_NSTAT * 8)
  res = {}
  for path in ('tma', 'ldg'):
    monkeypatch.setenv('WB2_DET_PATH', path)
    ctx.det_metrics(base, base, base if clim else None, _lib.F32, offs[0],
                    offs[1], offs[2] if clim else None, spec, skipna, out)
    res[path] = ctx.from_device(out, (nb, 3, _lib.DET_NSTAT), np.float64)
  for p in (df, dt_, dc, out):
    ctx.free(p)
  np.testing.assert_allclose(res['tma'], res['ldg'], rtol=2e-6, atol=1e-4)
  for ri, region in enumerate(oreg):
    s = res['tma'][:, ri]
    want, _ = orc.mse(f, dims, t, dims, lat, lon, region=region, skipna=skipna)
    np.testing.assert_allclose(s[:, 0] / s[:, 6], want, rtol=RTOL)
    want, _ = orc.mae(f, dims, t, dims, lat, lon, region=region, skipna=skipna)
    np.testing.assert_allclose(s[:, 1] / s[:, 6], want, rtol=RTOL)
    if clim:
      want, _ = orc.acc(f, dims, t, dims, c, dims, lat, lon, region=region,
                        skipna=skipna)
      acc = (s[:, 3] / s[:, 7]) / np.sqrt((s[:, 4] / s[:, 8]) *
                                          (s[:, 5] / s[:, 9]))
      np.testing.assert_allclose(acc, want, rtol=1e-4, atol=2e-6)
    else:
      assert (s[:, 3:6] == 0).all() and (s[:, 7:] == 0).all()


@pytest.mark.parametrize('skipna', [False, True])
@pytest.mark.parametrize('clim', [False, True])
def test_segmented_tma_path_many_regions(ctx, monkeypatch, skipna, clim):
  """Longitude boxes cut the row into many column segments: the
  lane-contiguous TMA kernel (det_tma_seg.cu) against the LDG kernel and the
  oracle.  20 regions exercise the split into launches of <= 16 (8 with
  skipna) regions; NaN/Inf sit in cells some regions mask out."""
  from weatherbench2_b200 import _lib, _spatial as sp, regions as R
  rs = np.random.RandomState(33)
  nlat, nlon, nb = 91, 360, 6
  lat, lon = _grid(nlat, nlon)
  dims = ('b', 'latitude', 'longitude')
  f = rs.standard_normal((nb, nlat, nlon)).astype(np.float32)
  t = rs.standard_normal((nb, nlat, nlon)).astype(np.float32)
  c = rs.standard_normal((nb, nlat, nlon)).astype(np.float32)
  f[:, 80:, 200:230] = np.nan     # only inside some boxes
  t[:, :5, 10:20] = np.inf
  if skipna:
    f[rs.rand(*f.shape) < 0.003] = np.nan
  boxes = [(None, None, None, None), (-20, 20, None, None), (20, 90, 0, 180),
           (35, 75, 347.5, 42.5), (25, 60, 240, 290), (25, 60, 145, 180),
           (-45, -12.5, 120, 175), (-37.5, -22.5, 15, 50),
           (-52.5, -20, 287.5, 327.5), (-90, -60, None, None)]
  boxes = boxes + [(-30 + i, 40 - i, 13.0 * i, 13.0 * i + 77) for i in
                   range(10)]
  preg, oreg = [], []
  for la0, la1, lo0, lo1 in boxes:
    lat_s = slice(la0, la1)
    if lo0 is not None and lo1 is not None and lo0 > lo1:
      lon_p = [slice(lo0, None), slice(0, lo1)]
    else:
      lon_p = slice(lo0, lo1)
    preg.append(R.SliceRegion(lat_slice=lat_s, lon_slice=lon_p))
    oreg.append(orc.SliceRegion(lat_slice=lat_s, lon_slice=lon_p))
  preg.append(R.ExtraTropicalRegion())
  oreg.append(orc.ExtraTropicalRegion())
  nreg = len(preg)
  df, dt_, dc = ctx.to_device(f), ctx.to_device(t), ctx.to_device(c)
  base = min(df, dt_, dc)
  slab = nlat * nlon
  offs = [np.arange(nb, dtype=np.int64) * slab + (p - base) // 4
          for p in (df, dt_, dc)]
  (_, spec), = sp.build_weights(ctx, lat, lon, preg, 'lat_lon', nlon)
  assert spec.nseg > 8 and spec.nregion == nreg == 21
  out = ctx.malloc(nb * nreg * _lib.DET_NSTAT * 8)
  res = {}
  for path in ('tma', 'ldg'):
    monkeypatch.setenv('WB2_DET_PATH', path)
    ctx.det_metrics(base, base, base if clim else None, _lib.F32, offs[0],
                    offs[1], offs[2] if clim else None, spec, skipna, out)
    res[path] = ctx.from_device(out, (nb, nreg, _lib.DET_NSTAT), np.float64)
  for p in (df, dt_, dc, out):
    ctx.free(p)
  np.testing.assert_array_equal(np.isnan(res['tma']), np.isnan(res['ldg']))
  np.testing.assert_allclose(res['tma'], res['ldg'], rtol=3e-6, atol=2e-4)
  for ri, region in enumerate(oreg):
    s = res['tma'][:, ri]
    kw = dict(lat=lat, lon=lon, region=region, skipna=skipna)
    want, _ = orc.mse(f, dims, t, dims, **kw)
    with np.errstate(invalid='ignore', divide='ignore'):
      np.testing.assert_allclose(s[:, 0] / s[:, 6], want, rtol=3e-6)
      want, _ = orc.bias(f, dims, t, dims, **kw)
      np.testing.assert_allclose(s[:, 2] / s[:, 6], want, rtol=1e-4,
                                 atol=3e-6)
      if clim:
        want, _ = orc.acc(f, dims, t, dims, c, dims, **kw)
        acc = (s[:, 3] / s[:, 7]) / np.sqrt((s[:, 4] / s[:, 8]) *
                                            (s[:, 5] / s[:, 9]))
        np.testing.assert_allclose(acc, want, rtol=1e-4, atol=3e-6)


def test_unaligned_offsets_take_scalar_path(ctx):
  """Slabs that start at odd element offsets (no 16-byte alignment)."""
  from weatherbench2_b200 import _lib, _spatial as sp
  rs = np.random.RandomState(2)
  nlat, nlon = 19, 37
  lat, lon = _grid(nlat, nlon)
  buf_f = rs.normal(size=3 + 4 * nlat * nlon).astype(np.float32)
  buf_t = rs.normal(size=3 + 4 * nlat * nlon).astype(np.float32)
  off = np.array([1 + i * nlat * nlon for i in range(4)], dtype=np.int64)
  df, dt_ = ctx.to_device(buf_f), ctx.to_device(buf_t)
  base = min(df, dt_)
  (_, spec), = sp.build_weights(ctx, lat, lon, [None], 'lat_lon', nlon)
  out = ctx.malloc(4 * _lib.DET_NSTAT * 8)
  ctx.det_metrics(base, base, None, _lib.F32, off + (df - base) // 4,
                  off + (dt_ - base) // 4, None, spec, False, out)
  st = ctx.from_device(out, (4, _lib.DET_NSTAT), np.float64)
  for p in (df, dt_, out):
    ctx.free(p)
  for i in range(4):
    f = buf_f[off[i]:off[i] + nlat * nlon].reshape(nlat, nlon)
    t = buf_t[off[i]:off[i] + nlat * nlon].reshape(nlat, nlon)
    want, _ = orc.mse(f, ('latitude', 'longitude'), t,
                      ('latitude', 'longitude'), lat, lon)
    np.testing.assert_allclose(st[i, 0] / st[i, 6], want, rtol=RTOL)
  assert (st[:, 3:6] == 0).all() and (st[:, 7:] == 0).all()


# ------------------------------------------------------------------------------
# operator level: the reference's classes on mock datasets (host inputs)
# ------------------------------------------------------------------------------
def _mock_pair(**kw):
  truth, forecast = td.get_random_truth_and_forecast(**kw)
  return (_ds(forecast['vars'], forecast['coords']),
          _ds(truth['vars'], truth['coords']), forecast, truth)


@pytest.mark.parametrize('cast', [np.float32, np.float64])
def test_metric_classes_match_oracle(cast):
  from weatherbench2_b200 import metrics
  truth, forecast = td.get_random_truth_and_forecast(
      variables=['geopotential', 'temperature'])
  fv = {k: (d, v.astype(cast)) for k, (d, v) in forecast['vars'].items()}
  tv = {k: (d, v.astype(cast)) for k, (d, v) in truth['vars'].items()}
  fds, tds = _ds(fv, forecast['coords']), _ds(tv, truth['coords'])
  lat, lon = truth['coords']['latitude'], truth['coords']['longitude']
  for name in ('geopotential', 'temperature'):
    fd, f = fv[name]
    tdm, t = tv[name]
    for cls, fn, tol in [(metrics.MSE, orc.mse, dict(rtol=RTOL)),
                         (metrics.MAE, orc.mae, dict(rtol=RTOL)),
                         (metrics.RMSESqrtBeforeTimeAvg,
                          orc.rmse_sqrt_before_time_avg, dict(rtol=RTOL)),
                         (metrics.Bias, orc.bias, dict(rtol=1e-4, atol=2e-6))]:
      got = cls().compute_chunk(fds, tds)[name]
      want, wd = fn(f, fd, t, tdm, lat, lon)
      assert got.dims == wd
      np.testing.assert_allclose(got.values, want, **tol)
  # compute() = time mean (metrics.py:117-138)
  got = metrics.MSE().compute(fds, tds)['geopotential']
  want, wd = orc.mse(*fv['geopotential'][::-1], *tv['geopotential'][::-1],
                     lat, lon)
  want, wd = orc.time_mean(want, wd)
  assert got.dims == wd
  np.testing.assert_allclose(got.values, want, rtol=RTOL)


def test_regions_and_batch_mode():
  from weatherbench2_b200 import metrics
  truth, forecast = td.get_random_truth_and_forecast(
      spatial_resolution_in_degrees=5)
  fv = {k: (d, v.astype(np.float32)) for k, (d, v) in forecast['vars'].items()}
  tv = {k: (d, v.astype(np.float32)) for k, (d, v) in truth['vars'].items()}
  fds, tds = _ds(fv, forecast['coords']), _ds(tv, truth['coords'])
  lat, lon = truth['coords']['latitude'], truth['coords']['longitude']
  oreg, preg = _regions_pair()
  fd, f = fv['geopotential']
  tdm, t = tv['geopotential']
  singles = []
  for o, p in zip(oreg, preg):
    got = metrics.MSE().compute_chunk(fds, tds, region=p)['geopotential']
    want, _ = orc.mse(f, fd, t, tdm, lat, lon, region=o)
    np.testing.assert_allclose(got.values, want, rtol=RTOL)
    singles.append(got.values)
  with metrics.batch(preg):
    for p, s in zip(preg, singles):
      got = metrics.MSE().compute_chunk(fds, tds, region=p)['geopotential']
      # same cells, different f32 summation order (more column segments)
      np.testing.assert_allclose(got.values, s, rtol=1e-6)


def test_land_region_and_combined():
  from weatherbench2_b200 import metrics, regions as R
  truth, forecast = td.get_random_truth_and_forecast(
      spatial_resolution_in_degrees=10)
  fv = {k: (d, v.astype(np.float32)) for k, (d, v) in forecast['vars'].items()}
  tv = {k: (d, v.astype(np.float32)) for k, (d, v) in truth['vars'].items()}
  fds, tds = _ds(fv, forecast['coords']), _ds(tv, truth['coords'])
  lat, lon = truth['coords']['latitude'], truth['coords']['longitude']
  rs = np.random.RandomState(0)
  lsm = rs.rand(lat.size, lon.size)
  lsm[lsm < 0.3] = 0.0
  fd, f = fv['geopotential']
  tdm, t = tv['geopotential']
  for thr in (None, 0.5):
    got = metrics.MSE().compute_chunk(
        fds, tds, region=R.LandRegion(lsm, threshold=thr))['geopotential']
    want, _ = orc.mse(f, fd, t, tdm, lat, lon,
                      region=orc.LandRegion(lsm, threshold=thr))
    np.testing.assert_allclose(got.values, want, rtol=RTOL)
  comb_p = R.CombinedRegion([R.SliceRegion(lat_slice=slice(-30, 60)),
                             R.LandRegion(lsm)])
  comb_o = orc.CombinedRegion([orc.SliceRegion(lat_slice=slice(-30, 60)),
                               orc.LandRegion(lsm, latitude=lat,
                                              longitude=lon)])
  got = metrics.MAE().compute_chunk(fds, tds, region=comb_p)['geopotential']
  want, _ = orc.mae(f, fd, t, tdm, lat, lon, region=comb_o)
  np.testing.assert_allclose(got.values, want, rtol=RTOL)


@pytest.mark.parametrize('invalid_value', [np.inf, np.nan])
def test_rmse_over_invalid_region(invalid_value):
  """weatherbench2/metrics_test.py:133-152."""
  from weatherbench2_b200 import metrics, regions as R, xarray_lite as xl
  lat = np.array([-45.0, 0.0, 45.0])
  tv = np.array([0.0, invalid_value, 0.0]).reshape(1, 1, 3)
  coords = {'latitude': lat, 'longitude': np.array([0.0]),
            'time': np.array([0])}
  dims = ('time', 'longitude', 'latitude')
  truth = xl.Dataset({'wind_speed': (dims, tv)}, coords)
  forecast = xl.Dataset({'wind_speed': (dims, tv + 1)}, coords)
  rmse = metrics.RMSESqrtBeforeTimeAvg()
  assert np.isnan(rmse.compute(forecast, truth)['wind_speed'].values)
  got = rmse.compute(forecast, truth, region=R.ExtraTropicalRegion())
  np.testing.assert_allclose(got['wind_speed'].values, 1.0)


def test_wind_vector_rmse():
  """weatherbench2/metrics_test.py:84-131 -> per level [0, 10, nan]."""
  from weatherbench2_b200 import metrics
  kw = dict(variables_3d=['u_component_of_wind', 'v_component_of_wind'],
            variables_2d=[], time_start='2022-01-01', time_stop='2022-01-02')
  forecast = td.mock_forecast_data(lead_stop='0 day', **kw)
  truth = td.mock_truth_data(**kw)

  def lvl(x, dims, vals):
    shape = [1] * x.ndim
    shape[dims.index('level')] = 3
    return x + np.array(vals, dtype=np.float32).reshape(shape)

  fd, fu = forecast['vars']['u_component_of_wind']
  _, fv = forecast['vars']['v_component_of_wind']
  tdm, tu = truth['vars']['u_component_of_wind']
  _, tv = truth['vars']['v_component_of_wind']
  fds = _ds({'u_component_of_wind': (fd, lvl(fu, fd, [0, 3, np.nan])),
             'v_component_of_wind': (fd, lvl(fv, fd, [0, -4, 1]))},
            forecast['coords'])
  tds = _ds({'u_component_of_wind': (tdm, lvl(tu, tdm, [0, -3, np.nan])),
             'v_component_of_wind': (tdm, lvl(tv, tdm, [0, 4, 1]))},
            truth['coords'])
  wv = metrics.WindVectorRMSESqrtBeforeTimeAvg(
      u_name='u_component_of_wind', v_name='v_component_of_wind',
      vector_name='wind_vector')
  result = wv.compute(fds, tds).values.squeeze()
  np.testing.assert_allclose(result, np.array([0, 10, np.nan]))
  # and as part of RMSE (metrics.py:261-268)
  res = metrics.RMSESqrtBeforeTimeAvg(wind_vector_rmse=[wv]).compute(fds, tds)
  np.testing.assert_allclose(res['wind_vector'].values.squeeze(),
                             [0, 10, np.nan])


def test_acc_with_dayofyear_climatology():
  """ACC incl. the day-of-year / level lookup (metrics.py:387-414) and the
  `_mean` suffix rule (metrics_test.py:154-170)."""
  import pandas as pd
  from weatherbench2_b200 import metrics, xarray_lite as xl
  truth, forecast = td.get_random_truth_and_forecast(time_resolution='1 day',
                                                     time_stop='2019-12-04')
  fd, f = forecast['vars']['geopotential']
  tdm, t = truth['vars']['geopotential']
  f = f.astype(np.float32)
  t = t.astype(np.float32)
  lat, lon = truth['coords']['latitude'], truth['coords']['longitude']
  rs = np.random.RandomState(9)
  # climatology (dayofyear, level, lon, lat), levels stored in another order
  clim_levels = np.array([850, 500, 700])
  clim = rs.normal(size=(366, 3, lon.size, lat.size)).astype(np.float32)
  cdims = ('dayofyear', 'level', 'longitude', 'latitude')
  ccoords = {'dayofyear': 1 + np.arange(366), 'level': clim_levels,
             'latitude': lat, 'longitude': lon}
  fds = _ds({'geopotential': (fd, f)}, forecast['coords'])
  tds = _ds({'geopotential': (tdm, t)}, truth['coords'])
  for vname in ('geopotential', 'geopotential_mean'):
    cds = xl.Dataset({vname: (cdims, clim)}, ccoords)
    got = metrics.ACC(climatology=cds).compute_chunk(fds, tds)['geopotential']
    # oracle: select the climatology onto (time, level) by label, as xarray does
    doy = pd.DatetimeIndex(truth['coords']['time']).dayofyear.values
    lev_pos = [list(clim_levels).index(l) for l in truth['coords']['level']]
    csel = clim[doy - 1][:, lev_pos]  # (time, level, lon, lat)
    want, wd = orc.acc(f, fd, t, tdm, csel,
                       ('time', 'level', 'longitude', 'latitude'), lat, lon)
    a, b, _ = orc.align(got.values, got.dims, want, wd)
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=2e-6)
  with pytest.raises(KeyError):
    metrics.ACC(climatology=xl.Dataset({'other': (cdims, clim)}, ccoords)
                ).compute_chunk(fds, tds)


def test_shared_dims_are_joined_by_label():
  """`forecast - truth` aligns shared dimensions by coordinate label (xarray's
  inner join): by-valid forecasts whose `time` is a subset / reordering of the
  truth's.  The join is an offset-table gather, no data is copied."""
  from weatherbench2_b200 import metrics, xarray_lite as xl
  rs = np.random.RandomState(12)
  lat, lon = _grid(19, 36)
  t_times = np.arange(10)
  f_times = np.array([7, 2, 3, 11, 5])  # 11 is not in truth
  dims = ('time', 'level', 'latitude', 'longitude')
  f = rs.normal(size=(5, 2, 19, 36)).astype(np.float32)
  t = rs.normal(size=(10, 2, 19, 36)).astype(np.float32)
  coords = {'level': np.array([500, 850]), 'latitude': lat, 'longitude': lon}
  fds = xl.Dataset({'z': (dims, f)}, dict(coords, time=f_times))
  tds = xl.Dataset({'z': (dims, t)}, dict(coords, time=t_times))
  got = metrics.MSE().compute_chunk(fds, tds)['z']
  keep = [0, 1, 2, 4]
  want, wd = orc.mse(f[keep], dims, t[[7, 2, 3, 5]], dims, lat, lon)
  assert got.dims == wd and got.shape == (4, 2)
  np.testing.assert_array_equal(got.coords['time'].values, [7, 2, 3, 5])
  np.testing.assert_allclose(got.values, want, rtol=RTOL)
  # ensemble path too
  x = rs.normal(size=(3, 5, 2, 19, 36)).astype(np.float32)
  eds = xl.Dataset({'z': (('realization',) + dims, x)},
                   dict(coords, time=f_times, realization=np.arange(3)))
  got = metrics.CRPS().compute_chunk(eds, tds)['z']
  want, wd = orc.crps(x[:, keep], ('realization',) + dims, t[[7, 2, 3, 5]],
                      dims, 'realization', lat, lon)
  a, b, _ = orc.align(got.values, got.dims, want, wd)
  np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)


def test_non_increasing_latitude_raises():
  from weatherbench2_b200 import metrics, xarray_lite as xl
  lat = np.array([45.0, 0.0, -45.0])
  coords = {'latitude': lat, 'longitude': np.arange(4) * 90.0,
            'time': np.arange(2)}
  dims = ('time', 'latitude', 'longitude')
  x = np.zeros((2, 3, 4), np.float32)
  ds = xl.Dataset({'a': (dims, x)}, coords)
  with pytest.raises(ValueError):
    metrics.MSE().compute_chunk(ds, ds)


def test_device_resident_inputs_match_host_inputs():
  import torch
  from weatherbench2_b200 import metrics, xarray_lite as xl
  rs = np.random.RandomState(4)
  lat, lon = _grid(91, 180)
  dims = ('init_time', 'level', 'latitude', 'longitude')
  f = rs.normal(size=(3, 2, 91, 180)).astype(np.float32)
  t = rs.normal(size=(3, 2, 91, 180)).astype(np.float32)
  coords = {'init_time': np.arange(3), 'level': np.array([500, 850]),
            'latitude': lat, 'longitude': lon}
  host = metrics.MSE().compute_chunk(xl.Dataset({'z': (dims, f)}, coords),
                                     xl.Dataset({'z': (dims, t)}, coords))
  dev = metrics.MSE().compute_chunk(
      xl.Dataset({'z': (dims, torch.from_numpy(f).cuda())}, coords),
      xl.Dataset({'z': (dims, torch.from_numpy(t).cuda())}, coords))
  np.testing.assert_array_equal(host['z'].values, dev['z'].values)
  want, _ = orc.mse(f, dims, t, dims, lat, lon)
  np.testing.assert_allclose(dev['z'].values, want, rtol=RTOL)


def test_host_streaming_many_groups(monkeypatch):
  """wb2_det_metrics_host with a staging buffer that only fits a few slabs:
  exercises the double-buffered group loop and the slab de-duplication."""
  monkeypatch.setenv('WB2_STAGE_MB', '1')
  from weatherbench2_b200 import _lib, metrics, xarray_lite as xl
  ctx2 = _lib.Context(0)  # fresh context so the 1 MB staging size is used
  monkeypatch.setattr(metrics, '_context', lambda: ctx2)
  rs = np.random.RandomState(8)
  lat, lon = _grid(121, 240)
  fdims = ('lead_time', 'init_time', 'latitude', 'longitude')
  tdims = ('init_time', 'latitude', 'longitude')
  f = rs.normal(size=(7, 5, 121, 240)).astype(np.float32)
  t = rs.normal(size=(5, 121, 240)).astype(np.float32)  # broadcast over lead
  coords = {'lead_time': np.arange(7), 'init_time': np.arange(5),
            'latitude': lat, 'longitude': lon}
  got = metrics.MAE().compute_chunk(xl.Dataset({'z': (fdims, f)}, coords),
                                    xl.Dataset({'z': (tdims, t)}, coords))
  want, wd = orc.mae(f, fdims, t, tdims, lat, lon)
  assert got['z'].dims == wd
  np.testing.assert_allclose(got['z'].values, want, rtol=RTOL)
  ctx2.close()
