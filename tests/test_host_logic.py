"""CPU-only tests of the host logic around the kernels: offset tables, region
weights (bit-exact index sets vs the oracle), the named-array container and
the lazy by-init gather."""
import numpy as np
import pytest

from oracle import wb2_oracle as orc
from weatherbench2_b200 import _spatial as sp
from weatherbench2_b200 import regions as R
from weatherbench2_b200 import xarray_lite as xl


def _grid(nlat, nlon):
  return np.linspace(-90, 90, nlat), np.linspace(0, 360, nlon, endpoint=False)


def test_lat_weights_match_oracle_bitwise():
  for n in (7, 33, 721):
    lat = np.linspace(-90, 90, n)
    np.testing.assert_array_equal(sp.lat_weights(lat),
                                  orc.get_lat_weights(lat))
  with pytest.raises(ValueError):
    sp.lat_weights(np.array([10.0, 0.0, -10.0]))


def _dense_weights(spec, layout):
  """W[r, lat, lon] reconstructed from a WeightSpec (what the kernel applies)."""
  seg_of_col = np.zeros(spec.ncol, dtype=int)
  for k in range(spec.nseg):
    seg_of_col[spec.seg_start[k]:spec.seg_start[k + 1]] = k
  colw = np.ones(spec.ncol) if spec.col_w is None else spec.col_w.astype(
      np.float64)
  w = (spec.row_w[:, :, None] * spec.seg_w[:, seg_of_col][:, None, :] *
       colw[None, None, :])
  return w if layout == 'lat_lon' else np.transpose(w, (0, 2, 1))


PAIRS = [
    (None, None),
    (orc.SliceRegion(lat_slice=slice(-20, 20)),
     R.SliceRegion(lat_slice=slice(-20, 20))),
    (orc.ExtraTropicalRegion(), R.ExtraTropicalRegion()),
    (orc.SliceRegion(lat_slice=slice(35, 75),
                     lon_slice=[slice(347.5, None), slice(0, 42.5)]),
     R.SliceRegion(lat_slice=slice(35, 75),
                   lon_slice=[slice(347.5, None), slice(0, 42.5)])),
    (orc.SliceRegion(lat_slice=[slice(None, -60), slice(60, None)],
                     lon_slice=slice(100.1, 250.3)),
     R.SliceRegion(lat_slice=[slice(None, -60), slice(60, None)],
                   lon_slice=slice(100.1, 250.3))),
    # overlapping slices select some rows twice (no de-duplication)
    (orc.SliceRegion(lat_slice=[slice(-30, 10), slice(0, 40)]),
     R.SliceRegion(lat_slice=[slice(-30, 10), slice(0, 40)])),
    (orc.CombinedRegion([orc.SliceRegion(lon_slice=slice(0, 180)),
                         orc.ExtraTropicalRegion()]),
     R.CombinedRegion([R.SliceRegion(lon_slice=slice(0, 180)),
                       R.ExtraTropicalRegion()])),
]


@pytest.mark.parametrize('layout', ['lat_lon', 'lon_lat'])
@pytest.mark.parametrize('nlat,nlon', [(37, 72), (721, 1440)])
def test_region_weights_are_bit_exact_index_sets(layout, nlat, nlon):
  """The kernel's per-cell weight equals the oracle's region.apply weights
  scattered back onto the full grid -- index sets identical, values to 1 ulp
  (float32 rounding of col_w only in the lon_lat layout)."""
  lat, lon = _grid(nlat, nlon)
  oregs = [p[0] for p in PAIRS]
  pregs = [p[1] for p in PAIRS]
  (ids, spec), = sp.build_weights(None, lat, lon, pregs, layout,
                                  nlon if layout == 'lat_lon' else nlat)
  assert ids == list(range(len(PAIRS)))
  dense = _dense_weights(spec, layout)
  wfull = np.broadcast_to(orc.get_lat_weights(lat)[:, None], (nlat, nlon))
  marker = np.arange(nlat * nlon, dtype=np.float64).reshape(nlat, nlon)
  for ri, oreg in enumerate(oregs):
    expect = np.zeros((nlat, nlon))
    if oreg is None:
      expect = wfull.copy()
    else:
      cells, w, _, _ = orc._region_apply(oreg, marker, wfull, lat, lon)
      np.add.at(expect, (cells.astype(int) // nlon, cells.astype(int) % nlon),
                w)
    # identical support (bit-exact region-mask indexing)
    np.testing.assert_array_equal(dense[ri] != 0, expect != 0)
    tol = 0 if layout == 'lat_lon' else 1e-7
    np.testing.assert_allclose(dense[ri], expect, rtol=tol, atol=0)
  assert spec.zero_skip


def test_land_region_groups_by_mask():
  lat, lon = _grid(19, 36)
  rs = np.random.RandomState(0)
  lsm = rs.rand(19, 36)

  class FakeCtx:
    def __init__(self):
      self.uploaded = []

    def to_device(self, a):
      self.uploaded.append(np.array(a))
      return 4096 * len(self.uploaded)

  ctx = FakeCtx()
  regs = [None, R.LandRegion(lsm), R.LandRegion(lsm, threshold=0.5),
          R.CombinedRegion([R.SliceRegion(lat_slice=slice(0, 90)),
                            R.LandRegion(lsm)])]
  groups = sp.build_weights(ctx, lat, lon, regs, 'lon_lat', 19, {})
  got = sorted(tuple(g[0]) for g in groups)
  assert got == [(0,), (1, 3), (2,)]
  land = [g for g in groups if tuple(g[0]) == (1, 3)][0][1]
  assert land.cell_w_dev is not None and land.nregion == 2
  # the uploaded mask is transposed to the (lon, lat) slab layout
  assert any(u.shape == (36, 19) and np.allclose(u, lsm.T.astype(np.float32))
             for u in ctx.uploaded)


def test_offset_tables_address_the_right_slabs():
  rs = np.random.RandomState(1)
  f = rs.rand(3, 4, 2, 5, 6).astype(np.float32)   # lead, init, level, lon, lat
  t = rs.rand(4, 2, 5, 6).astype(np.float32)      # init, level, lon, lat
  fda = xl.DataArray(f, ('lead', 'init', 'level', 'longitude', 'latitude'))
  tda = xl.DataArray(t, ('init', 'level', 'longitude', 'latitude'))
  fo = sp.prepare_operand(fda)
  to = sp.prepare_operand(tda, fo.layout, fo.dtype)
  assert fo.layout == 'lon_lat' and (fo.nrow, fo.ncol) == (5, 6)
  dims, shape = sp.broadcast_dims(fo, to)
  assert dims == ('lead', 'init', 'level') and shape == (3, 4, 2)
  of = sp.offset_table(fo, dims, shape).reshape(shape)
  ot = sp.offset_table(to, dims, shape).reshape(shape)
  for idx in np.ndindex(*shape):
    np.testing.assert_array_equal(
        f.ravel()[of[idx]:of[idx] + 30].reshape(5, 6), f[idx])
    np.testing.assert_array_equal(
        t.ravel()[ot[idx]:ot[idx] + 30].reshape(5, 6), t[idx[1:]])


def test_prepare_operand_handles_views_and_copies():
  x = np.arange(2 * 7 * 8, dtype=np.float32).reshape(2, 7, 8)
  # (time, lat, lon) view with lon contiguous but a lat window: row stride 8
  v = xl.DataArray(x[:, 1:6, :], ('time', 'latitude', 'longitude'))
  op = sp.prepare_operand(v)
  assert op.layout == 'lat_lon' and op.row_stride == 8 and op.nrow == 5
  # spatial dims not innermost -> copied into (..., lat, lon)
  y = np.transpose(x, (1, 0, 2))
  op = sp.prepare_operand(xl.DataArray(y, ('latitude', 'time', 'longitude')))
  assert op.layout == 'lat_lon' and op.outer_dims == ('time',)
  assert op.row_stride == 8
  # integer data is promoted to float64, float16 to float32
  op = sp.prepare_operand(xl.DataArray(np.ones((3, 4), np.int32),
                                       ('latitude', 'longitude')))
  assert op.dtype == np.float64


def test_gather_operand_matches_numpy_take():
  rs = np.random.RandomState(3)
  clim = rs.rand(12, 3, 4, 5).astype(np.float32)  # doy, level, lat, lon
  da = xl.DataArray(clim, ('dayofyear', 'level', 'latitude', 'longitude'))
  op = sp.prepare_operand(da)
  pos = rs.randint(0, 12, size=(2, 6))
  lev = np.array([2, 0])
  g = sp.gather_operand(op, {'dayofyear': (('init', 'lead'), pos),
                             'level': (('level',), lev)})
  dims = ('lead', 'init', 'level')
  shape = (6, 2, 2)
  off = sp.offset_table(g, dims, shape).reshape(shape)
  for l in range(6):
    for i in range(2):
      for k in range(2):
        np.testing.assert_array_equal(
            clim.ravel()[off[l, i, k]:off[l, i, k] + 20].reshape(4, 5),
            clim[pos[i, l], lev[k]])
  with pytest.raises(IndexError):
    sp.gather_operand(op, {'dayofyear': (('a',), np.array([12]))})


# ---- container ---------------------------------------------------------------
def test_dataarray_broadcast_arithmetic_and_mean():
  a = xl.DataArray(np.arange(6.0).reshape(2, 3), ('x', 'y'),
                   {'x': [10, 20], 'y': [1, 2, 3]})
  b = xl.DataArray(np.array([1.0, 2.0, 3.0, 4.0]), ('z',), {'z': np.arange(4)})
  c = a - b
  assert c.dims == ('x', 'y', 'z') and c.shape == (2, 3, 4)
  np.testing.assert_allclose(c.values, a.values[:, :, None] - b.values)
  d = b * a  # dims of the first operand first
  assert d.dims == ('z', 'x', 'y')
  m = c.mean('z', skipna=False)
  assert m.dims == ('x', 'y')
  np.testing.assert_allclose(m.values, a.values - 2.5)
  s = np.sqrt(a)
  assert isinstance(s, xl.DataArray)
  np.testing.assert_allclose(s.values, np.sqrt(a.values))
  x = a.values.copy()
  x[0, 0] = np.nan
  an = xl.DataArray(x, ('x', 'y'))
  assert np.isnan(an.mean('y', skipna=False).values[0])
  np.testing.assert_allclose(an.mean('y', skipna=True).values[0], 1.5)


def test_sel_label_rules():
  lat = np.linspace(-90, 90, 37)
  a = xl.DataArray(np.arange(37.0), ('latitude',), {'latitude': lat})
  np.testing.assert_array_equal(a.sel(latitude=slice(-20, 20)).values,
                                np.arange(37.0)[(lat >= -20) & (lat <= 20)])
  assert float(a.sel(latitude=45.0).values) == 27.0
  with pytest.raises(KeyError):
    a.sel(latitude=44.0)
  np.testing.assert_array_equal(
      a.sel(latitude=np.array([0.0, 90.0])).values, [18.0, 36.0])


def test_dataset_ops_concat_merge():
  coords = {'lead': np.arange(3), 'level': np.array([500, 850])}
  ds = xl.Dataset({'a': (('lead', 'level'), np.ones((3, 2))),
                   'b': (('lead',), np.arange(3.0))}, coords)
  other = xl.Dataset({'a': (('lead', 'level'), 2 * np.ones((3, 2)))}, coords)
  diff = ds - other
  assert list(diff.keys()) == ['a']  # inner join on variables
  np.testing.assert_allclose(diff['a'].values, -1)
  m = ds.mean('lead')
  assert m['a'].dims == ('level',) and m['b'].dims == ()
  parts = []
  for name in ('global', 'tropics'):
    parts.append(ds.expand_dims({'metric': np.array(['mse'], dtype=object),
                                 'region': np.array([name], dtype=object)}))
  cat = xl.concat(parts, 'region')
  assert cat['a'].dims == ('metric', 'region', 'lead', 'level')
  assert list(cat['a'].coords['region'].values) == ['global', 'tropics']
  other_metric = xl.concat(
      [ds[['a']].expand_dims({'metric': np.array(['acc'], dtype=object),
                              'region': np.array([n], dtype=object)})
       for n in ('global', 'tropics')], 'region')
  merged = xl.merge([cat, other_metric])
  assert merged['a'].shape == (2, 2, 3, 2)
  assert list(merged['a'].coords['metric'].values) == ['mse', 'acc']
  assert merged['b'].shape == (2, 2, 3)
  assert np.isnan(merged['b'].values[1]).all()  # 'acc' has no variable b


def test_lazy_truth_gather_equals_materialised_sel():
  from weatherbench2_b200 import evaluation
  rs = np.random.RandomState(5)
  times = np.arange('2020-01-01', '2020-01-11', dtype='datetime64[D]').astype(
      'datetime64[ns]')
  lat, lon = _grid(5, 8)
  truth = xl.Dataset({'z': (('time', 'level', 'latitude', 'longitude'),
                            rs.rand(10, 2, 5, 8).astype(np.float32))},
                     {'time': times, 'level': np.array([500, 850]),
                      'latitude': lat, 'longitude': lon})
  init = times[:4]
  lead = (np.arange(3) * 24 * 3600 * 10**9).astype('timedelta64[ns]')
  fc = xl.Dataset({'z': (('time', 'prediction_timedelta', 'level', 'latitude',
                          'longitude'), rs.rand(4, 3, 2, 5, 8).astype(
                              np.float32))},
                  {'time': init, 'prediction_timedelta': lead,
                   'level': np.array([500, 850]), 'latitude': lat,
                   'longitude': lon})
  fc = evaluation.apply_time_conventions(fc, by_init=True)
  assert fc['z'].dims[:2] == ('init_time', 'lead_time')
  assert fc['valid_time'].dims == ('init_time', 'lead_time')
  lazy = evaluation.select_truth_at_valid_time(truth, fc)['z']
  assert lazy.dims == ('init_time', 'lead_time', 'level', 'latitude',
                       'longitude')
  # addressing through the offset table == materialised gather
  op = sp.prepare_operand(lazy)
  dims, shape = op.outer_dims, op.outer_shape
  off = sp.offset_table(op, dims, shape).reshape(shape)
  tv = truth['z'].values
  for i in range(4):
    for l in range(3):
      for k in range(2):
        np.testing.assert_array_equal(
            tv.ravel()[off[i, l, k]:off[i, l, k] + 40].reshape(5, 8),
            tv[i + l, k])
  np.testing.assert_array_equal(lazy.values[2, 1], tv[3])
