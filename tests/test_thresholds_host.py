"""Host logic of `weatherbench2_b200.thresholds` (no GPU): `Threshold.compute`
reproduces the reference's selection (weatherbench2/thresholds.py:118-185) and
the kernels' offset tables address exactly the same climatology cells."""
import numpy as np
import pandas as pd
import pytest

from oracle import wb2_oracle as orc
import test_threshold_metrics_gpu as helpers
import wb2_testdata as td


def _level_clim(with_hour):
  """Climatology with (hour,) dayofyear, level, lat, lon and a level order that
  differs from the data's."""
  truth, _ = td.get_random_truth_and_forecast(
      variables=['geopotential'], time_stop='2019-12-03',
      time_resolution='6 hours', spatial_resolution_in_degrees=45)
  rs = np.random.RandomState(0)
  lat, lon = truth['coords']['latitude'], truth['coords']['longitude']
  levels = np.array([850, 500, 300, 700])  # data has (500, 700, 850)
  shape = (366, levels.size, lat.size, lon.size)
  dims = ('dayofyear', 'level', 'latitude', 'longitude')
  coords = {'dayofyear': np.arange(1, 367), 'level': levels, 'latitude': lat,
            'longitude': lon}
  if with_hour:
    shape = (4,) + shape
    dims = ('hour',) + dims
    coords['hour'] = np.array([0, 6, 12, 18])
  mean = rs.normal(size=shape).astype(np.float32)
  std = rs.uniform(0.5, 2, size=shape).astype(np.float32)
  quant = rs.normal(size=(3,) + shape).astype(np.float32)
  clim = helpers._ds(  # pylint: disable=protected-access
      {'geopotential': (dims, mean), 'geopotential_std': (dims, std),
       'geopotential_quantile': (('quantile',) + dims, quant)},
      dict(coords, quantile=np.array([0.1, 0.5, 0.9])))
  return truth, clim, mean, std, quant, levels


@pytest.mark.parametrize('with_hour', [False, True])
def test_threshold_compute_selects_like_the_reference(with_hour):
  from weatherbench2_b200 import thresholds
  truth, clim, mean, std, quant, levels = _level_clim(with_hour)
  tds = helpers._ds(**truth)  # pylint: disable=protected-access
  stamps = pd.DatetimeIndex(truth['coords']['time'])
  doy = stamps.dayofyear.values - 1
  hour = stamps.hour.values // 6
  lev = [list(levels).index(l) for l in truth['coords']['level']]
  tdims, _ = truth['vars']['geopotential']

  def select(a):  # .sel(level=truth.level).sel(dayofyear=..., hour=...)
    if with_hour:
      a = a[hour, doy]
    else:
      a = a[doy]
    return a[:, lev]  # dims (time, level, lat, lon)

  sel_dims = ('time', 'level', 'latitude', 'longitude')
  got = thresholds.GaussianQuantileThreshold(clim, 0.9).compute(tds)[
      'geopotential']
  want = orc.gaussian_quantile_threshold(select(mean), select(std), 0.9)
  a, b, _ = orc.align(got.values, got.dims, want, sel_dims)
  np.testing.assert_array_equal(a, b)
  got = thresholds.QuantileThreshold(clim, 0.5).compute(tds)['geopotential']
  a, b, _ = orc.align(got.values, got.dims, select(quant[1]), sel_dims)
  np.testing.assert_array_equal(a, b)
  # nearest quantile within 0.01, else KeyError (thresholds.py:79-89)
  got = thresholds.QuantileThreshold(clim, 0.505).compute(tds)['geopotential']
  np.testing.assert_array_equal(np.asarray(got.values).ravel().sum(),
                                a.ravel().sum())
  with pytest.raises(KeyError):
    thresholds.QuantileThreshold(clim, 0.7).compute(tds)


@pytest.mark.parametrize('with_hour', [False, True])
def test_kernel_offset_tables_address_the_same_cells(with_hour):
  """Read the climatology through the offset table the kernel gets and compare
  with `Threshold.compute`."""
  from weatherbench2_b200 import _spatial as sp, thresholds
  truth, clim, *_ = _level_clim(with_hour)
  tds = helpers._ds(**truth)  # pylint: disable=protected-access
  t_da = tds['geopotential']
  t_op = sp.prepare_operand(t_da, None, np.float32)
  dims, shape = sp.broadcast_dims(t_op)
  thr = thresholds.QuantileThreshold(clim, 0.9)
  kind, op = thr.kernel_operands(tds, 'geopotential', t_da, t_op.layout)
  assert kind == 'field'
  tab = sp.offset_table(op, dims, shape)
  flat = np.asarray(op.data).ravel()
  nrow, ncol = op.nrow, op.ncol
  got = np.stack([flat[o:o + nrow * op.row_stride].reshape(nrow, -1)[:, :ncol]
                  for o in tab]).reshape(shape + (nrow, ncol))
  want = thr.compute(tds)['geopotential']
  sp_dims = ('latitude', 'longitude') if op.layout == 'lat_lon' else (
      'longitude', 'latitude')
  a, b, _ = orc.align(got, tuple(dims) + sp_dims, np.asarray(want.values),
                      want.dims)
  np.testing.assert_array_equal(a, b)
  # Gaussian form: mean / std operands + z
  g = thresholds.GaussianQuantileThreshold(clim, 0.25)
  kind, m_op, s_op, z = g.kernel_operands(tds, 'geopotential', t_da,
                                          t_op.layout)
  assert kind == 'gaussian'
  from scipy import stats
  assert z == pytest.approx(stats.norm.ppf(0.25), abs=1e-15)
  tm = sp.offset_table(m_op, dims, shape)
  ts = sp.offset_table(s_op, dims, shape)
  fm, fs = np.asarray(m_op.data).ravel(), np.asarray(s_op.data).ravel()
  thr_cells = fm[tm] + np.float64(z) * fs[ts].astype(np.float64)
  want = g.compute(tds)['geopotential']
  first = np.asarray(want.values)
  a, b, _ = orc.align(first, want.dims,
                      np.zeros(shape + (nrow, ncol)), tuple(dims) + sp_dims)
  idx = (Ellipsis, 0, 0)
  np.testing.assert_array_equal(a[idx].ravel(), thr_cells)


def test_threshold_class_lookup_and_missing_variables():
  from weatherbench2_b200 import thresholds
  assert thresholds.get_threshold_cls('quantile') is (
      thresholds.QuantileThreshold)
  assert thresholds.get_threshold_cls('gaussian_quantile') is (
      thresholds.GaussianQuantileThreshold)
  with pytest.raises(NotImplementedError):
    thresholds.get_threshold_cls('other')
  truth, clim, *_ = _level_clim(False)
  tds = helpers._ds(**truth)  # pylint: disable=protected-access
  clim_no_std = clim[['geopotential']]
  with pytest.raises(KeyError):
    thresholds.GaussianQuantileThreshold(clim_no_std, 0.5).compute(tds)
