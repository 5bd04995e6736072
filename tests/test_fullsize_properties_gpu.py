"""Size-independent properties at BASELINE.json's full sizes (721 x 1440 x 13
levels, M = 50, 0.25 -> 1.5 degree regridding, 1440-point spectra), where the
oracle would take minutes: identities, linearity, conservation, Parseval,
run-to-run bit stability.  Inputs are generated and kept on the device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NLAT, NLON, NLEV = 721, 1440, 13


def _coords():
  return {'level': np.arange(NLEV), 'latitude': np.linspace(-90, 90, NLAT),
          'longitude': np.arange(NLON) * 0.25}


def _dev(shape, seed):
  import torch
  g = torch.Generator(device='cuda')
  g.manual_seed(seed)
  return torch.randn(shape, device='cuda', dtype=torch.float32, generator=g)


def test_k1_identities_and_linearity_fullsize():
  from weatherbench2_b200 import metrics, xarray_lite as xl
  dims = ('level', 'latitude', 'longitude')
  f = _dev((NLEV, NLAT, NLON), 1)
  t = _dev((NLEV, NLAT, NLON), 2)
  c = _dev((NLEV, NLAT, NLON), 3)

  def ds(x):
    return xl.Dataset({'z': (dims, x)}, _coords())

  mse = metrics.MSE().compute_chunk(ds(f), ds(t))['z'].values
  assert mse.shape == (NLEV,)
  # f, t ~ N(0, 1) independent: area-weighted MSE -> 2 within sampling noise
  np.testing.assert_allclose(mse, 2.0, rtol=5e-3)
  # identities
  np.testing.assert_array_equal(
      metrics.MSE().compute_chunk(ds(f), ds(f))['z'].values, 0.0)
  np.testing.assert_allclose(
      metrics.Bias().compute_chunk(ds(f + 3.0), ds(f))['z'].values, 3.0,
      rtol=1e-6)
  np.testing.assert_allclose(
      metrics.MAE().compute_chunk(ds(f + 3.0), ds(f))['z'].values, 3.0,
      rtol=1e-6)
  # homogeneity: MSE(a f, a t) = a^2 MSE(f, t); exact for a power of two
  np.testing.assert_array_equal(
      metrics.MSE().compute_chunk(ds(4 * f), ds(4 * t))['z'].values, 16 * mse)
  # symmetry in (f, t); RMSE^2 == MSE
  np.testing.assert_allclose(
      metrics.MSE().compute_chunk(ds(t), ds(f))['z'].values, mse, rtol=1e-12)
  rmse = metrics.RMSESqrtBeforeTimeAvg().compute_chunk(ds(f), ds(t))['z']
  np.testing.assert_allclose(rmse.values ** 2, mse, rtol=1e-12)
  # ACC: perfect forecast -> 1, anti-forecast -> -1, bounded otherwise
  clim = xl.Dataset({'z': (('dayofyear',) + dims, c[None])},
                    dict(_coords(), dayofyear=np.array([1])))
  coords_t = dict(_coords(), time=np.array(['2020-01-01'],
                                           dtype='datetime64[ns]'))
  tdims = ('time',) + dims
  ft = xl.Dataset({'z': (tdims, f[None])}, coords_t)
  tt = xl.Dataset({'z': (tdims, t[None])}, coords_t)
  acc = metrics.ACC(climatology=clim)
  np.testing.assert_allclose(acc.compute_chunk(ft, ft)['z'].values, 1.0,
                             rtol=1e-6)
  anti = xl.Dataset({'z': (tdims, (2 * c - f)[None])}, coords_t)
  np.testing.assert_allclose(acc.compute_chunk(anti, ft)['z'].values, -1.0,
                             rtol=1e-6)
  a = acc.compute_chunk(ft, tt)['z'].values
  assert (np.abs(a) < 1).all()
  # run-to-run bit stability (fixed-order reduction)
  np.testing.assert_array_equal(
      a, acc.compute_chunk(ft, tt)['z'].values)


def test_k2_identities_fullsize_m50():
  from weatherbench2_b200 import metrics, xarray_lite as xl
  m = 50
  dims = ('latitude', 'longitude')
  x = _dev((m, NLAT, NLON), 5)
  t = _dev((NLAT, NLON), 6)
  coords = {'latitude': np.linspace(-90, 90, NLAT),
            'longitude': np.arange(NLON) * 0.25, 'realization': np.arange(m)}
  fds = xl.Dataset({'z': (('realization',) + dims, x)}, coords)
  tds = xl.Dataset({'z': (dims, t)}, coords)
  skill = metrics.CRPSSkill().compute_chunk(fds, tds)['z'].values
  spread = metrics.CRPSSpread().compute_chunk(fds, tds)['z'].values
  crps = metrics.CRPS().compute_chunk(fds, tds)['z'].values
  np.testing.assert_allclose(crps, skill - 0.5 * spread, rtol=1e-12)
  # x, t iid N(0,1): E|X - Y| = E|X - X'| = 2 / sqrt(pi)
  np.testing.assert_allclose(skill, 2 / np.sqrt(np.pi), rtol=2e-3)
  np.testing.assert_allclose(spread, 2 / np.sqrt(np.pi), rtol=2e-3)
  var = metrics.EnsembleVariance().compute_chunk(fds, tds)['z'].values
  np.testing.assert_allclose(var, 1.0, rtol=2e-3)
  mse = metrics.EnsembleMeanMSE().compute_chunk(fds, tds)['z'].values
  np.testing.assert_allclose(mse, 1.0 + 1.0 / m, rtol=5e-3)
  deb = metrics.DebiasedEnsembleMeanMSE().compute_chunk(fds, tds)['z'].values
  np.testing.assert_allclose(deb, mse - var / m, rtol=1e-6)
  # spread is invariant to a permutation of the members and to a shift
  perm = np.random.RandomState(0).permutation(m)
  import torch
  fds_p = xl.Dataset({'z': (('realization',) + dims,
                            x[torch.from_numpy(perm).cuda()] + 7.0)}, coords)
  spread_p = metrics.CRPSSpread().compute_chunk(fds_p, tds)['z'].values
  np.testing.assert_allclose(spread_p, spread, rtol=2e-5)
  # identical members: spread 0, CRPS = MAE of that member
  same = xl.Dataset({'z': (('realization',) + dims,
                           x[:1].expand(m, NLAT, NLON).contiguous())}, coords)
  np.testing.assert_allclose(
      metrics.CRPSSpread().compute_chunk(same, tds)['z'].values, 0.0,
      atol=1e-7)
  one = xl.Dataset({'z': (dims, x[0])}, coords)
  np.testing.assert_allclose(
      metrics.CRPS().compute_chunk(same, tds)['z'].values,
      metrics.MAE().compute_chunk(one, tds)['z'].values, rtol=1e-5)


def test_k5_regrid_properties_fullsize():
  import torch
  from weatherbench2_b200 import regridding as rg
  from weatherbench2_b200 import _spatial as sp
  lon_s, lat_s = np.arange(NLON) * 0.25, np.linspace(-90, 90, NLAT)
  lon_t, lat_t = np.arange(240) * 1.5, np.linspace(-90, 90, 121)
  r = rg.ConservativeRegridder(rg.Grid.from_degrees(lon_s, lat_s),
                               rg.Grid.from_degrees(lon_t, lat_t))
  x = _dev((6, NLON, NLAT), 9)
  y = r.regrid_array(x)
  assert y.shape == (6, 240, 121) and y.is_cuda
  # constants are preserved; linear in the input
  ones = torch.full((1, NLON, NLAT), 2.5, device='cuda')
  np.testing.assert_allclose(r.regrid_array(ones).cpu().numpy(), 2.5,
                             rtol=1e-6)
  y2 = r.regrid_array(3.0 * x + 1.0)
  np.testing.assert_allclose(y2.cpu().numpy(), 3.0 * y.cpu().numpy() + 1.0,
                             rtol=1e-5, atol=1e-5)
  # conservation of the area-weighted global mean
  ws, wt = sp.lat_weights(lat_s), sp.lat_weights(lat_t)
  ms = (x.cpu().numpy().astype(np.float64) * ws).mean(axis=(1, 2))
  mt = (y.cpu().numpy().astype(np.float64) * wt).mean(axis=(1, 2))
  np.testing.assert_allclose(mt, ms, atol=2e-6)
  # each output lies within the range of its 7 x 7 source block (here: global)
  assert float(y.max()) <= float(x.max()) and float(y.min()) >= float(x.min())
  # NaN in one source cell touches exactly the targets whose stencil holds it
  xn = x[:1].clone()
  xn[0, 700, 300] = float('nan')
  yn = r.regrid_array(xn)
  assert not torch.isnan(yn).any()  # NaN-aware mean skips it
  assert (yn != y[:1]).sum().item() >= 1


def test_k4_spectrum_parseval_fullsize():
  from weatherbench2_b200 import derived_variables as dvs, xarray_lite as xl
  x = _dev((4, NLAT, NLON), 11) + 0.5
  lat = np.linspace(-90, 90, NLAT)
  lon = np.arange(NLON) * 0.25
  ds = xl.Dataset({'u': (('time', 'latitude', 'longitude'), x)},
                  {'time': np.arange(4), 'latitude': lat, 'longitude': lon})
  dv = dvs.ZonalEnergySpectrum('u')
  spec = dv.compute(ds)
  s = spec.values if isinstance(spec.values, np.ndarray) else spec.values
  s = np.asarray(s.cpu() if hasattr(s, 'cpu') else s, dtype=np.float64)
  assert s.shape == (4, NLAT, NLON // 2 + 1)
  xs = x.cpu().numpy().astype(np.float64)
  spacing = dv.lon_spacing_m(ds).values
  nyq = (xs[..., 0::2].sum(-1) - xs[..., 1::2].sum(-1)) / NLON
  circ = dv._circumference(lat)
  want = spacing[None, :] * (xs ** 2).sum(-1) + circ[None, :] * nyq ** 2
  # rows at the poles have ~0 circumference; compare where it is meaningful
  sel = np.abs(lat) < 89.9
  np.testing.assert_allclose(s.sum(-1)[:, sel], want[:, sel], rtol=2e-5)
  # k = 0 bin is the squared zonal mean times the circumference
  np.testing.assert_allclose(s[:, sel, 0],
                             (xs.mean(-1) ** 2 * circ[None, :])[:, sel],
                             rtol=1e-4)
  # time-summed spectrum == sum of per-time spectra
  tot = dv.compute(ds, time_sum_dim='time').values
  tot = np.asarray(tot.cpu() if hasattr(tot, 'cpu') else tot, np.float64)
  np.testing.assert_allclose(tot[sel], s.sum(0)[sel], rtol=1e-5)
