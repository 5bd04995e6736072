"""Vectors the REFERENCE's own metric classes produced in the build container
(tests/golden/make_reference_vectors.py: weatherbench2/metrics.py, regions.py,
thresholds.py executed on a re-implemented xarray subset, since xarray cannot
be installed there) against

  * the product operators on the NumPy stand-in context (same classes, same
    arguments as the reference call that made the vector), and
  * the oracle's restatement of the same call.

169 calls: every deterministic / ensemble / energy-score / spatial / Gaussian /
threshold metric, the region family, NaNs with and without skipna.  The CUDA
kernels face the same vectors in tests/test_zz_evaluation_cases_gpu.py."""
import os
import sys
import types
import warnings

import numpy as np
import pytest

import fake_ctx
from oracle import wb2_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import reference_cases as rc  # noqa: E402  pylint: disable=wrong-import-position

VECTORS = np.load(os.path.join(HERE, 'golden', 'reference_run_vectors.npz'))
BY_CASE = {}
for _key in VECTORS.files:
  if _key.startswith(('eval:', 'extra:')):  # other call shapes, see below
    continue
  _cid, _var, _dims = _key.split('|')
  BY_CASE.setdefault(_cid, {})[_var] = (
      tuple(d for d in _dims.split(',') if d), VECTORS[_key])


def product_lib():
  from weatherbench2_b200 import metrics, regions, thresholds
  return types.SimpleNamespace(metrics=metrics, regions=regions,
                               thresholds=thresholds)


def check_against_vectors(case, got, rtol=1e-5, atol=1e-6):
  want = BY_CASE[case['id']]
  assert set(got) == set(want), (sorted(got), sorted(want))
  for var, (dims, ref) in want.items():
    gd, gv = got[var]
    assert set(gd) == set(dims), (case['id'], var, gd, dims)
    gv = np.transpose(gv, [gd.index(d) for d in dims])
    np.testing.assert_array_equal(np.isnan(gv), np.isnan(ref),
                                  err_msg=f"{case['id']} {var}: NaN pattern")
    np.testing.assert_allclose(gv, ref, rtol=rtol, atol=atol, equal_nan=True,
                               err_msg=f"{case['id']} {var}")


def run_product(case, scope):
  from weatherbench2_b200 import xarray_lite as xl
  arr = rc.arrays()
  ds = rc.datasets(xl.Dataset, arr)
  with scope(), warnings.catch_warnings():
    warnings.simplefilter('ignore', RuntimeWarning)
    return rc.run_case(product_lib(), case, ds, arr, xl.DataArray)


def test_every_case_has_a_vector():
  assert set(BY_CASE) == {c['id'] for c in rc.CASES}
  assert len(rc.CASES) == 169


@pytest.mark.parametrize('case', rc.CASES, ids=lambda c: c['id'])
def test_product_operators_match_the_reference_run(case):
  check_against_vectors(case, run_product(case, fake_ctx.installed))


# ------------------------------------------------------------------------------
# The oracle, restating each call on plain arrays
# ------------------------------------------------------------------------------
EDIMS = ('realization',) + rc.DIMS


def _oracle_region(spec, arr):
  if spec is None:
    return None
  kind = spec['type']
  if kind == 'SliceRegion':
    return orc.SliceRegion(lat_slice=rc._slices(spec.get('lat')),  # pylint: disable=protected-access
                           lon_slice=rc._slices(spec.get('lon')))  # pylint: disable=protected-access
  if kind == 'ExtraTropicalRegion':
    return orc.ExtraTropicalRegion()
  if kind == 'LandRegion':
    return orc.LandRegion(arr[spec['mask']], threshold=spec.get('threshold'),
                          latitude=rc.LAT, longitude=rc.LON)
  return orc.CombinedRegion([_oracle_region(s, arr) for s in spec['regions']])


def _clim_at_times(a, lead_axes=0):
  """climatology (..., hour, dayofyear, level, lon, lat) -> (..., time, ...)."""
  import pandas as pd
  stamps = pd.DatetimeIndex(rc.TIMES)
  hour = np.asarray(stamps.hour) // 12
  doy = np.asarray(stamps.dayofyear) - 1
  idx = (slice(None),) * lead_axes + (hour, doy)
  return a[idx]


def _thresholds(kind, arr):
  """[(threshold array on (time, level, lon, lat))] per quantile."""
  if kind == 'QuantileThreshold':
    q = _clim_at_times(arr['clim/' + rc.Z + '_quantile'], lead_axes=1)
    return [q[k] for k in range(rc.QUANTILES.size)]
  mean = _clim_at_times(arr['clim/' + rc.Z])
  std = _clim_at_times(arr['clim/' + rc.Z + '_std'])
  return [orc.gaussian_quantile_threshold(mean, std, float(q))
          for q in rc.QUANTILES]


def run_oracle(case, arr):
  metric, kw = case['metric'], case.get('kwargs', {})
  region = _oracle_region(case.get('region'), arr)
  skipna = case.get('skipna', False)
  lat, lon = rc.LAT, rc.LON
  f_key, t_key = case['forecast'], case['truth']
  t_name = {'truth_z': 'truth'}.get(t_key, t_key)
  f_name = {'det_z': 'det'}.get(f_key, f_key)
  out = {}

  def sa(values, dims):
    return orc.spatial_average(values, dims, lat, lon, region, skipna)

  def finish(var, values, dims):
    if case.get('method') == 'compute':
      values, dims = orc.time_mean(values, dims, skipna=skipna)
    out[var] = (tuple(dims), values)

  simple = {'MSE': orc.mse, 'MAE': orc.mae, 'Bias': orc.bias,
            'RMSESqrtBeforeTimeAvg': orc.rmse_sqrt_before_time_avg}
  if metric in simple:
    variables = [v for v in (rc.Z, rc.U, rc.V) if f'{f_name}/{v}' in arr]
    for v in variables:
      r, d = simple[metric](arr[f'{f_name}/{v}'], rc.DIMS,
                            arr[f'{t_name}/{v}'], rc.DIMS, lat, lon, region,
                            skipna)
      finish(v, r, d)
    if 'wind_vector' in kw:
      r, d = orc.wind_vector_mse(
          arr[f'{f_name}/{rc.U}'], arr[f'{f_name}/{rc.V}'], rc.DIMS,
          arr[f'{t_name}/{rc.U}'], arr[f'{t_name}/{rc.V}'], rc.DIMS, lat, lon,
          region, skipna)
      finish('wind_vector', np.sqrt(r) if metric != 'MSE' else r, d)
    return out
  f = arr[f'{f_name}/{rc.Z}']
  t = arr[f'{t_name}/{rc.Z}']
  if metric == 'ACC':
    r, d = orc.acc(f, rc.DIMS, t, rc.DIMS, _clim_at_times(arr['clim/' + rc.Z]),
                   rc.DIMS, lat, lon, region, skipna)
    finish(rc.Z, r, d)
  elif metric in ('SpatialMSE', 'SpatialMAE', 'SpatialBias'):
    for v in (rc.Z, rc.U, rc.V):
      if f'{f_name}/{v}' in arr:
        r, d = orc.spatial_det_map(metric[7:].lower(), arr[f'{f_name}/{v}'],
                                   rc.DIMS, arr[f'{t_name}/{v}'], rc.DIMS)
        finish(v, r, d)
  elif metric in ('CRPS', 'CRPSSkill', 'EnsembleMeanMSE',
                  'EnsembleMeanRMSESqrtBeforeTimeAvg',
                  'DebiasedEnsembleMeanMSE', 'EnergyScore',
                  'EnergyScoreSkill'):
    fn = {'CRPS': orc.crps, 'CRPSSkill': orc.crps_skill,
          'EnsembleMeanMSE': orc.ensemble_mean_mse,
          'EnsembleMeanRMSESqrtBeforeTimeAvg':
              orc.ensemble_mean_rmse_sqrt_before_time_avg,
          'DebiasedEnsembleMeanMSE': orc.debiased_ensemble_mean_mse,
          'EnergyScore': orc.energy_score,
          'EnergyScoreSkill': orc.energy_score_skill}[metric]
    r, d = fn(f, EDIMS, t, rc.DIMS, 'realization', lat, lon, region=region,
              skipna=skipna)
    finish(rc.Z, r, d)
  elif metric in ('CRPSSpread', 'EnsembleVariance',
                  'EnsembleStddevSqrtBeforeTimeAvg', 'EnergyScoreSpread'):
    fn = {'CRPSSpread': orc.crps_spread,
          'EnsembleVariance': orc.ensemble_variance,
          'EnsembleStddevSqrtBeforeTimeAvg':
              orc.ensemble_stddev_sqrt_before_time_avg,
          'EnergyScoreSpread': orc.energy_score_spread}[metric]
    r, d = fn(f, EDIMS, 'realization', lat, lon, region=region, skipna=skipna)
    finish(rc.Z, r, d)
  elif metric in ('SpatialCRPS', 'SpatialCRPSSkill', 'SpatialCRPSSpread',
                  'SpatialEnsembleVariance', 'SpatialEnsembleMeanMSE',
                  'DebiasedSpatialEnsembleMeanMSE'):
    key = {'SpatialCRPS': 'crps', 'SpatialCRPSSkill': 'skill',
           'SpatialCRPSSpread': 'spread', 'SpatialEnsembleVariance': 'variance',
           'SpatialEnsembleMeanMSE': 'mse',
           'DebiasedSpatialEnsembleMeanMSE': 'debiased'}[metric]
    r, d = orc.spatial_ens_maps(f, EDIMS, t, rc.DIMS, 'realization',
                                skipna)[key]
    finish(rc.Z, r, d)
  elif metric in ('GaussianCRPS', 'GaussianVariance'):
    s = arr['gauss/' + rc.Z + '_std']
    point = (orc.gaussian_crps_pointwise(f, s, t) if metric == 'GaussianCRPS'
             else s.astype(np.float64) ** 2)
    finish(rc.Z, *sa(point, rc.DIMS))
  elif metric.startswith('Gaussian'):
    s = arr['gauss/' + rc.Z + '_std']
    fn = {'GaussianBrierScore': orc.gaussian_brier_pointwise,
          'GaussianIgnoranceScore': orc.gaussian_ignorance_pointwise,
          'GaussianRPS': orc.gaussian_rps_part_pointwise}[metric]
    parts = [sa(fn(f, s, t, thr), rc.DIMS)
             for thr in _thresholds(kw['thresholds'], arr)]
    _stack_thresholds(metric, parts, finish)
  else:  # ensemble threshold metrics, spatially averaged or maps
    spatial = metric.startswith('Spatial')
    name = metric.replace('Spatial', '')
    parts = []
    for thr in _thresholds(kw['thresholds'], arr):
      if 'Brier' in name:
        p = orc.ens_brier_pointwise(f, t, thr, 0, 'Debiased' in name, skipna)
      elif 'Ignorance' in name:
        p = orc.ens_ignorance_pointwise(f, t, thr, 0, skipna)
      else:
        p = orc.ens_rps_part_pointwise(f, t, thr, 0, skipna)
      parts.append((p, rc.DIMS) if spatial else sa(p, rc.DIMS))
    _stack_thresholds(name, parts, finish)
  return out


def _stack_thresholds(name, parts, finish):
  if name.endswith('RPS'):  # `.sum('quantile')`: xarray skips NaN by default
    finish(rc.Z, np.nansum(np.stack([p for p, _ in parts]), axis=0),
           parts[0][1])
  else:
    finish(rc.Z, np.stack([p for p, _ in parts]),
           ('quantile',) + tuple(parts[0][1]))


@pytest.mark.parametrize('case', rc.CASES, ids=lambda c: c['id'])
def test_oracle_matches_the_reference_run(case):
  with warnings.catch_warnings():
    warnings.simplefilter('ignore', RuntimeWarning)
    got = run_oracle(case, rc.arrays())
  check_against_vectors(case, got, rtol=2e-6, atol=1e-7)


# ------------------------------------------------------------------------------
# evaluate_in_memory: the result files the reference's own evaluation.py wrote
# ------------------------------------------------------------------------------
import reference_eval_cases as rec  # noqa: E402  pylint: disable=wrong-import-position

EVAL_VECTORS = {}
for _key in VECTORS.files:
  if _key.startswith('eval:'):
    _cid, _var, _dims = _key.split('|')
    EVAL_VECTORS.setdefault(_cid[5:], {})[_var] = (tuple(_dims.split(',')),
                                                   VECTORS[_key])


def check_product_evaluations(scope, tmp_path, rtol=1e-5, atol=1e-6):
  from weatherbench2_b200 import (config, evaluation, metrics, regions,
                                  xarray_lite as xl)
  lib = types.SimpleNamespace(config=config, metrics=metrics, regions=regions,
                              Dataset=xl.Dataset)
  seen = set()
  for case, data_config, eval_configs in rec.build(
      lib, lambda name, dataset: dataset, str(tmp_path)):
    with scope(), warnings.catch_warnings():
      warnings.simplefilter('ignore', RuntimeWarning)
      out = evaluation.evaluate_in_memory(data_config, eval_configs)
    for eval_name, res in out.items():
      key = f'{case}/{eval_name}'
      seen.add(key)
      want = EVAL_VECTORS[key]
      assert set(res.keys()) == set(want), key
      for var, (dims, ref) in want.items():
        da = res[var]
        assert set(da.dims) == set(dims), (key, da.dims, dims)
        got = np.asarray(da.transpose(*dims).values, dtype=np.float64)
        np.testing.assert_array_equal(np.isnan(got), np.isnan(ref),
                                      err_msg=key)
        np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol,
                                   equal_nan=True, err_msg=key)
  assert seen == set(EVAL_VECTORS) and len(seen) == 11


def test_evaluate_in_memory_matches_the_reference_run(tmp_path):
  """All eval configs of reference_eval_cases.py: plain / by region /
  temporal_mean=False / climatological, probabilistic-climatological and
  persistence forecasts / analysis as truth (by-init and by-valid) / latitude,
  longitude, level and time selection / pressure-level suffixes / by-valid step
  thinning -- product on the stand-in context vs the reference's result files."""
  check_product_evaluations(fake_ctx.installed, tmp_path)


# ------------------------------------------------------------------------------
# Rank histogram + central reliability, SEEPS, zonal spectrum + interpolation,
# wind speed
# ------------------------------------------------------------------------------
EXTRA_VECTORS = {}
for _key in VECTORS.files:
  if _key.startswith('extra:'):
    _name, _, _dims = _key.split('|')
    EXTRA_VECTORS[_name[6:]] = (tuple(d for d in _dims.split(',') if d),
                                VECTORS[_key])

EXTRA_TOL = {  # relative to the largest magnitude of the vector
    'spectrum': 1e-5, 'spectrum_interp': 2e-5, 'seeps': 1e-5,
    'spatial_seeps': 1e-5}


def check_product_extras(scope):
  from weatherbench2_b200 import (derived_variables, metrics, regions,
                                  regridding, xarray_lite as xl)
  lib = types.SimpleNamespace(metrics=metrics,
                              derived_variables=derived_variables,
                              regridding=regridding, regions=regions)
  with scope(), warnings.catch_warnings():
    warnings.simplefilter('ignore', RuntimeWarning)
    got = rc.run_extras(lib, xl.Dataset, rc.arrays())
  assert set(got) == set(EXTRA_VECTORS) and len(got) == 58
  for name, (dims, ref) in EXTRA_VECTORS.items():
    gd, gv = got[name]
    assert set(gd) == set(dims), (name, gd, dims)
    gv = np.transpose(np.asarray(gv, dtype=np.float64),
                      [gd.index(d) for d in dims])
    np.testing.assert_array_equal(np.isnan(gv), np.isnan(ref), err_msg=name)
    if name.startswith('rank_histogram'):
      np.testing.assert_array_equal(gv, ref, err_msg=name)
      continue
    finite = np.isfinite(ref)
    np.testing.assert_array_equal(gv[~finite & ~np.isnan(ref)],
                                  ref[~finite & ~np.isnan(ref)], err_msg=name)
    tol = EXTRA_TOL.get(name.split('/')[0], 1e-6)
    if name.startswith('region/'):  # index sets: exact
      assert gv.shape == ref.shape, (name, gv.shape, ref.shape)
      np.testing.assert_array_equal(gv, ref, err_msg=name)
      continue
    if name.startswith('regrid'):
      # float32 contraction (the reference: JAX einsum; here NumPy stands in)
      np.testing.assert_allclose(gv[finite], ref[finite], rtol=1e-5, atol=2e-6,
                                 err_msg=name)
    elif '/' in name or name == 'wind_speed':  # coordinates, sqrt(u^2 + v^2)
      np.testing.assert_allclose(gv[finite], ref[finite], rtol=1e-6, atol=0,
                                 err_msg=name)
    elif name.startswith('spectrum'):
      # per bin, relative to the row's total power (the bound the kernel tests
      # use: a float32 FFT cannot do better on the small bins)
      ax = dims.index('zonal_wavenumber' if name == 'spectrum' else 'frequency')
      power = np.nansum(np.abs(ref), axis=ax, keepdims=True)
      err = np.abs(np.where(finite, gv - ref, 0.0)) / np.where(power > 0, power,
                                                               1.0)
      assert err.max() <= tol, (name, err.max())
    else:
      scale = np.abs(ref[finite]).max() if finite.any() else 1.0
      assert np.abs(gv[finite] - ref[finite]).max() <= tol * scale, name


def test_product_extras_match_the_reference_run():
  """RankHistogram, central_reliability, SEEPS / SpatialSEEPS,
  ZonalEnergySpectrum (+ its frequency / wavelength coordinates),
  interpolate_spectral_frequencies and WindSpeed as the reference's own code
  computed them."""
  check_product_extras(fake_ctx.installed)


def test_oracle_extras_match_the_reference_run():
  x = rc.extra_inputs()
  arr = rc.arrays()
  f, t = arr['ens5/' + rc.Z], arr['truth/' + rc.Z]
  for bins in (None, 3):
    want_dims, want = EXTRA_VECTORS[f'rank_histogram/bins={bins}']
    got, gd = orc.rank_histogram_one_hot(f, EDIMS, t, rc.DIMS, 'realization',
                                         num_bins=bins)
    np.testing.assert_array_equal(
        np.transpose(got, [gd.index(d) for d in want_dims]), want)
  # SEEPS
  import pandas as pd
  vt = x['init'][:, None] + x['lead'][None, :]
  stamps = pd.DatetimeIndex(vt.ravel())
  hour = (np.asarray(stamps.hour) // 12).reshape(vt.shape)
  doy = (np.asarray(stamps.dayofyear) - 1).reshape(vt.shape)
  wet = x['thr'][hour, doy]
  p1 = x['dry'].mean(axis=(0, 1))
  dims = ('init_time', 'lead_time', 'longitude', 'latitude')
  point = orc.seeps_pointwise(x['pf'], x['pt'], wet, wet, p1)
  got, gd = orc.spatial_average(point, dims, rc.LAT, rc.LON, None, True)
  wd, want = EXTRA_VECTORS['seeps']
  np.testing.assert_allclose(np.transpose(got, [gd.index(d) for d in wd]), want,
                             rtol=2e-6)
  point = orc.seeps_pointwise(x['pf'], x['pt'], wet, wet, p1, min_p1=0.3,
                              max_p1=0.7)
  wd, want = EXTRA_VECTORS['spatial_seeps']
  np.testing.assert_allclose(np.transpose(point, [dims.index(d) for d in wd]),
                             want, rtol=2e-6, equal_nan=True)
  # zonal energy spectrum: values, frequency, wavelength
  lon36 = np.linspace(0, 360, 36, endpoint=False)
  sdims = ('time', 'level', 'latitude', 'longitude')
  spec, sd, freq, wavelength = orc.zonal_energy_spectrum(
      x['field'], sdims, rc.LAT, lon36)
  wd, want = EXTRA_VECTORS['spectrum']
  got = np.transpose(spec, [sd.index(d) for d in wd])
  assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
  wd, want = EXTRA_VECTORS['spectrum/frequency']
  got = freq if wd[0] == 'zonal_wavenumber' else freq.T
  finite = np.isfinite(want)
  np.testing.assert_array_equal(np.isfinite(got), finite)
  np.testing.assert_allclose(got[finite], want[finite], rtol=1e-12)
  wd, want = EXTRA_VECTORS['spectrum/wavelength']
  got = wavelength if wd[0] == 'zonal_wavenumber' else wavelength.T
  finite = np.isfinite(want)
  np.testing.assert_allclose(got[finite], want[finite], rtol=1e-12)
  fu, fv = x['field'], x['field'][::-1]
  np.testing.assert_array_equal(np.sqrt(fu**2 + fv**2),
                                EXTRA_VECTORS['wind_speed'][1])


def test_oracle_regridders_match_the_reference_run():
  """weatherbench2/regridding.py's own ConservativeRegridder / Bilinear /
  Nearest (run on a NumPy stand-in for jax.numpy) against the oracle."""
  rs = np.random.RandomState(31337)
  slon = np.linspace(0, 360, 24, endpoint=False)
  slat = np.linspace(-90, 90, 13)
  x = rs.standard_normal((3, 24, 13)).astype(np.float32)
  x[1, 3:6, 4:7] = np.nan
  x[2, :2, :] = np.nan
  tlon = np.linspace(0, 360, 10, endpoint=False)
  tlat = np.linspace(-90, 90, 7)
  grids = {
      'global': (orc.Grid(slon, slat), orc.Grid(tlon, tlat), x),
      'no_poles': (orc.Grid(slon, slat[1:-1], includes_poles=False),
                   orc.Grid(tlon, tlat), x[:, :, 1:-1]),
      'limited_area': (orc.Grid(slon[4:16], slat[2:10], periodic=False,
                                includes_poles=False),
                       orc.Grid(tlon[1:6], tlat[1:5], periodic=False,
                                includes_poles=False), x[:, 4:16, 2:10]),
  }
  fns = {'ConservativeRegridder': orc.conservative_regrid,
         'BilinearRegridder': orc.bilinear_regrid,
         'NearestRegridder': orc.nearest_regrid}
  for gname, (source, target, field) in grids.items():
    for cls, fn in fns.items():
      _, want = EXTRA_VECTORS[f'regrid/{cls}/{gname}']
      got = fn(field, source, target)
      np.testing.assert_array_equal(np.isnan(got), np.isnan(want),
                                    err_msg=f'{cls}/{gname}')
      np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6,
                                 equal_nan=True, err_msg=f'{cls}/{gname}')


def test_oracle_region_index_sets_match_the_reference_run():
  """regions.py's SliceRegion.apply, run by the reference itself with pandas'
  own `Index.slice_indexer` underneath, against the oracle's restatement of the
  label-slice rule (both ends inclusive, list-of-slices concatenated without
  de-duplication, reversed or outside bounds select nothing)."""
  lat = np.linspace(-90, 90, 19)
  lon = np.linspace(0, 360, 36, endpoint=False)
  data = np.arange(19 * 36, dtype=np.float32).reshape(19, 36)
  for name, spec in rc.REGION_SLICES.items():
    region = orc.SliceRegion(lat_slice=rc._slices(spec.get('lat')),  # pylint: disable=protected-access
                             lon_slice=rc._slices(spec.get('lon')))  # pylint: disable=protected-access
    x, w, rlat, rlon = orc._region_apply(region, data, np.ones_like(data),  # pylint: disable=protected-access
                                         lat, lon)
    np.testing.assert_array_equal(
        rlat, EXTRA_VECTORS[f'region/{name}/latitude'][1], err_msg=name)
    np.testing.assert_array_equal(
        rlon, EXTRA_VECTORS[f'region/{name}/longitude'][1], err_msg=name)
    np.testing.assert_array_equal(x, EXTRA_VECTORS[f'region/{name}/data'][1],
                                  err_msg=name)
    np.testing.assert_array_equal(w,
                                  EXTRA_VECTORS[f'region/{name}/weights'][1],
                                  err_msg=name)
  assert EXTRA_VECTORS['region/overlapping/latitude'][1].size == 10  # 5 + 5
  assert EXTRA_VECTORS['region/reversed_bounds/latitude'][1].size == 0
