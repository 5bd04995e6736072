"""The committed golden fixture (tests/golden/reference_known_answers.json:
known answers transcribed from the reference's own unit tests) for the
probabilistic / threshold / categorical metrics and the nearest / bilinear
regridders, against the oracle (CPU) and against the CUDA path (GPU)."""
import json
import os

import numpy as np
import pytest

from oracle import wb2_oracle as orc
import wb2_testdata as td

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, 'golden',
                                'reference_known_answers.json')))
T2M = '2m_temperature'
PRECIP = 'total_precipitation_24hr'


def _nan(x):
  return np.array([[np.nan if v is None else v for v in row] for row in x],
                  dtype=float)


def _expected(v):
  return np.inf if v == 'inf' else v


# ---- oracle (CPU) --------------------------------------------------------------
def test_oracle_gaussian_and_ensemble_scores_match_fixture():
  g = G['gaussian_crps']
  np.testing.assert_allclose(orc.gaussian_crps_pointwise(
      np.float32(g['forecast_mean']), np.float32(g['forecast_std']),
      np.float32(g['truth'])), g['expected'], rtol=1e-6)
  g = G['gaussian_brier']
  thr = orc.gaussian_quantile_threshold(np.float32(g['clim_mean']),
                                        np.float32(g['clim_std']),
                                        g['quantile'])
  for c in g['cases']:
    f = np.float32(1.0 + c['error'])
    t = np.float32(g['truth'])
    np.testing.assert_allclose(orc.gaussian_brier_pointwise(f, f, t, thr),
                               c['gaussian_quantile'], rtol=g['rtol'])
    np.testing.assert_allclose(
        orc.gaussian_brier_pointwise(f, f, t, np.float32(g['truth'])),
        c['quantile'], rtol=g['rtol'])
  g = G['gaussian_ignorance']
  for c in g['cases']:
    f = np.float32(1.0 + c['error'])
    np.testing.assert_allclose(orc.gaussian_ignorance_pointwise(
        f, f, np.float32(g['truth']), thr), c['expected'], rtol=g['rtol'])
  g = G['gaussian_rps']
  for c in g['cases']:
    f = np.float32(1.0 + c['error'])
    got = sum(orc.gaussian_rps_part_pointwise(f, f, np.float32(g['truth']),
                                              np.float32(q))
              for q in g['thresholds'])
    np.testing.assert_allclose(got, c['expected'], rtol=g['rtol'])
  g = G['ensemble_brier']
  thr = orc.gaussian_quantile_threshold(np.float32(g['clim_mean']),
                                        np.float32(g['clim_std']),
                                        g['quantile'])
  for c in g['cases']:
    x = (1.0 + c['error'] + c['ens_delta'] * np.array(g['member_offsets'])
         ).astype(np.float32)
    np.testing.assert_allclose(orc.ens_brier_pointwise(
        x, np.float32(g['truth']), thr, 0, False, False), c['expected'],
                               rtol=1e-4, atol=1e-12)
  g = G['ensemble_ignorance']
  for c in g['cases']:
    x = np.full(g['nmember'], 1.0 + c['error'], np.float32)
    got = orc.ens_ignorance_pointwise(x, np.float32(g['truth']), thr, 0, False)
    assert got == _expected(c['expected'])
  g = G['ensemble_rps']
  for c in g['cases']:
    x = np.full(g['nmember'], 1.0 + c['error'], np.float32)
    got = sum(orc.ens_rps_part_pointwise(x, np.float32(g['truth']),
                                         np.float32(q), 0, False)
              for q in g['thresholds'])
    assert got == c['expected']
  g = G['seeps']
  p1 = np.full((2, 3), g['dry_fraction'], np.float32)
  wet = np.full((2, 3), g['wet_threshold'], np.float32)
  t = np.full((2, 3), g['truth'], np.float32)
  for c in g['cases']:
    f = np.full((2, 3), c['forecast'], np.float32)
    np.testing.assert_allclose(orc.seeps_pointwise(f, t, wet, wet, p1),
                               c['expected'], atol=g['atol'])


def test_oracle_regridders_match_fixture():
  g = G['bilinear_longitude_periodicity']
  for periodic, key in ((True, 'periodic'), (False, 'not_periodic')):
    src = orc.Grid(np.array(g['source_lon']), np.array([0]), periodic, True)
    tgt = orc.Grid(np.array(g['target_lon']), np.array([0]), periodic, True)
    np.testing.assert_allclose(
        orc.bilinear_regrid(np.array(g['field']), src, tgt), _nan(g[key]),
        atol=1e-6)
  for c in G['bilinear_latitude_poles']['cases']:
    src = orc.Grid(np.array([0.0]), np.array(c['source_lat']), True,
                   c['poles'])
    tgt = orc.Grid(np.array([0.0]), np.array(c['target_lat']), True,
                   c['poles'])
    np.testing.assert_allclose(
        orc.bilinear_regrid(np.array(c['field'])[np.newaxis], src, tgt),
        _nan(c['expected']), atol=1e-6)
  g = G['nearest_exact']
  src = orc.Grid(np.array(g['source_lon']), np.array(g['source_lat']))
  tgt = orc.Grid(np.array(g['target_lon']), np.array(g['target_lat']))
  np.testing.assert_allclose(orc.nearest_regrid(np.array(g['field']), src, tgt),
                             g['expected'], atol=1e-6)


# ---- CUDA path (GPU) -----------------------------------------------------------
def _ds(vars, coords):  # pylint: disable=redefined-builtin
  from weatherbench2_b200 import xarray_lite as xl
  return xl.Dataset({k: (d, v) for k, (d, v) in vars.items()}, coords)


def _shift(ds, delta):
  return {'vars': {k: (d, v + delta) for k, (d, v) in ds['vars'].items()},
          'coords': ds['coords']}


def _gaussian_clim(truth, name):
  dims, arr = truth['vars'][name]
  first = np.take(arr, 0, axis=dims.index('time'))
  sdims = tuple(d for d in dims if d != 'time')
  coords = {k: v for k, v in truth['coords'].items() if k != 'time'}
  coords['dayofyear'] = np.arange(1, 367)
  rep = np.broadcast_to(first, (366,) + first.shape).copy()
  return _ds({name: (('dayofyear',) + sdims, rep),
              name + '_std': (('dayofyear',) + sdims, rep.copy())}, coords)


KW = dict(variables_3d=[], time_start='2022-01-01', time_stop='2022-01-02')


@pytest.mark.gpu
def test_cuda_gaussian_and_ensemble_scores_match_fixture():
  from weatherbench2_b200 import metrics, thresholds
  g = G['gaussian_brier']
  truth = _shift(td.mock_truth_data(variables_2d=[T2M], **KW), g['truth'])
  clim = _gaussian_clim(truth, T2M)
  for c in g['cases']:
    forecast = _shift(td.mock_forecast_data(
        variables_2d=[T2M, T2M + '_std'], lead_stop='1 day', **KW),
                      1.0 + c['error'])
    thr = thresholds.GaussianQuantileThreshold(clim, g['quantile'])
    res = metrics.GaussianBrierScore([thr]).compute(_ds(**forecast),
                                                    _ds(**truth))
    np.testing.assert_allclose(res[T2M].values, c['gaussian_quantile'],
                               rtol=g['rtol'])
    want = [cc['expected'] for cc in G['gaussian_ignorance']['cases']
            if cc['error'] == c['error']][0]
    res = metrics.GaussianIgnoranceScore([thr]).compute(_ds(**forecast),
                                                        _ds(**truth))
    np.testing.assert_allclose(res[T2M].values, want, rtol=g['rtol'])
  g = G['ensemble_brier']
  for c in g['cases']:
    forecast = td.mock_forecast_data(variables_2d=[T2M], ensemble_size=4,
                                     lead_stop='1 day', **KW)
    d, v = forecast['vars'][T2M]
    off = np.array(g['member_offsets']).reshape((4,) + (1,) * (v.ndim - 1))
    forecast['vars'][T2M] = (d, (v + 1.0 + c['error'] + c['ens_delta'] * off
                                 ).astype(np.float32))
    truth = _shift(td.mock_truth_data(variables_2d=[T2M], **KW), g['truth'])
    thr = thresholds.GaussianQuantileThreshold(_gaussian_clim(truth, T2M),
                                               g['quantile'])
    res = metrics.EnsembleBrierScore([thr]).compute(_ds(**forecast),
                                                    _ds(**truth))
    np.testing.assert_allclose(res[T2M].values, c['expected'], rtol=1e-4,
                               atol=1e-12)
  g = G['ensemble_ignorance']
  for c in g['cases']:
    forecast = _shift(td.mock_forecast_data(
        variables_2d=[T2M], ensemble_size=g['nmember'], lead_stop='1 day',
        **KW), 1.0 + c['error'])
    truth = _shift(td.mock_truth_data(variables_2d=[T2M], **KW), g['truth'])
    thr = thresholds.GaussianQuantileThreshold(_gaussian_clim(truth, T2M),
                                               g['quantile'])
    res = metrics.EnsembleIgnoranceScore([thr]).compute(_ds(**forecast),
                                                        _ds(**truth))
    np.testing.assert_allclose(res[T2M].values, _expected(c['expected']),
                               rtol=1e-4)


@pytest.mark.gpu
def test_cuda_regridders_match_fixture():
  from weatherbench2_b200 import regridding as rg
  g = G['bilinear_longitude_periodicity']
  for periodic, key in ((True, 'periodic'), (False, 'not_periodic')):
    kw = dict(includes_poles=True, periodic=periodic)
    r = rg.BilinearRegridder(
        rg.Grid(longitudes=np.array(g['source_lon']), latitudes=np.array([0]),
                **kw),
        rg.Grid(longitudes=np.array(g['target_lon']), latitudes=np.array([0]),
                **kw))
    np.testing.assert_allclose(r.regrid_array(np.array(g['field'])),
                               _nan(g[key]), atol=1e-6)
  for c in G['bilinear_latitude_poles']['cases']:
    kw = dict(includes_poles=c['poles'], periodic=True)
    r = rg.BilinearRegridder(
        rg.Grid(longitudes=np.array([0.0]),
                latitudes=np.array(c['source_lat']), **kw),
        rg.Grid(longitudes=np.array([0.0]),
                latitudes=np.array(c['target_lat']), **kw))
    np.testing.assert_allclose(
        r.regrid_array(np.array(c['field'])[np.newaxis]), _nan(c['expected']),
        atol=1e-6)
  g = G['nearest_exact']
  kw = dict(includes_poles=True, periodic=True)
  r = rg.NearestRegridder(
      rg.Grid(longitudes=np.array(g['source_lon']),
              latitudes=np.array(g['source_lat']), **kw),
      rg.Grid(longitudes=np.array(g['target_lon']),
              latitudes=np.array(g['target_lat']), **kw))
  np.testing.assert_allclose(r.regrid_array(np.array(g['field'])),
                             g['expected'], atol=1e-6)
