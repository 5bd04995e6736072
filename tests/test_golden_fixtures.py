"""The committed golden fixture (tests/golden/reference_known_answers.json,
transcribed from the reference's own tests by tests/golden/make_golden.py)
against the oracle (CPU) and against the CUDA path (GPU)."""
import json
import os

import numpy as np
import pytest

from oracle import wb2_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, 'golden',
                                'reference_known_answers.json')))


def _arr(x):
  return np.array([[np.nan if v is None else v for v in row]
                   if isinstance(row, list) else
                   (np.nan if row is None else row) for row in x], dtype=float)


def test_oracle_matches_fixture():
  g = G['lat_weights_6']
  np.testing.assert_allclose(orc.get_lat_weights(np.array(g['latitude'])),
                             g['weights'], rtol=1e-12)
  g = G['regrid_lat_weights']
  np.testing.assert_allclose(orc.conservative_latitude_weights(
      np.array(g['source_lat']), np.array(g['target_lat']), True, True),
                             g['weights'], atol=1e-6)
  g = G['regrid_lon_weights_same_branch']
  np.testing.assert_allclose(orc.conservative_longitude_weights(
      np.array(g['source_lon']), np.array(g['target_lon']), True, True) * 6,
                             g['weights_times_6'], atol=1e-4)
  g = G['align_phase_with']
  for x, y, e in g['cases']:
    assert orc.align_phase_with(x, y, g['period']) == e
  g = G['regrid_extrapolation']
  src = orc.Grid(longitudes=np.array(g['source_lon']),
                 latitudes=np.array(g['source_lat']), periodic=False,
                 includes_poles=False)
  tgt = orc.Grid(longitudes=np.array(g['target_lon']),
                 latitudes=np.array(g['target_lat']), periodic=False,
                 includes_poles=False)
  np.testing.assert_allclose(orc.conservative_regrid(np.array(g['field']),
                                                     src, tgt),
                             _arr(g['expected']), atol=1e-6)


def test_product_host_weights_match_fixture():
  from weatherbench2_b200 import _spatial as sp, regridding as rg
  g = G['lat_weights_6']
  np.testing.assert_allclose(sp.lat_weights(np.array(g['latitude'])),
                             g['weights'], rtol=1e-12)
  g = G['regrid_lat_weights']
  np.testing.assert_allclose(rg._conservative_latitude_weights(
      np.array(g['source_lat']), np.array(g['target_lat']), True, True),
                             g['weights'], atol=1e-6)
  g = G['align_phase_with']
  for x, y, e in g['cases']:
    assert rg._align_phase_with(x, y, g['period']) == e


@pytest.mark.gpu
def test_cuda_path_matches_fixture():
  from weatherbench2_b200 import metrics, regions as R, regridding as rg
  from weatherbench2_b200 import xarray_lite as xl
  g = G['regrid_extrapolation']
  kw = dict(includes_poles=False, periodic=False)
  r = rg.ConservativeRegridder(
      rg.Grid(longitudes=np.array(g['source_lon']),
              latitudes=np.array(g['source_lat']), **kw),
      rg.Grid(longitudes=np.array(g['target_lon']),
              latitudes=np.array(g['target_lat']), **kw))
  np.testing.assert_allclose(r.regrid_array(np.array(g['field'])),
                             _arr(g['expected']), atol=1e-6)
  g = G['rmse_over_invalid_region']
  for bad in (np.nan, np.inf):
    tv = np.array([0.0, bad, 0.0]).reshape(1, 1, 3)
    coords = {'latitude': np.array(g['latitude'], float),
              'longitude': np.array([0.0]), 'time': np.array([0])}
    dims = ('time', 'longitude', 'latitude')
    truth = xl.Dataset({'w': (dims, tv)}, coords)
    forecast = xl.Dataset({'w': (dims, tv + 1)}, coords)
    rmse = metrics.RMSESqrtBeforeTimeAvg()
    assert np.isnan(rmse.compute(forecast, truth)['w'].values)
    np.testing.assert_allclose(
        rmse.compute(forecast, truth,
                     region=R.ExtraTropicalRegion())['w'].values,
        g['extra_tropics'])
  g = G['wind_vector_rmse_per_level']
  lat = np.linspace(-90, 90, 7)
  lon = np.linspace(0, 360, 12, endpoint=False)
  dims = ('time', 'level', 'longitude', 'latitude')
  coords = {'time': np.arange(2), 'level': np.arange(3), 'latitude': lat,
            'longitude': lon}

  def field(vals):
    a = np.array([np.nan if v is None else v for v in vals], np.float32)
    return np.broadcast_to(a[None, :, None, None], (2, 3, 12, 7)).copy()

  fds = xl.Dataset({'u': (dims, field(g['forecast_u'])),
                    'v': (dims, field(g['forecast_v']))}, coords)
  tds = xl.Dataset({'u': (dims, field(g['truth_u'])),
                    'v': (dims, field(g['truth_v']))}, coords)
  wv = metrics.WindVectorRMSESqrtBeforeTimeAvg('u', 'v', 'wind')
  np.testing.assert_allclose(wv.compute(fds, tds).values,
                             _arr([g['expected']])[0])
