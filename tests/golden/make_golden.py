#!/usr/bin/env python
"""Writes tests/golden/reference_known_answers.json.

The reference (google-research/weatherbench2) cannot be imported in this
container or on the GPU box (it needs xarray / jax / apache_beam; none are
installed and there is no network), so golden vectors cannot be produced by
running it.  What the reference DOES publish are the known-answer constants
inside its own unit tests; this script transcribes them (with the test
file:line each one comes from) into a JSON fixture, so that both the oracle
(tests/test_oracle_golden.py) and the CUDA path are pinned to the same numbers
without reading /root/reference at test time.

Run:  python tests/golden/make_golden.py
"""
import json
import math
import os

S3 = math.sqrt(3.0)
GOLDEN = {
    'lat_weights_6': {
        'source': 'weatherbench2/metrics_test.py:63-82',
        'latitude': [-75, -45, -15, 15, 45, 75],
        'weights': [3 * (1 - S3 / 2), 3 * (S3 - 1) / 2, 1.5, 1.5,
                    3 * (S3 - 1) / 2, 3 * (1 - S3 / 2)],
    },
    'wind_vector_rmse_per_level': {
        'source': 'weatherbench2/metrics_test.py:84-131',
        'forecast_u': [0, 3, None], 'forecast_v': [0, -4, 1],
        'truth_u': [0, -3, None], 'truth_v': [0, 4, 1],
        'expected': [0, 10, None],
    },
    'rmse_over_invalid_region': {
        'source': 'weatherbench2/metrics_test.py:133-152',
        'latitude': [-45, 0, 45], 'global': None, 'extra_tropics': 1.0,
    },
    'regrid_lat_weights': {
        'source': 'weatherbench2/regridding_test.py:252-271',
        'source_lat': [-75, -45, -15, 15, 45, 75], 'target_lat': [-45, 45],
        'weights': [[1 - S3 / 2, (S3 - 1) / 2, 0.5, 0, 0, 0],
                    [0, 0, 0, 0.5, (S3 - 1) / 2, 1 - S3 / 2]],
    },
    'regrid_lon_weights_same_branch': {
        'source': 'weatherbench2/regridding_test.py:285-301',
        'source_lon': [0, 60, 120, 180, 240, 300],
        'target_lon': [0, 90, 180, 270],
        'weights_times_6': [[4, 1, 0, 0, 0, 1], [0, 3, 3, 0, 0, 0],
                            [0, 0, 1, 4, 1, 0], [0, 0, 0, 0, 3, 3]],
    },
    'regrid_extrapolation': {
        'source': 'weatherbench2/regridding_test.py:313-330',
        'source_lon': [1, 3, 5], 'source_lat': [1, 3],
        'target_lon': [0, 2, 4], 'target_lat': [0, 2],
        'field': [[1, 1], [2, 2], [3, 3]],
        'expected': [[None, None], [None, 1.5], [None, 2.5]],
    },
    'align_phase_with': {
        'source': 'weatherbench2/regridding_test.py:273-283',
        'period': 10,
        'cases': [[1, 0, 1], [-1, 0, -1], [5, 0, 5], [6, 0, -4], [1, 9, 11],
                  [5, 9, 5]],
    },
    # --- probabilistic / threshold / categorical metrics ----------------------
    # mock data are constant fields, so the point-wise score IS the average
    'gaussian_crps': {
        'source': 'weatherbench2/metrics_test.py:286-304',
        'forecast_mean': 1.0, 'forecast_std': 1.0, 'truth': 1.02,
        'expected': 0.23385455,
    },
    'gaussian_brier': {
        'source': 'weatherbench2/metrics_test.py:370-431',
        'truth': 1.0, 'clim_mean': 1.0, 'clim_std': 1.0, 'quantile': 0.8,
        # forecast mean = forecast std = 1 + error
        'cases': [{'error': 0.02, 'gaussian_quantile': 0.04421,
                   'quantile': 0.257883},
                  {'error': 1e6, 'gaussian_quantile': 0.70786,
                   'quantile': 0.707861}],
        'rtol': 1e-4,
    },
    'gaussian_ignorance': {
        'source': 'weatherbench2/metrics_test.py:436-475',
        'truth': 1.0, 'clim_mean': 1.0, 'clim_std': 1.0, 'quantile': 0.8,
        'cases': [{'error': 0.02, 'expected': 0.236055},
                  {'error': 1e6, 'expected': 1.841019}],
        'rtol': 1e-4,
    },
    'gaussian_rps': {
        'source': 'weatherbench2/metrics_test.py:480-534',
        'truth': 1.0, 'thresholds': [0.0, 1.0, 2.0],
        'cases': [{'error': 0.02, 'expected': 0.295746},
                  {'error': 1e6, 'expected': 0.758203}],
        'rtol': 1e-4,
    },
    'ensemble_brier': {
        'source': 'weatherbench2/metrics_test.py:989-1029',
        'truth': 1.0, 'clim_mean': 1.0, 'clim_std': 1.0, 'quantile': 0.2,
        'member_offsets': [-2, -1, 0, 1],
        # members = 1 + error + ens_delta * member_offsets
        'cases': [{'error': 0.0, 'ens_delta': 0.1, 'expected': 0.0},
                  {'error': 0.0, 'ens_delta': 1.0, 'expected': 0.25},
                  {'error': -10.0, 'ens_delta': 0.1, 'expected': 1.0}],
    },
    'ensemble_ignorance': {
        'source': 'weatherbench2/metrics_test.py:1294-1329',
        'truth': 1.0, 'clim_mean': 1.0, 'clim_std': 1.0, 'quantile': 0.2,
        'nmember': 4,
        'cases': [{'error': 0.0, 'expected': 0.0},
                  {'error': -10.0, 'expected': 'inf'}],
    },
    'ensemble_rps': {
        'source': 'weatherbench2/metrics_test.py:1334-1390',
        'truth': 1.5, 'thresholds': [0.0, 1.0, 2.0], 'nmember': 4,
        'cases': [{'error': 0.02, 'expected': 0.0},
                  {'error': -2.0, 'expected': 2.0}],
    },
    'seeps': {
        'source': 'weatherbench2/metrics_test.py:1392-1440',
        'dry_fraction': 0.4, 'wet_threshold': 1.0, 'truth': 0.0,
        'cases': [{'forecast': 0.0, 'expected': 0.0},
                  {'forecast': 0.5, 'expected': 1.25}],
        'atol': 1e-4,
    },
    # --- nearest / bilinear regridders ------------------------------------------
    'bilinear_longitude_periodicity': {
        'source': 'weatherbench2/regridding_test.py:495-525',
        'source_lon': [0.0, 90.0, 180.0, 270.0],
        'target_lon': [45.0, 135.0, 225.0, 315.0],
        'field': [[0.0], [1.0], [2.0], [3.0]],
        'periodic': [[0.5], [1.5], [2.5], [1.5]],
        'not_periodic': [[0.5], [1.5], [2.5], [None]],
    },
    'bilinear_latitude_poles': {
        'source': 'weatherbench2/regridding_test.py:527-572',
        'cases': [
            {'poles': True, 'source_lat': [-90.0, -30.0, 30.0, 90.0],
             'target_lat': [-60.0, 0.0, 60.0], 'field': [0.0, 1.0, 2.0, 3.0],
             'expected': [[0.5, 1.5, 2.5]]},
            {'poles': True, 'source_lat': [-60.0, 0.0, 60.0],
             'target_lat': [-90.0, -30.0, 30.0, 90.0],
             'field': [0.0, 1.0, 2.0], 'expected': [[0.0, 0.5, 1.5, 2.0]]},
            {'poles': False, 'source_lat': [-60.0, -20.0, 20.0, 60.0],
             'target_lat': [-70.0, 0.0, 70.0], 'field': [0.0, 1.0, 2.0, 3.0],
             'expected': [[None, 1.5, None]]}],
    },
    'nearest_exact': {
        'source': 'weatherbench2/regridding_test.py:574-591',
        'source_lon': [0, 90, 180, 270], 'source_lat': [-30, 0, 30],
        'target_lon': [0, 180], 'target_lat': [-30, 0, 30],
        'field': [[0, 1, 2], [4, 5, 6], [7, 8, 9], [10, 11, 12]],
        'expected': [[0, 1, 2], [7, 8, 9]],
    },
}

if __name__ == '__main__':
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                      'reference_known_answers.json')
  with open(path, 'w') as fh:
    json.dump(GOLDEN, fh, indent=1)
  print('wrote', path)
