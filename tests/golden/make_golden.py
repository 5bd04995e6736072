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
}

if __name__ == '__main__':
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                      'reference_known_answers.json')
  with open(path, 'w') as fh:
    json.dump(GOLDEN, fh, indent=1)
  print('wrote', path)
