"""The REFERENCE's own unit tests, run with the reference's own modules on the
stand-ins of tests/golden/xarray_shim (xarray, jax.numpy, apache_beam,
xarray_beam) -- the same stand-ins that produced
tests/golden/reference_run_vectors.npz.  The tests carry known answers and
statistical properties established with the real libraries, so their passing
is the measure of how faithful the stand-ins are.

Needs the reference checkout (/root/reference, read-only, build container
only): skipped elsewhere.  Nothing is copied from it."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = '/root/reference'
SHIM = os.path.join(ROOT, 'tests', 'golden', 'xarray_shim')

# (test file, tests that must pass, names allowed to fail: derived variables
# outside the hot path that need Dataset.rolling / coordinate arithmetic)
SUITES = [
    ('metrics_test.py', 72, ()),
    ('regions_test.py', 1, ()),
    ('regridding_test.py', 40, ()),
    ('derived_variables_test.py', 12, (
        'testAggregatePrecipitationAccumulation',
        'testPrecipitationAccumulation24hr', 'testPrecipitationAccumulation6hr',
        'testRelativeHumidity')),
]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'weatherbench2')),
                    reason='the reference checkout is only in the build '
                    'container')
@pytest.mark.parametrize('name,n_pass,may_fail', SUITES,
                         ids=[s[0] for s in SUITES])
def test_reference_own_tests_pass_on_the_stand_ins(name, n_pass, may_fail,
                                                   tmp_path):
  env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1',
             PYTHONPATH=os.pathsep.join([ROOT, SHIM, REFERENCE]))
  run = subprocess.run(
      [sys.executable, '-m', 'pytest',
       os.path.join(REFERENCE, 'weatherbench2', name), '-q', '-p',
       'no:cacheprovider'], cwd=str(tmp_path), env=env, capture_output=True,
      text=True, timeout=900, check=False)
  out = run.stdout + run.stderr
  failed = re.findall(r'^FAILED \S+::(\w+)', out, re.M)
  passed = int((re.search(r'(\d+) passed', out) or [0, 0])[1])
  assert sorted(failed) == sorted(may_fail), out[-3000:]
  assert passed == n_pass, out[-2000:]
