"""The REFERENCE's own unit tests, run with the reference's own modules on the
stand-ins of tests/golden/xarray_shim (xarray, jax.numpy, apache_beam,
xarray_beam) -- the same stand-ins that produced
tests/golden/reference_run_vectors.npz.  The tests carry known answers and
statistical properties established with the real libraries, so their passing
is the measure of how faithful the stand-ins are.

Needs the reference checkout (/root/reference, read-only, build container
only): skipped elsewhere.  Nothing is copied from it."""
import pytest

import reference_suite_runner as runner

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


@pytest.mark.skipif(not runner.available(),
                    reason='the reference checkout is only in the build '
                    'container')
@pytest.mark.parametrize('name,n_pass,may_fail', SUITES,
                         ids=[s[0] for s in SUITES])
def test_reference_own_tests_pass_on_the_stand_ins(name, n_pass, may_fail):
  passed, failed, tail = runner.result('shim', name)
  assert failed == sorted(may_fail), tail
  assert passed == n_pass, tail
