#!/bin/bash
# Runs the REFERENCE's own unit tests (read from /root/reference, not copied)
# with the reference's own modules on the stand-ins of tests/golden/xarray_shim
# (xarray, jax.numpy, apache_beam, xarray_beam).  This validates the stand-ins:
# the tests carry known answers and statistical properties established with the
# real libraries.  Only possible in the build container (/root/reference).
#   bash tests/golden/run_reference_tests.sh > profiles/r2_reference_tests_on_shim.txt
set -u
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
cd /tmp
export PYTHONDONTWRITEBYTECODE=1
export PYTHONPATH="$ROOT:$ROOT/tests/golden/xarray_shim:/root/reference"
for f in metrics_test regions_test regridding_test derived_variables_test; do
  echo "== weatherbench2/$f.py"
  python -m pytest "/root/reference/weatherbench2/$f.py" -q -p no:cacheprovider \
      2>&1 | grep -E "^FAILED|passed|failed" | sed 's/ - .*//'
done
