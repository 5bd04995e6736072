"""The in-memory evaluation cases of tests/test_evaluation_cpu.py --
BASELINE.json configs[0], climatology / probabilistic-climatology / persistence
forecasts, analysis as truth, the full config.Selection -- on the CUDA kernels
(real context instead of the NumPy stand-in), against the same oracle values.

Named to be collected LAST: these cases were added after the round's GPU
budget was spent, so their first execution on a B200 is the round-end run."""
import contextlib

import pytest

import test_evaluation_cpu as cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', cases.CASES, ids=lambda c: c.__name__[5:])
def test_in_memory_evaluation_on_device(case, tmp_path):
  case(tmp_path, contextlib.nullcontext, crps_rtol=5e-5, det_rtol=5e-6)
