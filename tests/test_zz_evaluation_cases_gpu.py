"""The in-memory evaluation cases of tests/test_evaluation_cpu.py --
BASELINE.json configs[0], climatology / probabilistic-climatology / persistence
forecasts, analysis as truth, the full config.Selection -- on the CUDA kernels
(real context instead of the NumPy stand-in), against the same oracle values.

Named to be collected LAST: these cases were added after the round's GPU
budget was spent, so their first execution on a B200 is the round-end run."""
import contextlib

import pytest

import test_evaluation_cpu as cases
import test_official_configs as official
import test_statistical_cases as stat

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', cases.CASES, ids=lambda c: c.__name__[5:])
def test_in_memory_evaluation_on_device(case, tmp_path):
  case(tmp_path, contextlib.nullcontext, crps_rtol=5e-5, det_rtol=5e-6)


def test_gaussian_crps_converges_on_device():
  stat.case_gaussian_crps_converges(contextlib.nullcontext, exact_rtol=1e-4)


@pytest.mark.parametrize('ensemble_size,num_bins', stat.RANK_HIST_CASES)
def test_rank_histogram_calibration_on_device(ensemble_size, num_bins):
  stat.case_rank_histogram_calibration(contextlib.nullcontext, ensemble_size,
                                       num_bins)


@pytest.mark.parametrize('case', stat.COARSE_GRID_CASES,
                         ids=lambda c: '-'.join(c))
def test_coarse_grid_interpolates_on_device(case):
  stat.case_coarse_grid_interpolates(contextlib.nullcontext, *case)


def test_official_deterministic_configs_on_device(tmp_path):
  official.case_official_deterministic_configs(tmp_path, contextlib.nullcontext)


def test_official_probabilistic_configs_on_device(tmp_path):
  official.case_official_probabilistic_configs(tmp_path, contextlib.nullcontext)


import test_reference_run_vectors as refrun  # noqa: E402  pylint: disable=wrong-import-position


@pytest.mark.parametrize('case', refrun.rc.CASES, ids=lambda c: c['id'])
def test_cuda_operators_match_the_reference_run(case):
  """The vectors the reference's own metrics.py produced (tests/golden/
  make_reference_vectors.py) against the CUDA kernels."""
  refrun.check_against_vectors(
      case, refrun.run_product(case, contextlib.nullcontext), rtol=2e-5,
      atol=2e-6)


def test_evaluate_in_memory_matches_the_reference_run_on_device(tmp_path):
  refrun.check_product_evaluations(contextlib.nullcontext, tmp_path, rtol=2e-5,
                                   atol=2e-6)


def test_extras_match_the_reference_run_on_device():
  refrun.check_product_extras(contextlib.nullcontext)
