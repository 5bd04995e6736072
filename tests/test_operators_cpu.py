"""The operator-level parity tests of the `-m gpu` suite, re-run WITHOUT a GPU
on the NumPy stand-in context (tests/fake_ctx.py): same test bodies, same
oracle values, but the raw C-ABI arguments the operators produce (offset
tables, gathers, region weight factors, member strides, result assembly) are
interpreted by NumPy instead of the CUDA kernels.  This is what the host logic
can be held to on a machine without a B200; the arithmetic of the kernels is
what the gpu-marked originals check."""
import functools

import fake_ctx
import test_det_metrics_gpu as det
import test_ens_metrics_gpu as ens
import test_evaluation_gpu as ev


def _on_stand_in(fn):
  @functools.wraps(fn)  # keeps the parametrize marks and the signature
  def run(*args, **kwargs):
    with fake_ctx.installed():
      return fn(*args, **kwargs)
  return run


# K1 operators: metric classes, regions (one pass for all), ACC climatology
# lookups, label joins, land masks
test_metric_classes_match_oracle = _on_stand_in(
    det.test_metric_classes_match_oracle)
test_acc_with_dayofyear_climatology = _on_stand_in(
    det.test_acc_with_dayofyear_climatology)
test_regions_and_batch_mode = _on_stand_in(det.test_regions_and_batch_mode)
test_land_region_and_combined = _on_stand_in(det.test_land_region_and_combined)
test_rmse_over_invalid_region = _on_stand_in(det.test_rmse_over_invalid_region)
test_shared_dims_are_joined_by_label = _on_stand_in(
    det.test_shared_dims_are_joined_by_label)
test_non_increasing_latitude_raises = _on_stand_in(
    det.test_non_increasing_latitude_raises)

# K2 operators: CRPS family, ensemble-mean / variance family, NaN rules
test_crps_parts_match_oracle = _on_stand_in(ens.test_crps_parts_match_oracle)
test_ensemble_mean_rmse_stddev_variance_debiased = _on_stand_in(
    ens.test_ensemble_mean_rmse_stddev_variance_debiased)
test_ensemble_size_1_gives_mae = _on_stand_in(
    ens.test_ensemble_size_1_gives_mae)
test_nan_forecasts = _on_stand_in(ens.test_nan_forecasts)
test_repeated_forecasts_are_okay = _on_stand_in(
    ens.test_repeated_forecasts_are_okay)
test_effect_of_large_bias_and_perfect_prediction = _on_stand_in(
    ens.test_effect_of_large_bias_and_perfect_prediction)
test_missing_ensemble_dim_raises = _on_stand_in(
    ens.test_missing_ensemble_dim_raises)

# the per-chunk loop
test_chunked_equals_unchunked = _on_stand_in(ev.test_chunked_equals_unchunked)

# wind-vector RMSE, energy score (K3), map outputs (K6 / K6e)
import test_maps_gpu as maps  # noqa: E402  pylint: disable=wrong-import-position

test_wind_vector_rmse = _on_stand_in(det.test_wind_vector_rmse)
test_energy_score = _on_stand_in(ens.test_energy_score)
test_energy_score_k3_regions_and_skipna_fallback = _on_stand_in(
    ens.test_energy_score_k3_regions_and_skipna_fallback)
test_spatial_det_maps_chunk_and_time_mean = _on_stand_in(
    maps.test_spatial_det_maps_chunk_and_time_mean)
test_spatial_det_maps_truth_gather_init_lead_layout = _on_stand_in(
    maps.test_spatial_det_maps_truth_gather_init_lead_layout)
test_spatial_ensemble_maps = _on_stand_in(maps.test_spatial_ensemble_maps)

# threshold / Gaussian metrics (K7): threshold tables, climatology gathers,
# quantile sums; the two probabilistic eval configs of scripts/evaluate.py
import test_eval_probabilistic_gpu as evp  # noqa: E402  pylint: disable=wrong-import-position
import test_threshold_metrics_gpu as thr  # noqa: E402  pylint: disable=wrong-import-position

for _name in ('test_gaussian_crps_and_variance_known_answers',
              'test_gaussian_brier_known_answers',
              'test_gaussian_ignorance_known_answers',
              'test_gaussian_rps_known_answers',
              'test_ensemble_brier_known_answers',
              'test_ensemble_brier_nan_propagates_unless_skipna',
              'test_ensemble_ignorance_and_rps_known_answers',
              'test_ensemble_threshold_metrics_match_oracle',
              'test_gaussian_metrics_match_oracle',
              'test_threshold_compute_matches_kernel_selection',
              'test_spatial_threshold_maps_match_oracle'):
  globals()[_name] = _on_stand_in(getattr(thr, _name))
test_ensemble_binary_config_with_regions = _on_stand_in(
    evp.test_ensemble_binary_config_with_regions)
test_probabilistic_and_spatial_configs = _on_stand_in(
    evp.test_probabilistic_and_spatial_configs)
test_metric_and_region_loop_many_metrics_regions = _on_stand_in(
    ev.test_metric_and_region_loop_many_metrics_regions)

# regridders (K5 / K8), zonal spectrum (K4 + latitude mean + interpolation),
# derived variables / preprocessing, SEEPS (K9), rank histogram (K10)
import test_derived_wind_gpu as wind  # noqa: E402  pylint: disable=wrong-import-position
import test_preprocessing_gpu as prep  # noqa: E402  pylint: disable=wrong-import-position
import test_rank_hist_gpu as rh  # noqa: E402  pylint: disable=wrong-import-position
import test_regrid_interp_gpu as rgi  # noqa: E402  pylint: disable=wrong-import-position
import test_regrid_spectrum_gpu as rgs  # noqa: E402  pylint: disable=wrong-import-position
import test_seeps_gpu as seeps  # noqa: E402  pylint: disable=wrong-import-position

for _mod, _names in (
    (rgs, ('test_regridding_extrapolation_known_answer',
           'test_expected_nans_and_oracle', 'test_regridding_nans_disc',
           'test_regrid_dataset_flips_latitude_and_keeps_dims',
           'test_spectrum_matches_oracle',
           'test_spectrum_longitude_first_layout_and_peak',
           'test_spectrum_parseval_and_time_sum')),
    (rgi, ('test_bilinear_longitude_periodicity_known_answers',
           'test_bilinear_latitude_poles_known_answers',
           'test_nearest_exact_known_answer',
           'test_regridders_match_oracle_on_random_fields')),
    (wind, ('test_wind_speed_is_bit_identical_to_numpy',
            'test_wind_speed_as_eval_derived_variable')),
    (prep, ('test_ensemble_mean_matches_numpy_mean',
            'test_interpolate_spectral_frequencies_matches_scipy_interp1d',
            'test_latitude_mean_spectrum_operator')),
    (seeps, ('test_seeps_known_answers',
             'test_seeps_random_fields_match_oracle')),
    (rh, ('test_rank_one_hot_matches_oracle_without_ties',
          'test_repeated_entries_get_random_bin',
          'test_partial_ties_stay_within_the_tied_bins',
          'test_bad_num_bins_raises'))):
  for _name in _names:
    globals()[_name] = _on_stand_in(getattr(_mod, _name))
