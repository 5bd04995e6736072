"""Per-chunk evaluation loop -- the API surface of weatherbench2/evaluation.py
that sits directly on the metric kernels:

  _metric_and_region_loop   (evaluation.py:388-438)
  _evaluate_all_metrics     (evaluation.py:441-483)
  evaluate_in_memory        (evaluation.py:486-517)

The reference loops metric x region in Python and every iteration re-reads the
chunk; here the loop runs inside `metrics.batch(...)`, so the first metric of a
family launches ONE kernel pass for all regions (and shares it with ACC's
climatology pass) and the remaining iterations only slice the result.

Zarr / netCDF I/O and the Beam pipeline (evaluation.py:520-828) are out of
scope: datasets are passed in memory through
`config.Paths`; results are returned and saved as `.npz`.
"""
from __future__ import annotations

import logging
import os
import typing as t

import numpy as np

from weatherbench2_b200 import config
from weatherbench2_b200 import metrics as metrics_lib
from weatherbench2_b200 import xarray_lite as xl


def make_latitude_increasing(dataset: xl.Dataset) -> xl.Dataset:
  """Flip the dataset if latitude is decreasing (evaluation.py:40-46)."""
  lat = dataset['latitude'].values
  if (np.diff(lat) < 0).all():
    dataset = dataset.isel(latitude=np.arange(lat.size)[::-1])
  return dataset


def apply_time_conventions(forecast: xl.Dataset, by_init: bool) -> xl.Dataset:
  """WeatherBench2 time names on a forecast (weatherbench2/schema.py:25-44)."""
  if 'prediction_timedelta' in forecast.coords:
    forecast = forecast.rename({'prediction_timedelta': 'lead_time'})
    if by_init:
      forecast = forecast.rename({'time': 'init_time'})
      init = forecast['init_time']
      lead = forecast['lead_time']
      valid = init.values[:, None] + lead.values[None, :]
      forecast = forecast.assign_coords(
          valid_time=(('init_time', 'lead_time'), valid))
    else:
      t_ = forecast['time']
      lead = forecast['lead_time']
      forecast = forecast.assign_coords(
          init_time=(('time', 'lead_time'),
                     t_.values[:, None] - lead.values[None, :]))
  return forecast


def select_truth_at_valid_time(truth: xl.Dataset, forecast: xl.Dataset
                               ) -> xl.Dataset:
  """`truth.sel(time=forecast.valid_time)` (evaluation.py:475) WITHOUT
  materialising the (init_time, lead_time, ...) copy: each variable becomes a
  LazyGather view whose label lookup is folded into the kernels' offset table.
  """
  vt = forecast['valid_time']
  out = xl.Dataset(attrs=truth.attrs)
  for k in truth.keys():
    v = truth[k]
    if 'time' not in v.dims:
      out[k] = v
      continue
    pos = xl._lookup(v.coords['time'].values, vt.values.ravel())  # pylint: disable=protected-access
    out[k] = LazyGather(v, {'time': (vt.dims, pos.reshape(vt.shape))},
                        extra_coords={'valid_time': xl.Coord(vt.dims,
                                                             vt.values)})
  return out


LazyGather = xl.LazyGather  # defined next to the container it extends


def _metric_and_region_loop(forecast: xl.Dataset, truth: xl.Dataset,
                            eval_config: config.Eval, skipna: bool,
                            compute_chunk: bool = False) -> xl.Dataset:
  """Metric results looping over metrics and regions (evaluation.py:388-438).
  Result variables have dims (metric, [region], ...)."""
  native = xl.is_native_xarray(forecast)
  forecast = xl.from_xarray(forecast)
  truth = xl.from_xarray(truth)
  logging.info('Starting _metric_and_region_loop')
  for name, dv in eval_config.derived_variables.items():
    forecast[name] = dv.compute(forecast)
    truth[name] = dv.compute(truth)

  regions = (list(eval_config.regions.values())
             if eval_config.regions is not None else [None])
  climatology = None
  for metric in eval_config.metrics.values():
    if isinstance(metric, metrics_lib.ACC):
      climatology = xl.from_xarray(metric.climatology)
      break

  results = []
  with metrics_lib.batch(regions, climatology):
    for name, metric in eval_config.metrics.items():
      logging.info('Logging metric: %s', name)
      if compute_chunk or not eval_config.temporal_mean:
        eval_fn = metric.compute_chunk
      else:
        eval_fn = metric.compute
      if eval_config.regions is not None:
        tmp_results = []
        for region_name, region in eval_config.regions.items():
          tmp = xl.from_xarray(eval_fn(forecast=forecast, truth=truth,
                                       region=region, skipna=skipna))
          tmp_results.append(tmp.expand_dims(
              {'metric': np.array([name], dtype=object),
               'region': np.array([region_name], dtype=object)}))
        result = xl.concat(tmp_results, 'region')
      else:
        result = xl.from_xarray(eval_fn(
            forecast=forecast, truth=truth, skipna=skipna)).expand_dims(
                {'metric': np.array([name], dtype=object)})
      results.append(result)
  merged = xl.merge(results)
  return xl.to_xarray(merged) if native else merged


def open_forecast_and_truth_datasets(data_config: config.Data,
                                     eval_config: config.Eval):
  """In-memory restatement of evaluation.py:189-365 for datasets passed
  through `config.Paths` (variable / level / time selection, time
  conventions, latitude orientation).  Zarr paths need xarray + zarr."""
  def _open(obj, what):
    if obj is None:
      return None
    if isinstance(obj, (str, os.PathLike)):
      if not xl.have_xarray():
        raise RuntimeError(
            f'{what}: zarr paths need xarray + zarr, which are not installed; '
            'pass an in-memory dataset in config.Paths instead')
      import xarray as xr  # pylint: disable=import-outside-toplevel
      return xl.from_xarray(xr.open_zarr(obj))
    return xl.from_xarray(obj)

  sel = data_config.selection
  forecast = _open(data_config.paths.forecast, 'forecast')
  obs = _open(data_config.paths.obs, 'obs')
  climatology = _open(data_config.paths.climatology, 'climatology')
  if data_config.rename_variables is not None:
    forecast = forecast.rename(data_config.rename_variables)
  obs = make_latitude_increasing(obs)
  forecast = make_latitude_increasing(forecast)
  if climatology is not None:
    climatology = make_latitude_increasing(climatology)
  for coord_name in ('latitude', 'longitude'):  # evaluation.py:49-61
    np.testing.assert_allclose(forecast[coord_name].values,
                               obs[coord_name].values, atol=1e-3)
  # _ensure_aligned_grid (evaluation.py:49-61): after the closeness check the
  # forecast's coordinates REPLACE the others', so that float32 / float64
  # coordinate labels cannot turn the later label joins into gathers
  grid = {k: forecast[k].values for k in ('latitude', 'longitude')}
  obs = obs.assign_coords(grid)
  if climatology is not None:
    for coord_name in ('latitude', 'longitude'):
      np.testing.assert_allclose(forecast[coord_name].values,
                                 climatology[coord_name].values, atol=1e-3)
    climatology = climatology.assign_coords(grid)
  forecast = apply_time_conventions(forecast, data_config.by_init)

  variables = list(sel.variables)
  forecast = forecast[[v for v in variables if v in forecast.keys()] +
                      [v for v in (sel.aux_variables or [])
                       if v in forecast.keys()]]
  obs = obs[[v for v in variables if v in obs.keys()]]
  if sel.levels is not None:
    lv = np.asarray(sel.levels)
    forecast = forecast.sel(level=lv) if 'level' in forecast.dims else forecast
    obs = obs.sel(level=lv) if 'level' in obs.dims else obs
  time_dim = 'init_time' if data_config.by_init else 'time'
  if sel.time_slice is not None and sel.time_slice != slice(None, None):
    forecast = forecast.sel({time_dim: _as_datetime_slice(sel.time_slice)})
  if eval_config.against_analysis:
    raise NotImplementedError('against_analysis is outside the hot path')
  return forecast, obs, climatology


def _as_datetime_slice(s: slice) -> slice:
  def cv(x):
    return None if x is None else np.datetime64(x, 'ns')
  return slice(cv(s.start), cv(s.stop))


def _get_output_path(data_config: config.Data, eval_name: str, fmt: str) -> str:
  # evaluation.py:368-380
  suffix = {'netcdf': 'nc', 'zarr': 'zarr', 'npz': 'npz'}[fmt]
  return os.path.join(
      data_config.paths.output_dir,
      f'{data_config.paths.output_file_prefix}{eval_name}.{suffix}')


def save_results(results: xl.Dataset, path: str) -> None:
  """netCDF when xarray is importable (evaluation.py:383-385), else .npz
  holding every variable, its dims and the coordinates."""
  if xl.have_xarray() and path.endswith('.nc'):
    xl.to_xarray(results).to_netcdf(path)
    return
  payload = {}
  for k in results.keys():
    payload[f'var:{k}'] = results[k].values
    payload[f'dims:{k}'] = np.array(results[k].dims, dtype=object)
  for k, c in results.coords.items():
    payload[f'coord:{k}'] = c.values
    payload[f'coorddims:{k}'] = np.array(c.dims, dtype=object)
  os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
  np.savez(path, **payload, allow_pickle=True)


def _evaluate_all_metrics(eval_name: str, eval_config: config.Eval,
                          data_config: config.Data, skipna: bool):
  """Evaluate a set of eval metrics in memory (evaluation.py:441-483)."""
  forecast, truth, climatology = open_forecast_and_truth_datasets(
      data_config, eval_config)
  if (eval_config.evaluate_climatology or eval_config.evaluate_persistence or
      eval_config.evaluate_probabilistic_climatology):
    raise NotImplementedError(
        'climatology / persistence forecasts are data preparation '
        '(evaluation.py:450-472), outside the hot path')
  del climatology
  if data_config.by_init:
    truth = select_truth_at_valid_time(truth, forecast)  # evaluation.py:475
  results = _metric_and_region_loop(forecast, truth, eval_config,
                                    skipna=skipna)
  fmt = 'netcdf' if xl.have_xarray() else 'npz'
  output_path = _get_output_path(data_config, eval_name, fmt)
  save_results(results, output_path)
  logging.info('Logging Saved results to %s', output_path)
  return results


def evaluate_in_memory(data_config: config.Data,
                       eval_configs: dict, skipna: bool = False) -> dict:
  """Run evaluation in memory, one results file per config.Eval
  (evaluation.py:486-517).  Also returns {eval_name: results}."""
  out = {}
  for eval_name, eval_config in eval_configs.items():
    out[eval_name] = _evaluate_all_metrics(eval_name, eval_config, data_config,
                                           skipna=skipna)
  return out
