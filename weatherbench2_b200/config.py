"""Evaluation configuration dataclasses -- same field names as
weatherbench2/config.py:28-138 so `scripts/evaluate.py` style callers drop in.
(Viz / Panel are plotting-only and out of scope.)"""
import dataclasses
import typing as t


@dataclasses.dataclass
class Selection:
  """Sub-set of forecast and truth data (weatherbench2/config.py:28-51)."""
  variables: t.Sequence[str]
  time_slice: slice
  levels: t.Optional[t.Sequence[int]] = None
  lat_slice: t.Optional[slice] = dataclasses.field(
      default_factory=lambda: slice(None, None))
  lon_slice: t.Optional[slice] = dataclasses.field(
      default_factory=lambda: slice(None, None))
  aux_variables: t.Optional[t.Sequence[str]] = None


@dataclasses.dataclass
class Paths:
  """Input / output locations (weatherbench2/config.py:54-70).

  `forecast`, `obs` and `climatology` may be in-memory datasets instead of
  zarr paths (zarr/xarray are not available on the GPU boxes; paths are
  honoured only when xarray + zarr are importable)."""
  forecast: t.Any
  obs: t.Any
  output_dir: str
  output_file_prefix: t.Optional[str] = ''
  climatology: t.Optional[t.Any] = None


@dataclasses.dataclass
class Data:
  """Selection + Paths (weatherbench2/config.py:73-93)."""
  selection: Selection
  paths: Paths
  by_init: t.Optional[bool] = True
  rename_variables: t.Optional[t.Dict[str, str]] = None
  pressure_level_suffixes: t.Optional[bool] = False


@dataclasses.dataclass
class Eval:
  """Evaluation configuration (weatherbench2/config.py:96-138)."""
  metrics: t.Dict[str, t.Any]
  regions: t.Optional[t.Dict[str, t.Any]] = None
  evaluate_persistence: t.Optional[bool] = False
  evaluate_climatology: t.Optional[bool] = False
  evaluate_probabilistic_climatology: t.Optional[bool] = False
  probabilistic_climatology_start_year: t.Optional[int] = None
  probabilistic_climatology_end_year: t.Optional[int] = None
  probabilistic_climatology_hour_interval: t.Optional[int] = None
  against_analysis: t.Optional[bool] = False
  derived_variables: t.Dict[str, t.Any] = dataclasses.field(
      default_factory=dict)
  temporal_mean: t.Optional[bool] = True
  output_format: str = 'netcdf'
