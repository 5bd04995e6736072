"""Region selectors -- same classes and `apply` contract as
weatherbench2/regions.py:25-158, plus `factors()`, the form the CUDA kernels
consume.

In the reference a region either slices the data (SliceRegion) or multiplies
the latitude weights by a mask (ExtraTropical / Land), and `_spatial_average`
then zeroes every cell whose weight is not positive
(weatherbench2/metrics.py:157-160).  Both are the same thing to a streaming
reduction: a per-cell weight factor, with weight-0 cells skipped.  `factors()`
returns that factor in separable form

    factor[lat, lon] = lat_factor[lat] * lon_factor[lon] * cell[lat, lon]

(`cell` is None unless a LandRegion is involved) so that one pass over the data
serves every region of an eval config.
"""
from __future__ import annotations

import dataclasses
import typing as t

import numpy as np

from weatherbench2_b200 import xarray_lite as xl


@dataclasses.dataclass
class RegionFactors:
  lat: np.ndarray                    # [nlat] float64 multiplicity / mask
  lon: np.ndarray                    # [nlon] float64
  cell: t.Optional[np.ndarray] = None  # [nlat, nlon] float64 or None

  def __mul__(self, other: 'RegionFactors') -> 'RegionFactors':
    cell = self.cell
    if other.cell is not None:
      cell = other.cell if cell is None else cell * other.cell
    return RegionFactors(self.lat * other.lat, self.lon * other.lon, cell)


@dataclasses.dataclass
class Region:
  """Region selector for spatially averaged metrics (regions.py:25-54)."""

  def apply(self, dataset, weights):
    """Returns (dataset, weights) restricted to the region (regions.py:40-54).

    Kept for API compatibility; works on the lite containers.  The metric
    kernels use `factors()` instead and never materialise the sliced data.
    """
    raise NotImplementedError

  def factors(self, latitude: np.ndarray, longitude: np.ndarray
              ) -> RegionFactors:
    raise NotImplementedError


def _slice_multiplicity(coord: np.ndarray, slices) -> np.ndarray:
  """How many times each index is selected by the concatenated label slices
  (regions.py:79-84: xr.concat of `.sel(slice)` results, no de-duplication)."""
  slices = slices if isinstance(slices, list) else [slices]
  mult = np.zeros(coord.size, dtype=np.float64)
  for s in slices:
    np.add.at(mult, xl.label_slice_indices(coord, s), 1.0)
  return mult


def _slice_indices(coord: np.ndarray, slices) -> np.ndarray:
  slices = slices if isinstance(slices, list) else [slices]
  return np.concatenate([xl.label_slice_indices(coord, s) for s in slices])


@dataclasses.dataclass
class SliceRegion(Region):
  """Latitude-longitude box selection (regions.py:57-95)."""

  lat_slice: t.Optional[t.Union[slice, list]] = dataclasses.field(
      default_factory=lambda: slice(None, None))
  lon_slice: t.Optional[t.Union[slice, list]] = dataclasses.field(
      default_factory=lambda: slice(None, None))

  def apply(self, dataset, weights):
    ilat = _slice_indices(dataset.latitude.values, self.lat_slice)
    ilon = _slice_indices(dataset.longitude.values, self.lon_slice)
    windex = {}
    if 'latitude' in weights.dims:
      windex['latitude'] = ilat
    if 'longitude' in weights.dims:
      windex['longitude'] = ilon
    return (dataset.isel(latitude=ilat, longitude=ilon), weights.isel(windex))

  def factors(self, latitude, longitude):
    return RegionFactors(_slice_multiplicity(latitude, self.lat_slice),
                         _slice_multiplicity(longitude, self.lon_slice))


@dataclasses.dataclass
class ExtraTropicalRegion(Region):
  """|lat| >= 20 mask on the weights (regions.py:98-109; the constant 20 is
  hard-coded at :108 and `threshold_lat` is ignored there -- kept as is)."""

  threshold_lat: t.Optional[float] = 20

  def apply(self, dataset, weights):
    region_weights = xl.DataArray(
        (np.abs(dataset.latitude.values) >= 20).astype(float), ('latitude',),
        {'latitude': dataset.latitude.values})
    return dataset, weights * region_weights

  def factors(self, latitude, longitude):
    return RegionFactors((np.abs(latitude) >= 20).astype(np.float64),
                         np.ones(longitude.size))


@dataclasses.dataclass
class LandRegion(Region):
  """Land-sea-mask weighting (regions.py:112-138)."""

  land_sea_mask: t.Any = None
  threshold: t.Optional[float] = None

  def _mask_lat_lon(self, latitude, longitude) -> np.ndarray:
    lsm = xl.from_xarray(self.land_sea_mask)
    if isinstance(lsm, xl.DataArray):
      extra = [d for d in lsm.dims if d not in ('latitude', 'longitude')]
      if extra:
        lsm = lsm.isel({d: 0 for d in extra})
      m = lsm.transpose('latitude', 'longitude').values
    else:
      m = np.asarray(lsm)
    if m.shape != (latitude.size, longitude.size):
      raise ValueError(
          f'land_sea_mask shape {m.shape} does not match the dataset grid '
          f'({latitude.size}, {longitude.size})')
    m = m.astype(np.float64)
    if self.threshold is not None:
      m = (m > self.threshold).astype(np.float64)  # regions.py:136-137
    elif np.isnan(m).any():
      # the mask becomes part of the weights of `dataset.weighted(...)`
      # (metrics.py:161), and xarray refuses weights with missing values
      raise ValueError('`weights` cannot contain missing values. Missing '
                       'values can be replaced by `weights.fillna(0)`.')
    return m

  def apply(self, dataset, weights):
    lat, lon = dataset.latitude.values, dataset.longitude.values
    land = xl.DataArray(self._mask_lat_lon(lat, lon),
                        ('latitude', 'longitude'),
                        {'latitude': lat, 'longitude': lon})
    return dataset, weights * land

  def factors(self, latitude, longitude):
    return RegionFactors(np.ones(latitude.size), np.ones(longitude.size),
                         self._mask_lat_lon(latitude, longitude))


@dataclasses.dataclass
class CombinedRegion(Region):
  """Sequentially applies regions (regions.py:141-158)."""

  regions: list = dataclasses.field(default_factory=list)

  def apply(self, dataset, weights):
    for region in self.regions:
      dataset, weights = region.apply(dataset, weights)
    return dataset, weights

  def factors(self, latitude, longitude):
    out = RegionFactors(np.ones(latitude.size), np.ones(longitude.size))
    for region in self.regions:
      out = out * region.factors(latitude, longitude)
    return out
