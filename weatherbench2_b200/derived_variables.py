"""Derived variables on the hot path: `ZonalEnergySpectrum` with the API of
weatherbench2/derived_variables.py:29-56, 531-626.  The rFFT, the power
spectrum and the circumference scaling run in csrc/spectrum.cu.
(The other derived variables of the reference -- wind speed, vorticity, ... --
are cheap stencils outside the scope of SURVEY.md section 8.)
"""
from __future__ import annotations

import dataclasses
import typing as t

import numpy as np

from weatherbench2_b200 import _lib
from weatherbench2_b200 import xarray_lite as xl

EARTH_RADIUS_M = 1000 * (6357 + 6378) / 2  # weatherbench2/schema.py:59


@dataclasses.dataclass
class DerivedVariable:
  """Derived variable base class (derived_variables.py:28-56)."""

  @property
  def base_variables(self) -> list[str]:
    return []

  @property
  def core_dims(self):
    raise NotImplementedError

  @property
  def all_input_core_dims(self) -> set[str]:
    return set().union(*self.core_dims[0])

  def compute(self, dataset):
    raise NotImplementedError


@dataclasses.dataclass
class ZonalEnergySpectrum(DerivedVariable):
  """Energy spectrum along the zonal direction
  (derived_variables.py:531-626).

  S[0] = C |F[0]|^2, S[k] = 2 C |F[k]|^2 for k > 0 (the Nyquist bin included),
  F = rfft(f, norm='forward'), C = circumference of the latitude circle.
  Output dims: the input dims with `longitude` replaced by a trailing
  `zonal_wavenumber`; coords `frequency` (1 / m) and `wavelength` (m) with dims
  (zonal_wavenumber, latitude).
  """

  variable_name: str

  @property
  def base_variables(self) -> list[str]:
    return [self.variable_name]

  @property
  def core_dims(self):
    return (['longitude'],), ['zonal_wavenumber']

  def _circumference(self, latitude: np.ndarray) -> np.ndarray:
    """derived_variables.py:578-581."""
    circum_at_equator = 2 * np.pi * EARTH_RADIUS_M
    return np.cos(np.asarray(latitude) * np.pi / 180) * circum_at_equator

  def lon_spacing_m(self, dataset) -> xl.DataArray:
    """Spacing (meters) between longitudes (derived_variables.py:583-590)."""
    ds = xl.from_xarray(dataset)
    lon = ds['longitude'].values
    lat = ds['latitude'].values
    diffs = np.diff(lon)
    if np.max(np.abs(diffs - diffs[0])) > 1e-3:
      raise ValueError(f'Expected uniform longitude spacing. {lon=}')
    return xl.DataArray(self._circumference(lat) * diffs[0] / 360,
                        ('latitude',), {'latitude': lat})

  def compute(self, dataset, time_sum_dim: t.Optional[str] = None):
    """Zonal power at wavenumber and frequency (derived_variables.py:592-626).

    time_sum_dim (extension): when given, the spectrum is additionally summed
    over that dimension on the device (the script's `xbeam.Mean(['time'])`,
    scripts/compute_zonal_energy_spectrum.py:234, is sum / count); the
    dimension is dropped from the result.
    """
    native = xl.is_native_xarray(dataset)
    ds = xl.from_xarray(dataset)
    spacing = self.lon_spacing_m(ds).values
    da = ds[self.variable_name]
    lat = ds['latitude'].values
    lon = ds['longitude'].values
    if 'latitude' not in da.dims or 'longitude' not in da.dims:
      raise ValueError(f'{self.variable_name} needs latitude and longitude')
    outer = tuple(d for d in da.dims if d not in ('latitude', 'longitude'))
    if time_sum_dim is not None:
      if time_sum_dim not in outer:
        raise ValueError(f'{time_sum_dim!r} is not a dimension of the data')
      outer = (time_sum_dim,) + tuple(d for d in outer if d != time_sum_dim)
    work = da.transpose(*(outer + ('latitude', 'longitude')))
    data = work.data
    ctx = _lib.default_context()
    nlat, nlon = lat.size, lon.size
    nk = nlon // 2 + 1
    oshape = tuple(work.sizes[d] for d in outer)
    nfield = int(np.prod(oshape)) if oshape else 1
    scale = self._circumference(lat)
    nt = oshape[0] if time_sum_dim is not None else 1
    nout = nfield // nt
    res_shape = (oshape[1:] if time_sum_dim is not None else oshape) + (nlat,
                                                                        nk)
    is_torch = xl._is_torch(data)  # pylint: disable=protected-access
    if is_torch and data.is_cuda:
      import torch  # pylint: disable=import-outside-toplevel
      x = data.to(torch.float32).contiguous()
      out = torch.zeros(res_shape, device=data.device, dtype=torch.float32)
      torch.cuda.current_stream(data.device).synchronize()
      ctx.zonal_spectrum(x.data_ptr(), nfield, nlat, nlon, scale,
                         out.data_ptr(), time_sum_dim is not None, nout)
      ctx.synchronize()
      values = out
    else:
      # host data: streamed through double-buffered staging; with a time sum
      # the accumulator stays in HBM and only the sum comes back
      x = np.ascontiguousarray(np.asarray(data), dtype=np.float32)
      values = ctx.pinned_result(res_shape, np.float32)  # overwritten below
      if not nfield:
        values[...] = 0
      if nfield:
        ctx.zonal_spectrum_host(x.ctypes.data, nfield, nlat, nlon, scale,
                                values.ctypes.data, time_sum_dim is not None,
                                nout)
    out_outer = outer[1:] if time_sum_dim is not None else outer
    dims = out_outer + ('latitude', 'zonal_wavenumber')
    base_frequency = np.fft.rfftfreq(nlon)  # derived_variables.py:614
    with np.errstate(divide='ignore'):
      frequency = base_frequency[:, None] / spacing[None, :]
      wavelength = 1 / frequency
    coords = {k: c for k, c in work.coords.items()
              if 'longitude' not in c.dims and all(d in dims for d in c.dims)}
    coords['zonal_wavenumber'] = xl.Coord(('zonal_wavenumber',), np.arange(nk))
    coords['frequency'] = xl.Coord(('zonal_wavenumber', 'latitude'), frequency,
                                   {'units': '1 / m'})
    coords['wavelength'] = xl.Coord(('zonal_wavenumber', 'latitude'),
                                    wavelength, {'units': 'm'})
    # the reference keeps the non-core dims in place and appends the new core
    # dim (apply_ufunc): (..., latitude, ..., zonal_wavenumber)
    ref_dims = tuple(d for d in da.dims if d != 'longitude' and
                     d != time_sum_dim) + ('zonal_wavenumber',)
    result = xl.DataArray(values, dims, coords, self.variable_name)
    if not is_torch:
      result = result.transpose(*ref_dims)
    return xl.to_xarray(result) if native else result


  def compute_latitude_mean(self, dataset, time_mean_dim: t.Optional[str] = None,
                            lat_slice: t.Optional[slice] = None):
    """North-star path (BASELINE.json): rFFT along longitude FOLLOWED BY the
    weighted meridional reduction, fused in one kernel
    (wb2_zonal_spectrum_latsum): the get_lat_weights-weighted latitude mean
    (weatherbench2/metrics.py:40-60) of `compute(dataset)`, optionally over the
    latitude band `lat_slice` (label-inclusive like SliceRegion) and averaged
    over `time_mean_dim`.  The per-latitude spectrum is never written (4 B read
    per cell, ~0 written).  Output dims: the remaining outer dims +
    `zonal_wavenumber`; no frequency / wavelength coordinates (they depend on
    latitude)."""
    from weatherbench2_b200 import _spatial as sp  # pylint: disable=import-outside-toplevel
    native = xl.is_native_xarray(dataset)
    ds = xl.from_xarray(dataset)
    self.lon_spacing_m(ds)  # uniform-longitude check (derived_variables.py:583-590)
    da = ds[self.variable_name]
    lat = ds['latitude'].values
    lon = ds['longitude'].values
    outer = tuple(d for d in da.dims if d not in ('latitude', 'longitude'))
    if time_mean_dim is not None:
      if time_mean_dim not in outer:
        raise ValueError(f'{time_mean_dim!r} is not a dimension of the data')
      outer = (time_mean_dim,) + tuple(d for d in outer if d != time_mean_dim)
    work = da.transpose(*(outer + ('latitude', 'longitude')))
    data = work.data
    nlat, nlon = lat.size, lon.size
    nk = nlon // 2 + 1
    oshape = tuple(work.sizes[d] for d in outer)
    nfield = int(np.prod(oshape)) if oshape else 1
    nt = oshape[0] if time_mean_dim is not None else 1
    nout = nfield // nt
    w = sp.lat_weights(lat)
    if lat_slice is not None:
      lo = -np.inf if lat_slice.start is None else lat_slice.start
      hi = np.inf if lat_slice.stop is None else lat_slice.stop
      w = w * ((lat >= lo) & (lat <= hi))
    if not w.sum() > 0:
      raise ValueError('the latitude band selects no latitude')
    scale = self._circumference(lat) * w / w.sum() / nt
    res_shape = (oshape[1:] if time_mean_dim is not None else oshape) + (nk,)
    ctx = _lib.default_context()
    if xl._is_torch(data) and data.is_cuda:  # pylint: disable=protected-access
      import torch  # pylint: disable=import-outside-toplevel
      x = data.to(torch.float32).contiguous()
      out = torch.empty(res_shape, device=x.device, dtype=torch.float32)
      ctx.zonal_spectrum_latsum(x.data_ptr(), nfield, nlat, nlon, scale,
                                out.data_ptr(), nout)
      values = out
    else:
      x = np.ascontiguousarray(np.asarray(data), dtype=np.float32)
      values = np.empty(res_shape, dtype=np.float32)
      ctx.zonal_spectrum_latsum_host(x.ctypes.data, nfield, nlat, nlon, scale,
                                     values.ctypes.data, nout)
    out_outer = outer[1:] if time_mean_dim is not None else outer
    dims = out_outer + ('zonal_wavenumber',)
    coords = {k: c for k, c in work.coords.items()
              if all(d in out_outer for d in c.dims)}
    coords['zonal_wavenumber'] = xl.Coord(('zonal_wavenumber',), np.arange(nk))
    result = xl.DataArray(values, dims, coords, self.variable_name)
    return xl.to_xarray(result) if native else result


def interpolate_spectral_frequencies(spectrum, wavenumber_dim: str,
                                     frequencies=None, method: str = 'linear'):
  """Interpolate the frequencies of `spectrum` (as produced by
  ZonalEnergySpectrum.compute) to common values
  (weatherbench2/derived_variables.py:629-682): every latitude row lives on its
  own frequency axis; the result replaces `wavenumber_dim` by `frequency`, NaN
  where a common frequency lies outside a row's range.  Default frequencies:
  the narrowest range present, `spectrum.sizes[wavenumber_dim]` points
  (:658-664).  Linear interpolation only (the kernel is a gather + lerp)."""
  if method != 'linear':
    raise NotImplementedError('only method="linear" runs on the device')
  native = xl.is_native_xarray(spectrum)
  sp_ = xl.from_xarray(spectrum)
  freq = sp_.coords['frequency']
  if set(freq.dims) != {wavenumber_dim, 'latitude'}:
    raise ValueError(f'{freq.dims=} was not a permutation of '
                     f'("{wavenumber_dim}", "latitude")')
  fv = np.asarray(freq.values, dtype=np.float64)
  if freq.dims[0] != wavenumber_dim:
    fv = fv.T  # (wavenumber, latitude)
  nk, nlat = fv.shape
  if frequencies is None:
    freq_min = fv.max(axis=1).min()
    freq_max = fv.min(axis=1).max()
    frequencies = np.linspace(freq_min, freq_max, num=nk)
  frequencies = np.asarray(frequencies, dtype=np.float64)
  if frequencies.ndim != 1:
    raise ValueError(f'Expected 1-D frequencies, found {frequencies.shape=}')
  table = np.ascontiguousarray(fv.T)  # (latitude, wavenumber), increasing in k
  if nk < 2 or not (np.diff(table, axis=1) > 0).all():
    raise ValueError('the frequency coordinate must increase with wavenumber')
  outer = tuple(d for d in sp_.dims if d not in ('latitude', wavenumber_dim))
  work = sp_.transpose(*(outer + ('latitude', wavenumber_dim)))
  data = work.data
  oshape = tuple(work.sizes[d] for d in outer)
  nfield = int(np.prod(oshape)) if oshape else 1
  nf = frequencies.size
  ctx = _lib.default_context()
  if xl._is_torch(data) and data.is_cuda:  # pylint: disable=protected-access
    import torch  # pylint: disable=import-outside-toplevel
    x = data.to(torch.float32).contiguous()
    out = torch.empty(oshape + (nlat, nf), device=x.device, dtype=torch.float32)
    ctx.spectrum_interp(x.data_ptr(), nfield, nlat, nk, table, frequencies,
                        out.data_ptr())
    values = out
  else:
    x = np.ascontiguousarray(np.asarray(data), dtype=np.float32)
    src = ctx.to_device(x)
    dst = ctx.malloc(max(4, nfield * nlat * nf * 4))
    try:
      ctx.spectrum_interp(src, nfield, nlat, nk, table, frequencies, dst)
      values = ctx.from_device(dst, oshape + (nlat, nf), np.float32)
    finally:
      ctx.free(src)
      ctx.free(dst)
  dims = outer + ('latitude', 'frequency')
  coords = {k: c for k, c in work.coords.items()
            if wavenumber_dim not in c.dims and k not in ('frequency',
                                                           'wavelength')}
  coords['frequency'] = xl.Coord(('frequency',), frequencies)
  with np.errstate(divide='ignore'):
    # "Interp didn't deal well with the infinite wavelength, so just reset"
    coords['wavelength'] = xl.Coord(('frequency',), 1 / frequencies,
                                    {'units': 'm'})
  result = xl.DataArray(values, dims, coords, sp_.name, sp_.attrs)
  # the reference's groupby('latitude').apply keeps latitude where it was and
  # puts frequency in the wavenumber's place
  ref_dims = tuple('frequency' if d == wavenumber_dim else d for d in sp_.dims)
  if not xl._is_torch(values):  # pylint: disable=protected-access
    result = result.transpose(*ref_dims)
  return xl.to_xarray(result) if native else result


# Wind variables live in _derived_wind.py (they need DerivedVariable).
from weatherbench2_b200._derived_wind import WindSpeed  # noqa: E402  pylint: disable=wrong-import-position

# The named derived variables of weatherbench2/derived_variables.py:724-773 that
# have a device implementation.  The remaining entries of the reference's table
# (divergence / vorticity, geostrophic winds, lapse rate, column integrals,
# relative humidity, precipitation accumulations) are data preparation off the
# hot path and are not reproduced; asking for one raises KeyError here instead
# of silently running on the host.
DERIVED_VARIABLE_DICT = {
    'wind_speed': WindSpeed(u_name='u_component_of_wind',
                            v_name='v_component_of_wind'),
    '10m_wind_speed': WindSpeed(u_name='10m_u_component_of_wind',
                                v_name='10m_v_component_of_wind'),
}
