"""Metric calls whose results the REFERENCE's own code produced in this
container (tests/golden/make_reference_vectors.py -> reference_run_vectors.npz).

The same case table drives three runs:
  * the generator: `lib` = the reference's weatherbench2.{metrics, regions,
    thresholds}, executed on a stand-in xarray (tests/golden/xarray_shim);
  * tests/test_reference_run_vectors.py: the oracle, restating each call;
  * the product (weatherbench2_b200) on the NumPy stand-in context (CPU) and on
    the CUDA kernels (GPU) -- same classes, same arguments as the reference.
Inputs are regenerated from seeds; only the reference's outputs are stored.
"""
import numpy as np

NLAT, NLON = 7, 12
LAT = np.linspace(-90, 90, NLAT)
LON = np.linspace(0, 360, NLON, endpoint=False)
LEVELS = np.array([500, 850])
# four 12-hourly times across the 2020 -> 2021 boundary (day 366, day 1)
TIMES = (np.datetime64('2020-12-30T12', 'ns') +
         np.arange(5) * np.timedelta64(12, 'h'))
DIMS = ('time', 'level', 'longitude', 'latitude')
Z, U, V = 'geopotential', 'u_component_of_wind', 'v_component_of_wind'
QUANTILES = np.array([0.1, 0.5, 0.9])


def arrays() -> dict:
  """Every input array, float32, from fixed seeds."""
  rs = np.random.RandomState(20240607)
  shape = (TIMES.size, LEVELS.size, NLON, NLAT)
  out = {}
  for name in ('truth', 'det'):
    for var in (Z, U, V):
      out[f'{name}/{var}'] = rs.standard_normal(shape).astype(np.float32)
  nan = out['det/' + Z].copy()
  nan[rs.uniform(size=shape) < 0.05] = np.nan
  out['det_nan/' + Z] = nan
  nan = out['truth/' + Z].copy()
  nan[rs.uniform(size=shape) < 0.03] = np.nan
  out['truth_nan/' + Z] = nan
  for m in (1, 2, 5):
    signal = rs.standard_normal(shape)
    out[f'ens{m}/{Z}'] = (signal[None] + rs.standard_normal((m,) + shape)
                          ).astype(np.float32)
  e = out['ens5/' + Z].copy()
  e[rs.uniform(size=e.shape) < 0.05] = np.nan
  out['ens5_nan/' + Z] = e
  out['gauss/' + Z] = rs.standard_normal(shape).astype(np.float32)
  out['gauss/' + Z + '_std'] = rs.uniform(0.5, 1.5, shape).astype(np.float32)
  cshape = (2, 366) + shape[1:]
  out['clim/' + Z] = (0.3 * rs.standard_normal(cshape)).astype(np.float32)
  out['clim/' + Z + '_std'] = rs.uniform(0.5, 1.5, cshape).astype(np.float32)
  q = np.sort(rs.standard_normal((QUANTILES.size,) + cshape), axis=0)
  out['clim/' + Z + '_quantile'] = q.astype(np.float32)
  out['lsm'] = (rs.uniform(size=(NLAT, NLON)) > 0.5).astype(np.float32)
  out['lsm_frac'] = rs.uniform(size=(NLAT, NLON)).astype(np.float32)
  return out


def datasets(Dataset, arr: dict) -> dict:  # pylint: disable=invalid-name
  coords = dict(time=TIMES, level=LEVELS, longitude=LON, latitude=LAT)
  out = {}
  for name in ('truth', 'det'):
    out[name] = Dataset({v: (DIMS, arr[f'{name}/{v}']) for v in (Z, U, V)},
                        coords)
  for name in ('det_nan', 'truth_nan'):
    out[name] = Dataset({Z: (DIMS, arr[f'{name}/{Z}'])}, coords)
  out['truth_z'] = Dataset({Z: (DIMS, arr['truth/' + Z])}, coords)
  out['det_z'] = Dataset({Z: (DIMS, arr['det/' + Z])}, coords)
  for name in ('ens1', 'ens2', 'ens5', 'ens5_nan'):
    a = arr[f'{name}/{Z}']
    out[name] = Dataset({Z: (('realization',) + DIMS, a)},
                        dict(coords, realization=np.arange(a.shape[0])))
  out['gauss'] = Dataset({Z: (DIMS, arr['gauss/' + Z]),
                          Z + '_std': (DIMS, arr['gauss/' + Z + '_std'])},
                         coords)
  ccoords = dict(hour=np.array([0, 12]), dayofyear=np.arange(1, 367),
                 level=LEVELS, longitude=LON, latitude=LAT)
  cdims = ('hour', 'dayofyear') + DIMS[1:]
  out['clim'] = Dataset(
      {Z: (cdims, arr['clim/' + Z]),
       Z + '_std': (cdims, arr['clim/' + Z + '_std']),
       Z + '_quantile': (('quantile',) + cdims, arr['clim/' + Z + '_quantile'])},
      dict(ccoords, quantile=QUANTILES))
  return out


def _slices(spec):
  if spec is None:
    return slice(None, None)
  got = [slice(a, b) for a, b in spec]
  return got[0] if len(got) == 1 else got


def build_region(lib, spec, arr, DataArray):  # pylint: disable=invalid-name
  if spec is None:
    return None
  kind = spec['type']
  if kind == 'SliceRegion':
    return lib.regions.SliceRegion(lat_slice=_slices(spec.get('lat')),
                                   lon_slice=_slices(spec.get('lon')))
  if kind == 'ExtraTropicalRegion':
    return lib.regions.ExtraTropicalRegion()
  if kind == 'LandRegion':
    lsm = DataArray(arr[spec['mask']], ('latitude', 'longitude'),
                    {'latitude': LAT, 'longitude': LON})
    return lib.regions.LandRegion(land_sea_mask=lsm,
                                  threshold=spec.get('threshold'))
  if kind == 'CombinedRegion':
    return lib.regions.CombinedRegion(regions=[
        build_region(lib, s, arr, DataArray) for s in spec['regions']])
  raise ValueError(kind)


def build_metric(lib, case, ds):
  kwargs = dict(case.get('kwargs', {}))
  if kwargs.pop('climatology', False):
    kwargs['climatology'] = ds['clim']
  if 'thresholds' in kwargs:
    cls = getattr(lib.thresholds, kwargs.pop('thresholds'))
    kwargs['thresholds'] = [cls(climatology=ds['clim'], quantile=float(q))
                            for q in QUANTILES]
  if 'wind_vector' in kwargs:
    cls = getattr(lib.metrics, kwargs.pop('wind_vector'))
    inner = cls(u_name=U, v_name=V, vector_name='wind_vector')
    key = 'wind_vector_mse' if case['metric'] == 'MSE' else 'wind_vector_rmse'
    kwargs[key] = [inner]
  return getattr(lib.metrics, case['metric'])(**kwargs)


def run_case(lib, case, ds, arr, DataArray):  # pylint: disable=invalid-name
  """{variable: (dims, values)} of one metric call."""
  metric = build_metric(lib, case, ds)
  region = build_region(lib, case.get('region'), arr, DataArray)
  fn = getattr(metric, case.get('method', 'compute_chunk'))
  kw = dict(skipna=case.get('skipna', False))
  if region is not None:
    kw['region'] = region
  result = fn(ds[case['forecast']], ds[case['truth']], **kw)
  return {str(k): (tuple(result[k].dims), np.asarray(result[k].values))
          for k in result.keys()}


EUROPE = {'type': 'SliceRegion', 'lat': [(35, 75)],
          'lon': [(360 - 12.5, None), (0, 42.5)]}
TROPICS = {'type': 'SliceRegion', 'lat': [(-20, 20)]}
EXTRA = {'type': 'ExtraTropicalRegion'}
LAND = {'type': 'LandRegion', 'mask': 'lsm'}
LAND_THR = {'type': 'LandRegion', 'mask': 'lsm_frac', 'threshold': 0.4}
TROPICS_LAND = {'type': 'CombinedRegion', 'regions': [TROPICS, LAND]}
POLES = {'type': 'SliceRegion', 'lat': [(None, -60), (60, None)]}


def _cases():
  out = []

  def add(metric, forecast, truth, **kw):
    case = dict(metric=metric, forecast=forecast, truth=truth, **kw)
    tag = [metric, forecast, case.get('method', 'compute_chunk')]
    if case.get('region'):
      r = case['region']
      tag.append(next(k for k, v in globals().items() if v is r))
    if case.get('skipna'):
      tag.append('skipna')
    if case.get('kwargs'):
      tag.append('-'.join(f'{k}={v}' for k, v in case['kwargs'].items()))
    case['id'] = '/'.join(tag)
    out.append(case)

  # deterministic family (weatherbench2/metrics.py:175-414)
  for metric in ('MSE', 'MAE', 'Bias', 'RMSESqrtBeforeTimeAvg'):
    add(metric, 'det', 'truth')
    add(metric, 'det', 'truth', method='compute')
    for region in (EUROPE, TROPICS, EXTRA, LAND, LAND_THR, TROPICS_LAND, POLES):
      add(metric, 'det', 'truth', region=region)
    for skipna in (False, True):
      add(metric, 'det_nan', 'truth_nan', skipna=skipna)
      add(metric, 'det_nan', 'truth_nan', skipna=skipna, region=EUROPE)
  add('MSE', 'det', 'truth', kwargs={'wind_vector': 'WindVectorMSE'})
  add('RMSESqrtBeforeTimeAvg', 'det', 'truth', region=TROPICS,
      kwargs={'wind_vector': 'WindVectorRMSESqrtBeforeTimeAvg'})
  for region in (None, EUROPE, LAND):
    add('ACC', 'det_z', 'truth_z', region=region,
        kwargs={'climatology': True})
  for skipna in (False, True):
    add('ACC', 'det_nan', 'truth_nan', skipna=skipna,
        kwargs={'climatology': True})
  for metric in ('SpatialMSE', 'SpatialMAE', 'SpatialBias'):
    add(metric, 'det', 'truth')
    add(metric, 'det_nan', 'truth_nan', method='compute', skipna=True)
  # ensemble family (metrics.py:598-846, 1185-1517)
  ens_metrics = ('CRPS', 'CRPSSkill', 'CRPSSpread', 'EnsembleMeanMSE',
                 'EnsembleMeanRMSESqrtBeforeTimeAvg', 'EnsembleVariance',
                 'EnsembleStddevSqrtBeforeTimeAvg', 'DebiasedEnsembleMeanMSE',
                 'EnergyScore', 'EnergyScoreSkill', 'EnergyScoreSpread')
  for metric in ens_metrics:
    for name in ('ens1', 'ens2', 'ens5'):
      add(metric, name, 'truth_z')
    add(metric, 'ens5', 'truth_z', region=TROPICS_LAND)
    add(metric, 'ens5', 'truth_z', method='compute')
    if not metric.startswith('Energy'):
      add(metric, 'ens5_nan', 'truth_z', skipna=True)
  for metric in ('SpatialCRPS', 'SpatialCRPSSkill', 'SpatialCRPSSpread',
                 'SpatialEnsembleVariance', 'SpatialEnsembleMeanMSE',
                 'DebiasedSpatialEnsembleMeanMSE'):
    add(metric, 'ens5', 'truth_z')
    add(metric, 'ens2', 'truth_z', method='compute')
  # Gaussian forecasts (metrics.py:849-1158)
  add('GaussianCRPS', 'gauss', 'truth_z')
  add('GaussianCRPS', 'gauss', 'truth_z', region=EUROPE)
  add('GaussianVariance', 'gauss', 'truth_z')
  for metric in ('GaussianBrierScore', 'GaussianIgnoranceScore', 'GaussianRPS'):
    for thr in ('GaussianQuantileThreshold', 'QuantileThreshold'):
      add(metric, 'gauss', 'truth_z', kwargs={'thresholds': thr})
  # threshold metrics on ensembles (metrics.py:1523-1891)
  for metric in ('EnsembleBrierScore', 'DebiasedEnsembleBrierScore',
                 'EnsembleIgnoranceScore', 'EnsembleRPS'):
    for thr in ('GaussianQuantileThreshold', 'QuantileThreshold'):
      add(metric, 'ens5', 'truth_z', kwargs={'thresholds': thr})
    add(metric, 'ens5', 'truth_z', region=TROPICS,
        kwargs={'thresholds': 'GaussianQuantileThreshold'})
    add(metric, 'ens5_nan', 'truth_z', skipna=True,
        kwargs={'thresholds': 'GaussianQuantileThreshold'})
  for metric in ('SpatialEnsembleBrierScore',
                 'SpatialDebiasedEnsembleBrierScore',
                 'SpatialEnsembleIgnoranceScore', 'SpatialEnsembleRPS'):
    add(metric, 'ens5', 'truth_z',
        kwargs={'thresholds': 'GaussianQuantileThreshold'})
  return out


CASES = _cases()


# ------------------------------------------------------------------------------
# Calls with other shapes: rank histogram + central reliability, SEEPS, zonal
# energy spectrum + frequency interpolation, wind speed
# ------------------------------------------------------------------------------
PRECIP = 'total_precipitation_24hr'


def extra_inputs():
  rs = np.random.RandomState(424242)
  h = np.timedelta64(1, 'h').astype('timedelta64[ns]')
  init = TIMES[:3]
  lead = np.array([0, 12]) * h
  shape = (init.size, lead.size, NLON, NLAT)
  wet = lambda: np.where(rs.uniform(size=shape) < 0.3, 0.0,  # noqa: E731
                         rs.gamma(0.6, 0.002, size=shape)).astype(np.float32)
  pf, pt = wet(), wet()
  pf.flat[::17] = np.float32(0.00025)   # exactly AT the dry threshold
  pf.flat[5::41] = np.nan
  cshape = (2, 366, NLON, NLAT)
  return dict(
      init=init, lead=lead, pf=pf, pt=pt,
      dry=rs.uniform(0.0, 1.0, cshape).astype(np.float32),
      thr=rs.uniform(0.0005, 0.004, cshape).astype(np.float32),
      field=(rs.standard_normal((3, 2, NLAT, 36)) + 0.3).astype(np.float32),
      hist=np.array([0.2, 0.05, 0.1, 0.1, 0.25, 0.3]))


def run_extras(lib, Dataset, arr):  # pylint: disable=invalid-name
  """{name: (dims, values)}; `lib` has metrics and derived_variables."""
  x = extra_inputs()
  ds = datasets(Dataset, arr)
  out = {}

  def put(name, da):
    out[name] = (tuple(da.dims), np.asarray(da.values))

  for bins in (None, 3):
    r = lib.metrics.RankHistogram(
        ensemble_dim='realization', num_bins=bins).compute_chunk(
            ds['ens5'], ds['truth_z'])[Z]
    put(f'rank_histogram/bins={bins}', r)
  hist = Dataset({Z: (('bins',), x['hist'])}, {'bins': np.arange(6)})
  rel = lib.metrics.central_reliability(hist)[Z]
  put('central_reliability', rel)
  put('central_reliability/desired_prob', rel['desired_prob'])
  # SEEPS on a by-init forecast (needs the valid_time coordinate)
  pdims = ('init_time', 'lead_time', 'longitude', 'latitude')
  pcoords = dict(init_time=x['init'], lead_time=x['lead'], longitude=LON,
                 latitude=LAT,
                 valid_time=(('init_time', 'lead_time'),
                             x['init'][:, None] + x['lead'][None, :]))
  cdims = ('hour', 'dayofyear', 'longitude', 'latitude')
  clim = Dataset({PRECIP + '_seeps_dry_fraction': (cdims, x['dry']),
                  PRECIP + '_seeps_threshold': (cdims, x['thr'])},
                 dict(hour=np.array([0, 12]), dayofyear=np.arange(1, 367),
                      longitude=LON, latitude=LAT))
  pfd = Dataset({PRECIP: (pdims, x['pf'])}, pcoords)
  ptd = Dataset({PRECIP: (pdims, x['pt'])}, pcoords)
  put('seeps', lib.metrics.SEEPS(climatology=clim).compute_chunk(
      pfd, ptd)[PRECIP])
  put('spatial_seeps', lib.metrics.SpatialSEEPS(
      climatology=clim, min_p1=0.3, max_p1=0.7).compute_chunk(pfd, ptd)[PRECIP])
  # zonal energy spectrum of a (time, level, latitude, longitude) field
  lon36 = np.linspace(0, 360, 36, endpoint=False)
  sds = Dataset({U: (('time', 'level', 'latitude', 'longitude'), x['field']),
                 V: (('time', 'level', 'latitude', 'longitude'),
                     x['field'][::-1].copy())},
                dict(time=TIMES[:3], level=LEVELS, latitude=LAT,
                     longitude=lon36))
  spec = lib.derived_variables.ZonalEnergySpectrum(U).compute(sds)
  put('spectrum', spec)
  put('spectrum/frequency', spec['frequency'])
  put('spectrum/wavelength', spec['wavelength'])
  inner = spec.isel(latitude=slice(1, NLAT - 1))  # the poles have zero length
  interp = lib.derived_variables.interpolate_spectral_frequencies(
      inner, 'zonal_wavenumber')
  put('spectrum_interp', interp)
  put('spectrum_interp/frequency', interp['frequency'])
  put('wind_speed', lib.derived_variables.WindSpeed(
      u_name=U, v_name=V).compute(sds))
  if hasattr(lib, 'regridding'):
    out.update(run_regridders(lib.regridding, Dataset))
  if hasattr(lib, 'regions'):
    out.update(run_region_index_sets(lib.regions, Dataset))
  return out


REGION_SLICES = {
    'tropics': dict(lat=[(-20, 20)]),
    'extra_tropics': dict(lat=[(None, -20), (20, None)]),
    'europe_wrap': dict(lat=[(35, 75)], lon=[(360 - 12.5, None), (0, 42.5)]),
    'on_grid_points': dict(lat=[(-30, 60)], lon=[(30, 200)]),
    'between_points': dict(lat=[(-27.3, 61.2)], lon=[(33.3, 196.1)]),
    'overlapping': dict(lat=[(-30, 10), (0, 40)]),  # duplicates are kept
    'single_point': dict(lat=[(0, 0)], lon=[(180, 180)]),
    'reversed_bounds': dict(lat=[(20, -20)]),       # selects nothing
    'outside': dict(lon=[(400, 500)]),              # selects nothing
}


def run_region_index_sets(regions, Dataset):  # pylint: disable=invalid-name
  """SliceRegion.apply (regions.py:57-98) on a 10-degree grid: which rows /
  columns each label slice selects (both ends inclusive, list-of-slices
  concatenated without de-duplication), as coordinate labels and weights."""
  lat = np.linspace(-90, 90, 19)
  lon = np.linspace(0, 360, 36, endpoint=False)
  data = np.arange(19 * 36, dtype=np.float32).reshape(19, 36)
  ds = Dataset({Z: (('latitude', 'longitude'), data)},
               dict(latitude=lat, longitude=lon))
  weights = ds[Z] * 0 + 1
  out = {}
  for name, spec in REGION_SLICES.items():
    region = regions.SliceRegion(lat_slice=_slices(spec.get('lat')),
                                 lon_slice=_slices(spec.get('lon')))
    sub, w = region.apply(ds, weights)
    out[f'region/{name}/latitude'] = (('latitude',),
                                      np.asarray(sub['latitude'].values))
    out[f'region/{name}/longitude'] = (('longitude',),
                                       np.asarray(sub['longitude'].values))
    out[f'region/{name}/data'] = (tuple(sub[Z].dims),
                                  np.asarray(sub[Z].values))
    out[f'region/{name}/weights'] = (tuple(w.dims), np.asarray(w.values))
  return out


def run_regridders(rg, Dataset):  # pylint: disable=invalid-name
  """The three regridders of weatherbench2/regridding.py on a field with a NaN
  patch: global -> global, and source / target grids without poles or without
  periodic longitudes (uncovered target cells come out NaN)."""
  rs = np.random.RandomState(31337)
  out = {}
  fdims = ('field', 'longitude', 'latitude')
  slon = np.linspace(0, 360, 24, endpoint=False)
  slat = np.linspace(-90, 90, 13)
  x = rs.standard_normal((3, 24, 13)).astype(np.float32)
  x[1, 3:6, 4:7] = np.nan
  x[2, :2, :] = np.nan  # across the longitude seam
  tlon = np.linspace(0, 360, 10, endpoint=False)
  tlat = np.linspace(-90, 90, 7)
  grids = {
      'global': (rg.Grid.from_degrees(slon, slat),
                 rg.Grid.from_degrees(tlon, tlat)),
      'no_poles': (rg.Grid(longitudes=slon, latitudes=slat[1:-1],
                           periodic=True, includes_poles=False),
                   rg.Grid(longitudes=tlon, latitudes=tlat, periodic=True,
                           includes_poles=True)),
      'limited_area': (rg.Grid(longitudes=slon[4:16], latitudes=slat[2:10],
                               periodic=False, includes_poles=False),
                       rg.Grid(longitudes=tlon[1:6], latitudes=tlat[1:5],
                               periodic=False, includes_poles=False)),
  }
  crop = {'global': x, 'no_poles': x[:, :, 1:-1],
          'limited_area': x[:, 4:16, 2:10]}
  for gname, (source, target) in grids.items():
    for cls in ('ConservativeRegridder', 'BilinearRegridder',
                'NearestRegridder'):
      r = getattr(rg, cls)(source, target)
      out[f'regrid/{cls}/{gname}'] = (
          fdims, np.asarray(r.regrid_array(crop[gname]), dtype=np.float32))
  # regrid_dataset: dims (time, latitude, longitude), latitude decreasing
  ds = Dataset({Z: (('time', 'latitude', 'longitude'),
                    np.ascontiguousarray(np.transpose(x, (0, 2, 1))[:, ::-1]))},
               dict(time=np.arange(3), latitude=slat[::-1], longitude=slon))
  source, target = grids['global']
  res = rg.ConservativeRegridder(source, target).regrid_dataset(ds)[Z]
  out['regrid_dataset'] = (tuple(res.dims), np.asarray(res.values,
                                                        dtype=np.float32))
  return out
