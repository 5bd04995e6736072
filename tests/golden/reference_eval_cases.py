"""`evaluate_in_memory` runs whose result files the REFERENCE's own
weatherbench2/evaluation.py wrote in this container
(tests/golden/make_reference_vectors.py; xarray / apache_beam / xarray_beam /
zarr replaced by the stand-ins of tests/golden/xarray_shim, datasets handed
over through an in-memory "zarr store" registry).

`build(lib, source)` returns [(case name, config.Data, {eval name: config.Eval})]
for `lib` = the reference's modules or the product's; `source(name, dataset)`
turns a dataset into whatever `config.Paths` takes (a registered path for the
reference, the dataset itself for the product).
"""
import numpy as np

H = np.timedelta64(1, 'h').astype('timedelta64[ns]')
NLAT, NLON = 9, 12
LAT = np.linspace(-90, 90, NLAT)
LON = np.linspace(0, 360, NLON, endpoint=False)
LEVELS = (500, 700, 850)
Z = 'geopotential'


def _times(start, stop, step_h):
  return np.arange(np.datetime64(start, 'ns'), np.datetime64(stop, 'ns'),
                   step_h * H)


def datasets(Dataset):  # pylint: disable=invalid-name
  """Seeded inputs in WeatherBench2's on-disk conventions."""
  rs = np.random.RandomState(802701)
  coords = dict(level=np.array(LEVELS), latitude=LAT, longitude=LON)
  grid = (len(LEVELS), NLON, NLAT)
  gdims = ('level', 'longitude', 'latitude')
  out = {}
  ttimes = _times('2018-01-01', '2021-01-12', 6)
  out['truth'] = Dataset(
      {Z: (('time',) + gdims,
           rs.standard_normal((ttimes.size,) + grid).astype(np.float32))},
      dict(coords, time=ttimes))
  lead = np.array([0, 6, 12, 30]) * H
  init = _times('2020-12-27', '2021-01-04', 12)
  fdims = ('time', 'prediction_timedelta') + gdims
  out['by_init'] = Dataset(
      {Z: (fdims, rs.standard_normal((init.size, lead.size) + grid).astype(
          np.float32))},
      dict(coords, time=init, prediction_timedelta=lead))
  elead = np.array([0, 6, 12, 18, 24]) * H  # evenly spaced: analysis pairing
  out['by_init_even'] = Dataset(
      {Z: (fdims, rs.standard_normal((init.size, elead.size) + grid).astype(
          np.float32))},
      dict(coords, time=init, prediction_timedelta=elead))
  vlead = np.array([0, 12, 24]) * H
  valid = _times('2020-12-27', '2021-01-04', 12)
  out['by_valid'] = Dataset(
      {Z: (fdims, rs.standard_normal((valid.size, vlead.size) + grid).astype(
          np.float32))},
      dict(coords, time=valid, prediction_timedelta=vlead))
  # the same by-valid forecast stored with pressure-level suffixes
  a = out['by_valid'][Z].values
  out['by_valid_suffixed'] = Dataset(
      {f'{Z}_{lev}': (fdims[:2] + gdims[1:], a[:, :, i])
       for i, lev in [(2, 850), (0, 500), (1, 700)]},
      dict(latitude=LAT, longitude=LON, time=valid,
           prediction_timedelta=vlead))
  cdims = ('hour', 'dayofyear') + gdims
  out['clim'] = Dataset(
      {Z: (cdims, rs.standard_normal((4, 366) + grid).astype(np.float32))},
      dict(coords, hour=np.array([0, 6, 12, 18]),
           dayofyear=np.arange(1, 367)))
  return out


def build(lib, source, output_dir):
  cfg, m, r = lib.config, lib.metrics, lib.regions
  ds = datasets(lib.Dataset)
  regions = {'global': r.SliceRegion(),
             'tropics': r.SliceRegion(lat_slice=slice(-20, 20)),
             'extra-tropics': r.ExtraTropicalRegion()}

  def data(forecast, by_init, **sel):
    sel.setdefault('variables', [Z])
    sel.setdefault('time_slice', slice('2020-12-28', '2021-01-02'))
    kw = {}
    if forecast == 'by_valid_suffixed':
      kw['pressure_level_suffixes'] = True
    return cfg.Data(
        selection=cfg.Selection(**sel),
        paths=cfg.Paths(forecast=source(forecast, ds[forecast]),
                        obs=source('truth', ds['truth']),
                        climatology=source('clim', ds['clim']),
                        output_dir=output_dir), by_init=by_init, **kw)

  rmse = {'rmse': m.RMSESqrtBeforeTimeAvg(), 'mse': m.MSE()}
  acc = {'acc': m.ACC(climatology=ds['clim']), 'bias': m.Bias()}
  return [
      ('by_init', data('by_init', True, levels=[500, 850]), {
          'plain': cfg.Eval(metrics=dict(rmse, **acc)),
          'by_region': cfg.Eval(metrics=rmse, regions=regions),
          'temporal': cfg.Eval(metrics=rmse, temporal_mean=False),
          'climatology': cfg.Eval(metrics=rmse, evaluate_climatology=True),
      }),
      # (2018 / 2019 hold no day 366: valid times on 2020-12-31 would be a
      # KeyError in the reference as well, so this case starts in January)
      ('by_init_january', data(
          'by_init', True, time_slice=slice('2021-01-01', '2021-01-02')), {
              'probabilistic_climatology': cfg.Eval(
                  metrics={'crps': m.CRPS(ensemble_dim='number'),
                           'ensemble_mean_mse': m.EnsembleMeanMSE(
                               ensemble_dim='number')},
                  evaluate_probabilistic_climatology=True,
                  probabilistic_climatology_start_year=2018,
                  probabilistic_climatology_end_year=2019,
                  probabilistic_climatology_hour_interval=6)}),
      ('by_init_analysis', data(
          'by_init_even', True, time_slice=slice('2020-12-28', '2020-12-31')), {
              'vs_analysis': cfg.Eval(metrics=rmse, against_analysis=True)}),
      ('by_valid', data('by_valid', False), {
          'plain': cfg.Eval(metrics=rmse, regions=regions),
          'persistence': cfg.Eval(metrics=rmse, evaluate_persistence=True),
          'climatology': cfg.Eval(metrics=rmse, evaluate_climatology=True),
          'vs_analysis': cfg.Eval(metrics=rmse, against_analysis=True),
      }),
      ('by_valid_box', data(
          'by_valid_suffixed', False, levels=[500, 850],
          lat_slice=slice(-45, 60), lon_slice=slice(30, 200)), {
              'plain': cfg.Eval(metrics=rmse)}),
  ]
