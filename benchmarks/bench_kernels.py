#!/usr/bin/env python
"""Per-kernel micro-benchmarks (device-resident synthetic inputs, CUDA events
on the launching stream, >= 3 warm-up launches, inputs >> L2).  One JSON line
per kernel with the achieved algorithmic HBM GB/s and the fraction of the
measured peak (MEASURED_PEAKS.json).  Used for profiles/ and DESIGN.md; the
headline contract line is bench.py.

  python benchmarks/bench_kernels.py --kernel ens      # K2, BASELINE configs[2]
  python benchmarks/bench_kernels.py --kernel det      # K1 variants
  python benchmarks/bench_kernels.py --kernel regrid   # K5, configs[3]
  python benchmarks/bench_kernels.py --kernel spectrum # K4, configs[4]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NLAT, NLON = 721, 1440


def peak():
  try:
    return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))[
        'hbm_gbs'])
  except Exception:  # pylint: disable=broad-except
    return 6650.0


def timeit(fn, steps, warmup=3):
  import torch
  for _ in range(warmup):
    fn()
  torch.cuda.synchronize()
  e0 = torch.cuda.Event(enable_timing=True)
  e1 = torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(steps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / steps


def setup():
  import torch
  from weatherbench2_b200 import _lib
  torch.cuda.set_device(0)
  stream = torch.cuda.Stream()
  torch.cuda.set_stream(stream)
  ctx = _lib.Context(0)
  ctx.set_stream(stream.cuda_stream)
  return torch, _lib, ctx


def report(name, ms, nbytes, units, unit_name, extra=None):
  gbs = nbytes / (ms * 1e-3) / 1e9
  line = {'kernel': name, 'ms': ms, 'algorithmic_GB': nbytes / 1e9,
          'achieved_GBps': gbs, 'peak_GBps': peak(), 'frac': gbs / peak(),
          unit_name + '_per_s': units / (ms * 1e-3)}
  line.update(extra or {})
  print(json.dumps(line), flush=True)


def bench_ens(args):
  torch, _lib, ctx = setup()
  from weatherbench2_b200 import _spatial as sp
  lat = np.linspace(-90, 90, NLAT)
  lon = np.arange(NLON) * 0.25
  for m in args.members:
    nfield = args.fields
    x = torch.randn((m, nfield, NLAT, NLON), device='cuda',
                    dtype=torch.float32)
    x += torch.randn((1, nfield, NLAT, NLON), device='cuda')  # shared signal
    t = torch.randn((nfield, NLAT, NLON), device='cuda', dtype=torch.float32)
    (_, spec), = sp.build_weights(ctx, lat, lon, [None], 'lat_lon', NLON)
    slab = NLAT * NLON
    base = min(x.data_ptr(), t.data_ptr())
    off_x = np.arange(nfield, dtype=np.int64) * slab + (x.data_ptr() -
                                                        base) // 4
    off_t = np.arange(nfield, dtype=np.int64) * slab + (t.data_ptr() -
                                                        base) // 4
    out = torch.zeros((nfield, _lib.ENS_NSTAT), device='cuda',
                      dtype=torch.float64)
    for skipna in (False, True):
      fn = lambda: ctx.ens_metrics(base, base, _lib.F32, m, nfield * slab,
                                   off_x, off_t, spec, skipna, out.data_ptr())
      ms = timeit(fn, args.steps)
      pts = nfield * slab
      report(f'ens_metrics M={m} skipna={skipna}', ms, pts * (4 * m + 4), pts,
             'grid_points', {'members': m, 'fields': nfield})
    del x, t


def bench_energy(args):
  torch, _lib, ctx = setup()
  from weatherbench2_b200 import _spatial as sp
  lat = np.linspace(-90, 90, NLAT)
  lon = np.arange(NLON) * 0.25
  for m in args.members:
    nfield = args.fields
    x = torch.randn((m, nfield, NLAT, NLON), device='cuda',
                    dtype=torch.float32)
    t = torch.randn((nfield, NLAT, NLON), device='cuda', dtype=torch.float32)
    (_, spec), = sp.build_weights(ctx, lat, lon, [None], 'lat_lon', NLON)
    slab = NLAT * NLON
    base = min(x.data_ptr(), t.data_ptr())
    off_x = np.arange(nfield, dtype=np.int64) * slab + (x.data_ptr() -
                                                        base) // 4
    off_t = np.arange(nfield, dtype=np.int64) * slab + (t.data_ptr() -
                                                        base) // 4
    out = torch.zeros((nfield, 1, 4, m), device='cuda', dtype=torch.float64)
    fn = lambda: ctx.energy_score(base, base, _lib.F32, m, nfield * slab,
                                  off_x, off_t, spec, out.data_ptr())
    ms = timeit(fn, args.steps)
    pts = nfield * slab
    report(f'energy_score M={m}', ms, pts * (4 * m + 4), pts, 'grid_points',
           {'members': m, 'fields': nfield})
    del x, t


def bench_det(args):
  torch, _lib, ctx = setup()
  from weatherbench2_b200 import _spatial as sp, regions as R
  lat = np.linspace(-90, 90, NLAT)
  lon = np.arange(NLON) * 0.25
  nfield = args.fields
  f = torch.randn((nfield, NLAT, NLON), device='cuda', dtype=torch.float32)
  t = torch.randn_like(f)
  c = torch.randn_like(f)
  slab = NLAT * NLON
  base = min(f.data_ptr(), t.data_ptr(), c.data_ptr())
  offs = [np.arange(nfield, dtype=np.int64) * slab + (v.data_ptr() - base) // 4
          for v in (f, t, c)]
  regs13 = [None, R.SliceRegion(lat_slice=slice(-20, 20)),
            R.ExtraTropicalRegion(),
            R.SliceRegion(lat_slice=slice(20, 90)),
            R.SliceRegion(lat_slice=slice(-90, -20)),
            R.SliceRegion(lat_slice=slice(35, 75),
                          lon_slice=[slice(347.5, None), slice(0, 42.5)]),
            R.SliceRegion(lat_slice=slice(25, 60), lon_slice=slice(240, 290)),
            R.SliceRegion(lat_slice=slice(25, 60), lon_slice=slice(145, 180)),
            R.SliceRegion(lat_slice=slice(25, 60), lon_slice=slice(102.5, 150)),
            R.SliceRegion(lat_slice=slice(-45, -12.5),
                          lon_slice=slice(120, 175)),
            R.SliceRegion(lat_slice=slice(-37.5, -22.5),
                          lon_slice=slice(15, 50)),
            R.SliceRegion(lat_slice=slice(-52.5, -20),
                          lon_slice=slice(287.5, 327.5)),
            R.SliceRegion(lat_slice=slice(-90, -60))]
  cases = [('global', [None]), ('3 lat-band regions', regs13[:3]),
           ('13 regions (lat/lon boxes)', regs13)]
  for label, regs in cases:
    (_, spec), = sp.build_weights(ctx, lat, lon, regs, 'lat_lon', NLON)
    out = torch.zeros((nfield, len(regs), _lib.DET_NSTAT), device='cuda',
                      dtype=torch.float64)
    for clim in (True, False):
      for path in ('tma', 'ldg'):
        os.environ['WB2_DET_PATH'] = path
        fn = lambda: ctx.det_metrics(
            base, base, base if clim else None, _lib.F32, offs[0], offs[1],
            offs[2] if clim else None, spec, False, out.data_ptr())
        ms = timeit(fn, args.steps)
        cells = nfield * slab
        report(f'det_metrics {label} clim={clim} path={path} nseg={spec.nseg}',
               ms, cells * (12 if clim else 8), cells, 'grid_cells')
  os.environ.pop('WB2_DET_PATH', None)


def bench_regrid(args):
  torch, _lib, ctx = setup()
  from weatherbench2_b200 import regridding
  src = regridding.Grid.from_degrees(np.arange(NLON) * 0.25,
                                     np.linspace(-90, 90, NLAT))
  tgt = regridding.Grid.from_degrees(np.arange(240) * 1.5,
                                     np.linspace(-90, 90, 121))
  rg = regridding.ConservativeRegridder(source=src, target=tgt)
  nfield = args.fields
  x = torch.randn((nfield, NLON, NLAT), device='cuda', dtype=torch.float32)
  out = torch.empty((nfield, 240, 121), device='cuda', dtype=torch.float32)
  fn = lambda: rg.regrid_device(ctx, x.data_ptr(), out.data_ptr(), nfield)
  ms = timeit(fn, args.steps)
  cells = nfield * NLAT * NLON
  report('regrid_conservative 0.25->1.5deg', ms,
         cells * 4 + nfield * 240 * 121 * 4, cells, 'source_cells')


def bench_spectrum(args):
  torch, _lib, ctx = setup()
  nfield = args.fields
  x = torch.randn((nfield, NLAT, NLON), device='cuda', dtype=torch.float32)
  nk = NLON // 2 + 1
  scale = np.cos(np.deg2rad(np.linspace(-90, 90, NLAT))) * 4.0e7
  out = torch.empty((nfield, NLAT, nk), device='cuda', dtype=torch.float32)
  fn = lambda: ctx.zonal_spectrum(x.data_ptr(), nfield, NLAT, NLON, scale,
                                  out.data_ptr())
  ms = timeit(fn, args.steps)
  cells = nfield * NLAT * NLON
  report('zonal_spectrum per-time output', ms, cells * 4 + nfield * NLAT * nk *
         4, cells, 'grid_cells')
  acc = torch.zeros((13, NLAT, nk), device='cuda', dtype=torch.float32)
  fn = lambda: ctx.zonal_spectrum(x.data_ptr(), nfield, NLAT, NLON, scale,
                                  acc.data_ptr(), True, 13)
  ms = timeit(fn, args.steps)
  report('zonal_spectrum time-summed output', ms, cells * 4, cells,
         'grid_cells')
  red = torch.zeros((13, nk), device='cuda', dtype=torch.float32)
  fn = lambda: ctx.zonal_spectrum_latsum(x.data_ptr(), nfield, NLAT, NLON,
                                         scale, red.data_ptr(), 13)
  ms = timeit(fn, args.steps)
  report('zonal_spectrum + latitude-weighted reduction (latsum)', ms,
         cells * 4, cells, 'grid_cells')


def bench_maps(args):
  """K6 / K6e: Spatial* maps; per-time maps (ngroup 1) and a 10-step time mean
  fused in (ngroup 10)."""
  torch, _lib, ctx = setup()
  slab = NLAT * NLON
  nfield = 390  # 10 times x 39 (variable, level) maps
  f = torch.randn((nfield, NLAT, NLON), device='cuda', dtype=torch.float32)
  t = torch.randn((nfield, NLAT, NLON), device='cuda', dtype=torch.float32)
  base = min(f.data_ptr(), t.data_ptr())
  for ngroup in (1, 10):
    nout = nfield // ngroup
    # group-major tables: output j averages times g of (variable, level) j
    idx = (np.arange(ngroup)[None, :] * nout + np.arange(nout)[:, None])
    off_f = (idx * slab + (f.data_ptr() - base) // 4).astype(np.int64).ravel()
    off_t = (idx * slab + (t.data_ptr() - base) // 4).astype(np.int64).ravel()
    out = torch.empty((nout, NLAT, NLON), device='cuda', dtype=torch.float32)
    fn = lambda: ctx.det_maps(base, base, _lib.F32, _lib.MAP_MSE, nout, ngroup,
                              off_f, off_t, NLAT, NLON, NLON, False,
                              out.data_ptr())
    ms = timeit(fn, args.steps)
    cells = nfield * slab
    report(f'det_maps MSE ngroup={ngroup}', ms, cells * 8 + nout * slab * 4,
           cells, 'grid_cells')
  del f, t
  for m in args.members:
    ntime, nmap = 5, 8
    x = torch.randn((m, ntime * nmap, NLAT, NLON), device='cuda',
                    dtype=torch.float32)
    t = torch.randn((ntime * nmap, NLAT, NLON), device='cuda',
                    dtype=torch.float32)
    base = min(x.data_ptr(), t.data_ptr())
    for ngroup, mask in ((1, _lib.ENS_CRPS), (ntime, _lib.ENS_CRPS),
                         (ntime, 63)):
      nout = ntime * nmap // ngroup
      idx = (np.arange(ngroup)[None, :] * nout + np.arange(nout)[:, None])
      off_x = (idx * slab + (x.data_ptr() - base) // 4).astype(
          np.int64).ravel()
      off_t = (idx * slab + (t.data_ptr() - base) // 4).astype(
          np.int64).ravel()
      nsel = bin(mask).count('1')
      out = torch.empty((nsel, nout, NLAT, NLON), device='cuda',
                        dtype=torch.float32)
      fn = lambda: ctx.ens_maps(base, base, _lib.F32, m, ntime * nmap * slab,
                                nout, ngroup, off_x, off_t, NLAT, NLON, NLON,
                                mask, False, out.data_ptr())
      ms = timeit(fn, args.steps)
      pts = ntime * nmap * slab
      report(f'ens_maps M={m} ngroup={ngroup} nsel={nsel}', ms,
             pts * (4 * m + 4) + nsel * nout * slab * 4, pts, 'grid_points')
    del x, t


def bench_threshold(args):
  """K7: ensemble Brier / debiased / ignorance / RPS for 4 Gaussian-quantile
  thresholds from one pass, and the Gaussian-forecast entry."""
  torch, _lib, ctx = setup()
  from weatherbench2_b200 import _spatial as sp
  lat = np.linspace(-90, 90, NLAT)
  lon = np.arange(NLON) * 0.25
  (_, spec), = sp.build_weights(ctx, lat, lon, [None], 'lat_lon', NLON)
  slab = NLAT * NLON
  nfield = args.fields
  z = np.array([-1.2816, -0.5244, 0.5244, 1.2816])
  for m in args.members:
    x = torch.randn((m, nfield, NLAT, NLON), device='cuda',
                    dtype=torch.float32)
    t = torch.randn((nfield, NLAT, NLON), device='cuda', dtype=torch.float32)
    cm = torch.zeros((NLAT, NLON), device='cuda', dtype=torch.float32)
    cs = torch.ones((NLAT, NLON), device='cuda', dtype=torch.float32)
    base = min(p.data_ptr() for p in (x, t, cm, cs))
    rel = lambda p: (p.data_ptr() - base) // 4
    off_x = np.arange(nfield, dtype=np.int64) * slab + rel(x)
    off_t = np.arange(nfield, dtype=np.int64) * slab + rel(t)
    off_a = np.zeros(nfield, dtype=np.int64) + rel(cm)
    off_b = np.zeros(nfield, dtype=np.int64) + rel(cs)
    out = torch.zeros((nfield, 4, 1, 8), device='cuda', dtype=torch.float64)
    fn = lambda: ctx.ens_threshold_metrics(
        base, base, m, nfield * slab, off_x, off_t, 4, base, off_a, base,
        off_b, z, spec, False, out.data_ptr())
    ms = timeit(fn, args.steps)
    pts = nfield * slab
    report(f'ens_threshold_metrics M={m} 4 thresholds', ms,
           pts * (4 * m + 4), pts, 'grid_points',
           {'note': 'climatology mean/std slab (8 MB) is L2-resident'})
    del x
  f = torch.randn((nfield * 10, NLAT, NLON), device='cuda',
                  dtype=torch.float32)
  s = torch.rand((nfield * 10, NLAT, NLON), device='cuda',
                 dtype=torch.float32) + 0.5
  t = torch.randn((nfield * 10, NLAT, NLON), device='cuda',
                  dtype=torch.float32)
  nf = nfield * 10
  base = min(p.data_ptr() for p in (f, s, t, cm, cs))
  rel = lambda p: (p.data_ptr() - base) // 4
  offs = [np.arange(nf, dtype=np.int64) * slab + rel(p) for p in (f, s, t)]
  off_a = np.zeros(nf, dtype=np.int64) + rel(cm)
  off_b = np.zeros(nf, dtype=np.int64) + rel(cs)
  out = torch.zeros((nf, 4, 1, 8), device='cuda', dtype=torch.float64)
  for nq in (0, 4):
    fn = lambda: ctx.gaussian_metrics(
        base, base, base, offs[0], offs[1], offs[2], nq, base if nq else None,
        off_a if nq else None, base if nq else None, off_b if nq else None,
        z[:nq] if nq else None, spec, False, out.data_ptr())
    ms = timeit(fn, args.steps)
    cells = nf * slab
    report(f'gaussian_metrics nthreshold={nq}', ms, cells * 12, cells,
           'grid_cells')


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--kernel', required=True,
                  choices=['ens', 'energy', 'det', 'regrid', 'spectrum',
                           'maps', 'threshold'])
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--fields', type=int, default=39)
  ap.add_argument('--members', type=int, nargs='+', default=[50])
  a = ap.parse_args()
  {'ens': bench_ens, 'energy': bench_energy, 'det': bench_det, 'regrid': bench_regrid,
   'spectrum': bench_spectrum, 'maps': bench_maps,
   'threshold': bench_threshold}[a.kernel](a)
