#!/usr/bin/env python
"""Headline benchmark: grid-cells/s of the fused weighted RMSE + Bias + ACC
kernel on 721 x 1440 x 13-level fields (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W          # this framework
  python bench.py --impl reference --gpus N ...          # CPU reference arm

A "step" is one pass of the hot path over one evaluation chunk: one init time
x 10 lead times x 6 variables x 13 levels = 780 fields of 721 x 1440 float32
for forecast, truth and climatology (9.72 GB of HBM traffic, far larger than
the 126 MB L2, so no flush is needed between steps).  Under torchrun every rank
owns a different chunk (weak scaling; chunks = init times are independent,
weatherbench2/evaluation.py:583-599) and the time-summed statistics are
all-reduced once at the end (the NCCL equivalent of xbeam.Mean,
evaluation.py:740-744).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput,
`e2e` = the same chunk through the public operator API with pinned HOST inputs
(H2D inside the timed region), `roofline` = achieved algorithmic HBM GB/s of
the dominant kernel vs MEASURED_PEAKS.json, `cpu_baseline` = the oracle port
timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

NLAT, NLON, NLEV, NVAR, NLEAD = 721, 1440, 13, 6, 10
VARIABLES = ['geopotential', 'temperature', 'u_component_of_wind',
             'v_component_of_wind', 'specific_humidity', 'vertical_velocity']
LEVELS = [50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000]
BYTES_PER_CELL = 12  # f + t + c, float32


def _ncu_traffic():
  """dram bytes per launch of the dominant kernel from the committed ncu
  capture (profiles/r1_k1_traffic.json); None if absent or another path."""
  if os.environ.get('WB2_DET_PATH') == 'ldg':
    return None
  try:
    with open(os.path.join(ROOT, 'profiles', 'r1_k1_traffic.json')) as fh:
      return float(json.load(fh)['dram_bytes_per_launch'])
  except Exception:  # pylint: disable=broad-except
    return None


def _peak_gbs():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    with open(path) as fh:
      return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  except Exception:  # pylint: disable=broad-except
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
  """Samples SM clocks / throttle reasons with nvidia-smi while timing."""
  QUERY = ('clocks.sm,clocks.max.sm,power.draw,'
           'clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, index=0):
    self.index = index
    self.proc = None
    self.lines = []

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', f'--query-gpu={self.QUERY}',
           '--format=csv,noheader,nounits', '-lms', '100', '-i',
           str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
          text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:  # pylint: disable=broad-except
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:  # pylint: disable=broad-except
      self.proc.kill()
    sm, smax, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
             'sw_power_cap']
    for line in self.lines:
      parts = [p.strip() for p in line.split(',')]
      if len(parts) < 7:
        continue
      try:
        sm.append(float(parts[0]))
        smax.append(float(parts[1]))
      except ValueError:
        continue
      for n, v in zip(names, parts[3:7]):
        if v.lower().startswith('active'):
          reasons.add(n)
    return {'sm_mhz': float(np.median(sm)) if sm else None,
            'sm_max_mhz': float(max(smax)) if smax else None,
            'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference needs xarray/jax, not installed)
# ------------------------------------------------------------------------------
def _oracle_inputs(seed):
  rs = np.random.RandomState(seed)
  shape = (NLEV, NLAT, NLON)
  return tuple(rs.standard_normal(shape).astype(np.float32) for _ in range(3))


def _oracle_compute(f, t, c):
  """MSE + Bias + ACC with the oracle (op-for-op restatement of
  weatherbench2/metrics.py) on 1 variable x 13 levels x 721 x 1440; returns
  the seconds of the metric computation only."""
  from oracle import wb2_oracle as orc
  lat = np.linspace(-90, 90, NLAT)
  lon = np.arange(NLON) * 0.25
  dims = ('level', 'latitude', 'longitude')
  t0 = time.perf_counter()
  orc.rmse_sqrt_before_time_avg(f, dims, t, dims, lat, lon)
  orc.bias(f, dims, t, dims, lat, lon)
  orc.acc(f, dims, t, dims, c, dims, lat, lon)
  return time.perf_counter() - t0


def _oracle_sample_step(seed):
  f, t, c = _oracle_inputs(seed)
  return _oracle_compute(f, t, c), f.size


def _worker(args):
  """One host process of the reference arm: inputs are generated once (not
  timed), then `reps` steps of the metric computation."""
  seed, reps = args
  os.environ.setdefault('OMP_NUM_THREADS', '1')
  f, t, c = _oracle_inputs(seed)
  total_t, total_cells = 0.0, 0
  for _ in range(reps):
    total_t += _oracle_compute(f, t, c)
    total_cells += f.size
  return total_t, total_cells


def cpu_baseline_single():
  """Oracle on one core, bounded sample (~10-20 s)."""
  _oracle_sample_step(0)  # warm-up (page faults, imports)
  reps, tt, cells = 3, 0.0, 0
  for r in range(reps):
    dt, n = _oracle_sample_step(r + 1)
    tt += dt
    cells += n
  return {'value': cells / tt, 'unit': 'grid-cells/s', 'cores': 1,
          'kind': 'port',
          'sample': f'{reps} x (1 variable x {NLEV} levels x {NLAT}x{NLON} = '
                    f'{NLEV * NLAT * NLON} cells), RMSE+Bias+ACC via '
                    'oracle/wb2_oracle.py (NumPy restatement of the xarray '
                    'path), time.perf_counter around the compute only'}


def run_reference(args):
  """--impl reference: the oracle port on ALL host cores (the reference itself
  needs xarray, which is not installed on this box: `kind` = "port")."""
  rank = int(os.environ.get('RANK', 0))
  if rank != 0:
    return
  import multiprocessing as mp
  cores = os.cpu_count() or 1
  workers = max(1, min(cores, 256))  # all host cores
  ctxm = mp.get_context('fork')
  with ctxm.Pool(workers) as pool:
    for _ in range(max(1, min(args.warmup, 1))):
      pool.map(_worker, [(i, 1) for i in range(workers)])
    res = pool.map(_worker, [(100 + i, args.steps) for i in range(workers)])
  # all processes compute concurrently; the job takes as long as the slowest
  # one spends in the metric code (input generation is not part of the path)
  wall = max(r[0] for r in res)
  cells = sum(r[1] for r in res)
  value = cells / wall
  sample = (f'{workers} processes x {args.steps} steps x (1 variable x {NLEV} '
            f'levels x {NLAT}x{NLON}), RMSE+Bias+ACC via oracle/wb2_oracle.py')
  line = {
      'impl': 'reference', 'metric': 'grid-cells/s', 'value': value,
      'unit': 'grid-cells/s', 'n_gpus': args.gpus, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': 1e3 * wall / max(1, args.steps),
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'f32', 'data': 'synthetic',
      'config': _config(args.gpus),
      'cpu_baseline': {'value': value, 'unit': 'grid-cells/s',
                       'cores': workers, 'kind': 'port', 'sample': sample},
      'e2e': {'value': value, 'unit': 'grid-cells/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


def _config(n_gpus):
  return {'workload': 'configs[1]: RMSE+Bias+ACC, 6 vars x 13 levels x '
                      '721x1440, chunk = 1 init x 10 lead (780 fields, '
                      '9.72 GB f32 per step per GPU)',
          'cells_per_step_per_gpu': NLEAD * NVAR * NLEV * NLAT * NLON,
          'regions': 1, 'skipna': False,
          'l2_policy': 'inputs (9.72 GB) >> L2 (126 MB); no flush needed',
          'parallelism': f'chunks sharded over {n_gpus} GPU(s), one NCCL '
                         'all-reduce of the time-sum at the end'}


# ------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------
def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--e2e-steps', type=int, default=2)
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-cpu', action='store_true')
  ap.add_argument('--traffic', type=float, default=None,
                  help='dram bytes per launch from an ncu --set full capture')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
    return
  args.warmup = max(args.warmup, 3)

  import torch
  import torch.distributed as dist
  from weatherbench2_b200 import _lib, _spatial as sp

  rank = int(os.environ.get('RANK', 0))
  world = int(os.environ.get('WORLD_SIZE', 1))
  local = int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(local)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  dev = torch.device('cuda', local)
  ctx = _lib.Context(local)
  # one explicit (non-legacy) stream shared by torch and the library, so that
  # torch.cuda.Event timing sees the kernels
  stream = torch.cuda.Stream(device=dev)
  torch.cuda.set_stream(stream)
  ctx.set_stream(stream.cuda_stream)

  # ---- synthetic chunk, resident in HBM -------------------------------------
  nfield = NLEAD * NVAR * NLEV
  slab = NLAT * NLON
  gen = torch.Generator(device=dev)
  gen.manual_seed(802701 + rank)
  f = torch.randn((nfield, NLAT, NLON), device=dev, dtype=torch.float32,
                  generator=gen)
  t = torch.randn((nfield, NLAT, NLON), device=dev, dtype=torch.float32,
                  generator=gen)
  c = torch.randn((nfield, NLAT, NLON), device=dev, dtype=torch.float32,
                  generator=gen)
  lat = np.linspace(-90, 90, NLAT)
  lon = np.arange(NLON) * 0.25
  (_, spec), = sp.build_weights(ctx, lat, lon, [None], 'lat_lon', NLON)
  base = min(f.data_ptr(), t.data_ptr(), c.data_ptr())
  offs = [np.arange(nfield, dtype=np.int64) * slab + (x.data_ptr() - base) // 4
          for x in (f, t, c)]
  nstat = _lib.DET_NSTAT
  total_steps = args.warmup + args.steps
  out = torch.zeros((total_steps, nfield, nstat), device=dev,
                    dtype=torch.float64)

  def step(i):
    ctx.det_metrics(base, base, base, _lib.F32, offs[0], offs[1], offs[2],
                    spec, False, out[i].data_ptr())

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for i in range(args.warmup):
    step(i)
  _ = out[:args.warmup].sum(dim=0)  # warm the reduction used after the loop
  if world > 1:
    dist.all_reduce(_)
  barrier()
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  launches0 = ctx.launch_count
  ev0 = torch.cuda.Event(enable_timing=True)
  ev1 = torch.cuda.Event(enable_timing=True)
  ev_k = torch.cuda.Event(enable_timing=True)
  barrier()
  ev0.record()
  for i in range(args.steps):
    step(args.warmup + i)
  ev_k.record()
  # time mean over the steps of this rank, then ONE all-reduce (sum, count)
  tsum = out[args.warmup:].sum(dim=0)
  if world > 1:
    dist.all_reduce(tsum)
  ev1.record()
  barrier()
  launches = ctx.launch_count - launches0
  ms_total = ev0.elapsed_time(ev1)
  ms_kernels = ev0.elapsed_time(ev_k)
  tmax = torch.tensor([ms_total, ms_kernels], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
  ms_total, ms_kernels = [float(x) for x in tmax.tolist()]
  clocks = sampler.stop() if rank == 0 else None

  cells_per_step = nfield * slab
  value = world * cells_per_step * args.steps / (ms_total * 1e-3)
  peak, peak_src = _peak_gbs()
  kernel_ms = ms_kernels / args.steps
  achieved = cells_per_step * BYTES_PER_CELL / (kernel_ms * 1e-3) / 1e9

  # sanity of the numbers we just produced (cheap, on rank 0): statistic 6 is
  # the weight sum = nlat * nlon
  wsum = float(out[args.warmup, 0, 6].item())
  assert abs(wsum - NLAT * NLON) < 1e-3 * NLAT * NLON, wsum

  line = {
      'metric': 'grid-cells/s', 'value': value, 'unit': 'grid-cells/s',
      'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic', 'config': _config(world), 'clocks': clocks,
      'gpu_launches': int(launches),
      'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak,
                   'unit': 'GB/s', 'frac': achieved / peak,
                   'traffic': (args.traffic if args.traffic is not None
                               else _ncu_traffic()),
                   'peak_source': peak_src,
                   'kernel': ('det_metrics_kernel<float,4,CLIM> (LDG path)'
                              if os.environ.get('WB2_DET_PATH') == 'ldg' else
                              'det_tma_kernel<CLIM,!SKIPNA> (TMA ring)') +
                             ' + finalize',
                   'kernel_ms': kernel_ms,
                   'algorithmic_bytes_per_launch':
                       cells_per_step * BYTES_PER_CELL},
  }

  # ---- end to end through the public operator API with HOST inputs ----------
  if not args.no_e2e:
    line['e2e'] = run_e2e(ctx, args, rank, world, dev, f, t, c, lat, lon)
  del f, t, c
  if rank == 0 and not args.no_cpu:
    line['cpu_baseline'] = cpu_baseline_single()
  if rank == 0:
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def run_e2e(ctx, args, rank, world, dev, f, t, c, lat, lon):
  """evaluation._metric_and_region_loop on pinned-host datasets: every step
  copies the chunk host->device inside the timed region (wb2_det_metrics_host)
  and reads the result back."""
  import pandas as pd
  import torch
  import torch.distributed as dist
  from weatherbench2_b200 import config, evaluation, metrics, _lib
  from weatherbench2_b200 import xarray_lite as xl

  nfield = NLEAD * NVAR * NLEV
  shape5 = (NLEAD, NVAR, NLEV, NLAT, NLON)

  def to_pinned(x):
    h = ctx.pinned_empty(shape5, np.float32)
    _lib.check(ctx.lib.wb2_memcpy_d2h(ctx.handle, h.ctypes.data, x.data_ptr(),
                                      h.nbytes))
    return h

  hf, ht, hc = to_pinned(f), to_pinned(t), to_pinned(c)
  init = np.array(['2020-01-01T00'], dtype='datetime64[ns]')
  lead = (np.arange(NLEAD) * 24 * 3600 * 10**9).astype('timedelta64[ns]')
  valid = init[:, None] + lead[None, :]
  levels = np.array(LEVELS)
  fcoords = {'init_time': init, 'lead_time': lead, 'level': levels,
             'latitude': lat, 'longitude': lon,
             'valid_time': (('init_time', 'lead_time'), valid)}
  fdims = ('init_time', 'lead_time', 'level', 'latitude', 'longitude')
  tdims = ('time', 'level', 'latitude', 'longitude')
  cdims = ('dayofyear', 'level', 'latitude', 'longitude')
  doy = pd.DatetimeIndex(valid.ravel()).dayofyear.values
  forecast = xl.Dataset({v: (fdims, hf[None, :, i]) for i, v in
                         enumerate(VARIABLES)}, fcoords)
  truth_src = xl.Dataset({v: (tdims, ht[:, i]) for i, v in
                          enumerate(VARIABLES)},
                         {'time': valid.ravel(), 'level': levels,
                          'latitude': lat, 'longitude': lon})
  clim = xl.Dataset({v: (cdims, hc[:, i]) for i, v in enumerate(VARIABLES)},
                    {'dayofyear': doy, 'level': levels, 'latitude': lat,
                     'longitude': lon})
  truth = evaluation.select_truth_at_valid_time(truth_src, forecast)
  eval_config = config.Eval(
      metrics={'rmse': metrics.RMSESqrtBeforeTimeAvg(), 'bias': metrics.Bias(),
               'acc': metrics.ACC(climatology=clim)}, temporal_mean=False)

  def one():
    return evaluation._metric_and_region_loop(  # pylint: disable=protected-access
        forecast, truth, eval_config, skipna=False, compute_chunk=True)

  one()  # warm-up (staging buffers, page tables)
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(args.e2e_steps):
    res = one()
  torch.cuda.synchronize()
  dt = time.perf_counter() - t0
  tm = torch.tensor([dt], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
  dt = float(tm.item())
  d2h = sum(res[v].values.nbytes for v in res.keys())
  cells = nfield * NLAT * NLON
  for h in (hf, ht, hc):
    ctx.host_free(h.ctypes.data)
  return {'value': world * cells * args.e2e_steps / dt, 'unit': 'grid-cells/s',
          'h2d_bytes_per_step': int(3 * cells * 4),
          'd2h_bytes_per_step': int(d2h), 'steps': args.e2e_steps,
          'ms_per_step': 1e3 * dt / args.e2e_steps,
          'api': 'evaluation._metric_and_region_loop(RMSE+Bias+ACC) on '
                 'pinned-host datasets -> wb2_det_metrics_host'}


if __name__ == '__main__':
  main()
