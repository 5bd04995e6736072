#!/usr/bin/env python
"""Zonal-energy-spectrum sweep scaling (BASELINE.json configs[4]): time-mean
spectra of 5 variables x 37 levels on the 721 x 1440 grid.  Every rank owns
`--steps` chunks of `--times-per-chunk` time steps (weak scaling), accumulates the time SUM of its spectra on
the device inside K4 (a CTA owns its output rows and walks time in order), and
one NCCL all-reduce of the 185 x 721 x 721 float32 sums (385 MB) finishes the
mean -- the NCCL form of `xbeam.Mean(['time'])`
(scripts/compute_zonal_energy_spectrum.py:234).

  python benchmarks/bench_spectrum_sweep.py --steps 4
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
      benchmarks/bench_spectrum_sweep.py --gpus 8 --steps 4

Same timing rules as bench.py: >= 3 warm-ups, inputs (3.07 GB per chunk)
>> L2, CUDA events on the shared stream, barrier + synchronize on both sides,
max over ranks.  One JSON line on rank 0; `value` is grid cells per second over
all ranks, collective included.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NLAT, NLON, NLEV, NVAR = 721, 1440, 37, 5


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=4)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--times-per-chunk', type=int, default=4)
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3)

  import torch
  import torch.distributed as dist
  from weatherbench2_b200 import _lib, derived_variables

  rank = int(os.environ.get('RANK', 0))
  world = int(os.environ.get('WORLD_SIZE', 1))
  local = int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
  stream = torch.cuda.Stream(device=dev)
  torch.cuda.set_stream(stream)
  ctx = _lib.Context(local)
  ctx.set_stream(stream.cuda_stream)

  nmap = NVAR * NLEV
  gen = torch.Generator(device=dev)
  gen.manual_seed(802701 + rank)
  # one chunk of T time steps resident (T x 0.77 GB); every step re-reads it
  T = args.times_per_chunk
  x = torch.randn((T, nmap, NLAT, NLON), device=dev, dtype=torch.float32,
                  generator=gen)
  lat = np.linspace(-90, 90, NLAT)
  scale = derived_variables.ZonalEnergySpectrum('x')._circumference(lat)  # pylint: disable=protected-access
  nk = NLON // 2 + 1
  acc = torch.zeros((nmap, NLAT, nk), device=dev, dtype=torch.float32)

  def step(i):
    del i
    ctx.zonal_spectrum(x.data_ptr(), T * nmap, NLAT, NLON, scale,
                       acc.data_ptr(), True, nmap)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for i in range(args.warmup):
    step(i)
  if world > 1:
    dist.all_reduce(acc)
  acc.zero_()
  barrier()
  e0 = torch.cuda.Event(enable_timing=True)
  e1 = torch.cuda.Event(enable_timing=True)
  launches0 = ctx.launch_count
  barrier()
  e0.record()
  for i in range(args.steps):
    step(i)
  if world > 1:
    dist.all_reduce(acc)
  e1.record()
  barrier()
  ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  ms = float(ms.item())
  cells = T * nmap * NLAT * NLON
  value = world * cells * args.steps / (ms * 1e-3)
  try:
    peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))[
        'hbm_gbs'])
  except Exception:  # pylint: disable=broad-except
    peak = 6650.0
  gbs = cells * 4 * args.steps / (ms * 1e-3) / 1e9  # per GPU, collective incl.
  mean = acc / float(args.steps * world * T)
  total = float(mean.double().sum().item())
  assert np.isfinite(total) and total > 0, total
  if rank == 0:
    print(json.dumps({
        'metric': 'grid-cells/s', 'value': value, 'unit': 'grid-cells/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'configs[4]: zonal energy spectrum, time mean, '
                               f'{NVAR} vars x {NLEV} levels x {NLAT}x{NLON} '
                               f'per time step, chunks of {T} time steps ({cells * 4 / 1e9:.2f} GB)',
                   'chunks_per_rank': args.steps,
                   'collective': f'one all_reduce of {acc.numel() * 4 / 1e6:.0f}'
                                 ' MB at the end (inside the timed region)'},
        'gpu_launches': int(ctx.launch_count - launches0),
        'roofline': {'bound': 'hbm (instruction-bound kernel, DESIGN.md)',
                     'achieved': gbs, 'peak': peak, 'unit': 'GB/s',
                     'frac': gbs / peak},
        'mean_power': total / (nmap * NLAT)}))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
