#!/usr/bin/env python
"""CRPS sweep scaling (BASELINE.json configs[2] / the north star's 8-GPU
target): CRPS skill + spread + ensemble-mean MSE + variance from ONE pass of K2
over a 50-member ensemble, 3 variables x 13 levels x 721 x 1440 per
(init, lead) chunk (8.26 GB), chunks sharded over the ranks (weak scaling: every
rank owns `--steps` chunks), one NCCL all-reduce of the time sums at the end.

  python benchmarks/bench_crps_sweep.py --steps 10
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
      benchmarks/bench_crps_sweep.py --gpus 8 --steps 10

Same timing rules as bench.py: >= 3 warm-ups, inputs >> L2, CUDA events on the
shared stream, barrier + synchronize on both sides, max over ranks.  One JSON
line on rank 0; `value` is grid points (each carrying M members) per second
over all ranks.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NLAT, NLON, NLEV, NVAR = 721, 1440, 13, 3


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--members', type=int, default=50)
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3)

  import torch
  import torch.distributed as dist
  from weatherbench2_b200 import _lib, _spatial as sp

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

  m = args.members
  nfield = NVAR * NLEV
  slab = NLAT * NLON
  gen = torch.Generator(device=dev)
  gen.manual_seed(802701 + rank)
  x = torch.randn((m, nfield, NLAT, NLON), device=dev, dtype=torch.float32,
                  generator=gen)
  t = torch.randn((nfield, NLAT, NLON), device=dev, dtype=torch.float32,
                  generator=gen)
  lat = np.linspace(-90, 90, NLAT)
  lon = np.arange(NLON) * 0.25
  (_, spec), = sp.build_weights(ctx, lat, lon, [None], 'lat_lon', NLON)
  base = min(x.data_ptr(), t.data_ptr())
  off_x = np.arange(nfield, dtype=np.int64) * slab + (x.data_ptr() - base) // 4
  off_t = np.arange(nfield, dtype=np.int64) * slab + (t.data_ptr() - base) // 4
  total = args.warmup + args.steps
  out = torch.zeros((total, nfield, _lib.ENS_NSTAT), device=dev,
                    dtype=torch.float64)

  def step(i):
    ctx.ens_metrics(base, base, _lib.F32, m, nfield * slab, off_x, off_t, spec,
                    False, out[i].data_ptr())

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for i in range(args.warmup):
    step(i)
  w = out[:args.warmup].sum(dim=0)
  if world > 1:
    dist.all_reduce(w)
  barrier()
  e0 = torch.cuda.Event(enable_timing=True)
  e1 = torch.cuda.Event(enable_timing=True)
  launches0 = ctx.launch_count
  barrier()
  e0.record()
  for i in range(args.steps):
    step(args.warmup + i)
  tsum = out[args.warmup:].sum(dim=0)
  if world > 1:
    dist.all_reduce(tsum)
  e1.record()
  barrier()
  ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  ms = float(ms.item())
  pts = nfield * slab
  value = world * pts * args.steps / (ms * 1e-3)
  try:
    peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))[
        'hbm_gbs'])
  except Exception:  # pylint: disable=broad-except
    peak = 6650.0
  gbs = pts * (4 * m + 4) * args.steps / (ms * 1e-3) / 1e9  # per GPU
  # CRPS of the time-summed statistics (sanity: finite, ~ 2/sqrt(pi)/2 scale)
  st = tsum.cpu().numpy() / (args.steps * world)
  crps = float(np.mean(st[:, 0] / st[:, 5] - 0.5 * st[:, 1] / st[:, 6]))
  assert np.isfinite(crps), crps
  if rank == 0:
    print(json.dumps({
        'metric': 'grid-points/s (M members each)', 'value': value,
        'unit': 'grid-points/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': f'configs[2]: CRPS skill+spread+ens-mean MSE+'
                               f'variance, M={m}, {NVAR} vars x {NLEV} levels '
                               f'x {NLAT}x{NLON} per chunk '
                               f'({pts * (4 * m + 4) / 1e9:.2f} GB)',
                   'chunks_per_rank': args.steps},
        'gpu_launches': int(ctx.launch_count - launches0),
        'roofline': {'bound': 'hbm (ALU-pipe ridge, DESIGN.md)',
                     'achieved': gbs, 'peak': peak, 'unit': 'GB/s',
                     'frac': gbs / peak},
        'mean_crps': crps}))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
