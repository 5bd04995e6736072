"""CPU-only, world_size = 2 over gloo: init-time chunks sharded across ranks,
one all-reduce of [sum, count] -- the N > 1 path of bench.py / evaluate_sharded
(weatherbench2/evaluation.py:693-744 replaced by torch.distributed).  The
per-chunk compute is injected (oracle-based) so no GPU is needed."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from weatherbench2_b200 import distributed as wd
from weatherbench2_b200 import xarray_lite as xl


def test_shard_indices_partition():
  for n in (0, 1, 7, 8, 730):
    for world in (1, 2, 3, 8):
      parts = [wd.shard_indices(n, r, world) for r in range(world)]
      np.testing.assert_array_equal(np.concatenate(parts), np.arange(n))
      sizes = [len(p) for p in parts]
      assert max(sizes) - min(sizes) <= 1


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _make_data(ninit=5, nlead=3, nlat=7, nlon=12, nan=False):
  rs = np.random.RandomState(0)
  lat = np.linspace(-90, 90, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  times = (np.datetime64('2020-01-01', 'ns') +
           np.arange(ninit + nlead) * np.timedelta64(1, 'D'))
  lead = np.arange(nlead) * np.timedelta64(1, 'D').astype('timedelta64[ns]')
  f = rs.normal(size=(ninit, nlead, nlat, nlon)).astype(np.float32)
  t = rs.normal(size=(ninit + nlead, nlat, nlon)).astype(np.float32)
  if nan:
    f[1, 0, 2, 3] = np.nan
  valid = times[:ninit, None] + lead[None, :]
  forecast = xl.Dataset(
      {'z': (('init_time', 'lead_time', 'latitude', 'longitude'), f)},
      {'init_time': times[:ninit], 'lead_time': lead, 'latitude': lat,
       'longitude': lon, 'valid_time': (('init_time', 'lead_time'), valid)})
  truth = xl.Dataset({'z': (('time', 'latitude', 'longitude'), t)},
                     {'time': times, 'latitude': lat, 'longitude': lon})
  return forecast, truth


def _oracle_loop(forecast, truth, eval_config, skipna, compute_chunk=True):
  """Stand-in for evaluation._metric_and_region_loop: per-(init, lead) MSE
  computed with the oracle on the materialised truth gather."""
  from oracle import wb2_oracle as orc
  del eval_config, compute_chunk
  f = forecast['z']
  t = truth['z']
  r, dims = orc.mse(f.values, f.dims, t.values, t.dims,
                    forecast['latitude'].values, forecast['longitude'].values,
                    skipna=False)
  return xl.Dataset({'z': (dims, r)},
                    {'init_time': forecast['init_time'].values,
                     'lead_time': forecast['lead_time'].values})


def _worker(rank, world, port, skipna, nan, outdir):
  import torch.distributed as dist
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  forecast, truth = _make_data(nan=nan)
  res = wd.evaluate_sharded(forecast, truth, None, skipna=skipna,
                            loop_fn=_oracle_loop)
  np.save(os.path.join(outdir, f'rank{rank}.npy'), res['z'].values)
  dist.destroy_process_group()


@pytest.mark.parametrize('skipna,nan', [(False, False), (False, True),
                                        (True, True)])
def test_two_rank_gloo_matches_single_process(tmp_path, skipna, nan):
  world = 2
  mp.spawn(_worker, args=(world, _free_port(), skipna, nan, str(tmp_path)),
           nprocs=world, join=True)
  forecast, truth = _make_data(nan=nan)
  from weatherbench2_b200 import evaluation
  full = _oracle_loop(forecast,
                      evaluation.select_truth_at_valid_time(truth, forecast),
                      None, skipna)['z']
  want = (np.nanmean(full.values, axis=0) if skipna
          else full.values.mean(axis=0))
  for r in range(world):
    got = np.load(tmp_path / f'rank{r}.npy')
    np.testing.assert_allclose(got, want, rtol=1e-12, equal_nan=True)
  if nan and not skipna:
    assert np.isnan(want).any()


def test_single_process_without_process_group():
  forecast, truth = _make_data()
  res = wd.evaluate_sharded(forecast, truth, None, loop_fn=_oracle_loop,
                            chunk_size=2)
  from weatherbench2_b200 import evaluation
  full = _oracle_loop(forecast,
                      evaluation.select_truth_at_valid_time(truth, forecast),
                      None, False)['z']
  np.testing.assert_allclose(res['z'].values, full.values.mean(axis=0),
                             rtol=1e-12)
  assert res['z'].dims == ('lead_time',)


def test_by_valid_chunks_align_truth_by_label_not_position():
  """chunk_dim='time': the truth record is longer than the forecast's and
  starts earlier; chunks must pick the truth of the SAME time labels
  (the reference aligns by label; ADVICE r1: positional slicing paired
  forecast chunks with the wrong truth)."""
  rs = np.random.RandomState(3)
  nlat, nlon = 5, 8
  lat = np.linspace(-90, 90, nlat)
  lon = np.linspace(0, 360, nlon, endpoint=False)
  ttimes = (np.datetime64('2020-01-01', 'ns') +
            np.arange(12) * np.timedelta64(1, 'D'))
  ftimes = ttimes[4:9]  # starts 4 days into the truth record
  f = rs.normal(size=(5, nlat, nlon)).astype(np.float32)
  t = rs.normal(size=(12, nlat, nlon)).astype(np.float32)
  dims = ('time', 'latitude', 'longitude')
  forecast = xl.Dataset({'z': (dims, f)}, {'time': ftimes, 'latitude': lat,
                                           'longitude': lon})
  truth = xl.Dataset({'z': (dims, t)}, {'time': ttimes, 'latitude': lat,
                                        'longitude': lon})

  def loop(fc, tr, eval_config, skipna, compute_chunk=True):
    from oracle import wb2_oracle as orc
    del eval_config, compute_chunk
    np.testing.assert_array_equal(fc['time'].values, tr['time'].values)
    r, d = orc.mse(fc['z'].values, fc['z'].dims, tr['z'].values, tr['z'].dims,
                   lat, lon, skipna=skipna)
    return xl.Dataset({'z': (d, r)}, {'time': fc['time'].values})

  res = wd.evaluate_sharded(forecast, truth, None, chunk_dim='time',
                            chunk_size=2, loop_fn=loop)
  from oracle import wb2_oracle as orc
  want, _ = orc.mse(f, dims, t[4:9], dims, lat, lon)
  np.testing.assert_allclose(res['z'].values, want.mean(), rtol=1e-12)
  # a forecast time the truth does not have is an error, not a silent drop
  bad = xl.Dataset({'z': (dims, f)},
                   {'time': ftimes + np.timedelta64(100, 'D'),
                    'latitude': lat, 'longitude': lon})
  with pytest.raises(KeyError):
    wd.evaluate_sharded(bad, truth, None, chunk_dim='time', loop_fn=loop)


def test_prefetching_feeder_gives_the_same_result():
  forecast, truth = _make_data(ninit=7)
  a = wd.evaluate_sharded(forecast, truth, None, loop_fn=_oracle_loop,
                          chunk_size=2)
  b = wd.evaluate_sharded(forecast, truth, None, loop_fn=_oracle_loop,
                          chunk_size=2, prefetch=2, num_threads=3)
  np.testing.assert_array_equal(a['z'].values, b['z'].values)


# ---- the product loop itself (operators -> offset tables -> C-ABI arguments)
# under gloo, with the NumPy stand-in context of tests/fake_ctx.py ----------------
def _eval_config():
  from weatherbench2_b200 import config, metrics, regions as R
  return config.Eval(
      metrics={'rmse': metrics.RMSESqrtBeforeTimeAvg(), 'bias': metrics.Bias()},
      regions={'global': R.SliceRegion(),
               'tropics': R.SliceRegion(lat_slice=slice(-20, 20)),
               'extra-tropics': R.ExtraTropicalRegion()})


def _product_worker(rank, world, port, prefetch, outdir):
  import torch.distributed as dist
  import fake_ctx
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  forecast, truth = _make_data(ninit=5)
  with fake_ctx.installed() as fake:
    res = wd.evaluate_sharded(forecast, truth, _eval_config(), prefetch=prefetch)
  np.save(os.path.join(outdir, f'rank{rank}.npy'), res['z'].values)
  np.save(os.path.join(outdir, f'calls{rank}.npy'), len(fake.calls))
  dist.destroy_process_group()


@pytest.mark.parametrize('prefetch', [0, 2])
def test_two_rank_gloo_runs_the_product_loop(tmp_path, prefetch):
  from oracle import wb2_oracle as orc
  import fake_ctx
  world = 2
  mp.spawn(_product_worker, args=(world, _free_port(), prefetch, str(tmp_path)),
           nprocs=world, join=True)
  got = [np.load(tmp_path / f'rank{r}.npy') for r in range(world)]
  np.testing.assert_array_equal(got[0], got[1])
  # 5 init-time chunks over 2 ranks: 3 + 2, one K1 pass per chunk for both
  # metrics and all three regions
  assert [int(np.load(tmp_path / f'calls{r}.npy')) for r in range(world)] == [
      3, 2]
  forecast, truth = _make_data(ninit=5)
  f = forecast['z'].values
  vt = forecast['valid_time'].values
  pos = {t: i for i, t in enumerate(truth['time'].values)}
  tg = np.stack([np.stack([truth['z'].values[pos[v]] for v in row])
                 for row in vt])
  dims = ('time', 'lead_time', 'latitude', 'longitude')
  lat, lon = forecast['latitude'].values, forecast['longitude'].values
  regs = [None, orc.SliceRegion(lat_slice=slice(-20, 20)),
          orc.ExtraTropicalRegion()]
  for ri, reg in enumerate(regs):
    want, wdims = orc.rmse_sqrt_before_time_avg(f, dims, tg, dims, lat, lon,
                                                region=reg)
    np.testing.assert_allclose(got[0][0, ri], want.mean(axis=wdims.index('time')),
                               rtol=2e-6)
    want, wdims = orc.bias(f, dims, tg, dims, lat, lon, region=reg)
    np.testing.assert_allclose(got[0][1, ri], want.mean(axis=wdims.index('time')),
                               rtol=2e-6, atol=1e-9)
  # single process, same loop: identical numbers
  with fake_ctx.installed():
    single = wd.evaluate_sharded(forecast, truth, _eval_config())
  np.testing.assert_allclose(single['z'].values, got[0], rtol=1e-12)


# ---- evaluate_with_beam drop-in == evaluate_in_memory (the reference's own
# consistency test, weatherbench2/evaluation_test.py:30-128) -----------------------
def _beam_worker(rank, world, port, by_init, outdir):
  import torch.distributed as dist
  import fake_ctx
  from weatherbench2_b200 import evaluation
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  import test_evaluation_cpu as ev
  dc, eval_configs = ev.consistency_setup(os.path.join(outdir, 'beam'),
                                          by_init)
  chunk_dim = 'init_time' if by_init else 'time'
  with fake_ctx.installed():
    out = evaluation.evaluate_with_beam(dc, eval_configs,
                                        input_chunks={chunk_dim: 2},
                                        runner='DirectRunner', num_threads=2)
  for name, ds in out.items():
    np.save(os.path.join(outdir, f'{name}.rank{rank}.npy'),
            ds['geopotential'].values)
  dist.destroy_process_group()


@pytest.mark.parametrize('by_init', [True, False])
def test_in_memory_and_distributed_consistency(tmp_path, by_init):
  import fake_ctx
  from weatherbench2_b200 import evaluation
  world = 2
  mp.spawn(_beam_worker, args=(world, _free_port(), by_init, str(tmp_path)),
           nprocs=world, join=True)
  import test_evaluation_cpu as ev
  dc, eval_configs = ev.consistency_setup(tmp_path / 'mem', by_init)
  with fake_ctx.installed():
    mem = evaluation.evaluate_in_memory(dc, eval_configs)
  for name in eval_configs:
    got = [np.load(tmp_path / f'{name}.rank{r}.npy') for r in range(world)]
    np.testing.assert_array_equal(got[0], got[1])
    want = mem[name]['geopotential'].values
    assert got[0].shape == want.shape, name
    np.testing.assert_allclose(got[0], want, rtol=1e-12, atol=1e-15,
                               err_msg=name)
    # rank 0 wrote the file, as evaluate_in_memory does
    assert (tmp_path / 'beam' / f'{name}.npz').exists()


def _idle_rank_worker(rank, world, port, temporal_mean, outdir):
  import torch.distributed as dist
  import fake_ctx
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  forecast, truth = _make_data(ninit=2)
  with fake_ctx.installed() as fake:
    res = wd.evaluate_sharded(forecast, truth, _eval_config(),
                              temporal_mean=temporal_mean)
  np.save(os.path.join(outdir, f'rank{rank}.npy'), res['z'].values)
  np.save(os.path.join(outdir, f'calls{rank}.npy'), len(fake.calls))
  dist.destroy_process_group()


@pytest.mark.parametrize('temporal_mean', [True, False])
def test_more_ranks_than_chunks(tmp_path, temporal_mean):
  """Two init times over three ranks: the rank without a chunk still takes
  part in the collective (zeros for the mean, nothing for the gather) and
  every rank ends with the single-process result."""
  import fake_ctx
  world = 3
  mp.spawn(_idle_rank_worker,
           args=(world, _free_port(), temporal_mean, str(tmp_path)),
           nprocs=world, join=True)
  forecast, truth = _make_data(ninit=2)
  with fake_ctx.installed():
    want = wd.evaluate_sharded(forecast, truth, _eval_config(),
                               temporal_mean=temporal_mean)['z'].values
  for r in range(world):
    np.testing.assert_allclose(np.load(tmp_path / f'rank{r}.npy'), want,
                               rtol=1e-12)
  calls = [int(np.load(tmp_path / f'calls{r}.npy')) for r in range(world)]
  # chunks 0 and 1 go to ranks 0 and 1; for the mean the idle rank evaluates
  # one probe chunk only to learn the payload shape
  assert calls[:2] == [1, 1] and calls[2] == (1 if temporal_mean else 0)
