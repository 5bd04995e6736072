"""ChunkFeeder (the DatasetToChunks replacement, evaluation.py:693-705) on the
CPU: chunk contents and keys, prefetch depth, memory-mapped sources, error
propagation."""
import threading
import time

import numpy as np
import pytest

from weatherbench2_b200 import feeder as fd
from weatherbench2_b200 import xarray_lite as xl


def _dataset(tmp_path=None, ntime=11):
  rs = np.random.RandomState(0)
  shape = (ntime, 3, 5, 8)
  a = rs.standard_normal(shape).astype(np.float32)
  if tmp_path is not None:  # a memory-mapped source, like a chunked store
    path = tmp_path / 'a.dat'
    mm = np.memmap(path, dtype=np.float32, mode='w+', shape=shape)
    mm[:] = a
    mm.flush()
    a_src = np.memmap(path, dtype=np.float32, mode='r', shape=shape)
  else:
    a_src = a
  dims = ('time', 'level', 'latitude', 'longitude')
  ds = xl.Dataset({'a': (dims, a_src), 'b': (dims, a * 2),
                   'orog': (dims[2:], a[0, 0])},
                  {'time': np.arange(ntime) * 6, 'level': np.arange(3),
                   'latitude': np.linspace(-90, 90, 5),
                   'longitude': np.arange(8) * 45.0})
  return ds, a


@pytest.mark.parametrize('chunk_size,depth,threads', [(1, 2, 1), (4, 1, 3),
                                                      (3, 3, 2)])
def test_chunks_keys_and_contents(tmp_path, chunk_size, depth, threads):
  ds, a = _dataset(tmp_path)
  f = fd.ChunkFeeder(ds, 'time', chunk_size, depth=depth, num_threads=threads,
                     pin=False)
  seen = []
  for key, chunk in f:
    s = key['time']
    seen.append(s)
    e = min(11, s + chunk_size)
    np.testing.assert_array_equal(chunk['a'].values, a[s:e])
    np.testing.assert_array_equal(chunk['b'].values, a[s:e] * 2)
    np.testing.assert_array_equal(chunk['time'].values, np.arange(s, e) * 6)
    assert chunk['orog'].dims == ('latitude', 'longitude')
    assert not isinstance(chunk['a'].data, np.memmap)  # a real copy
    assert chunk['a'].data.flags['C_CONTIGUOUS']
  assert seen == list(range(0, 11, chunk_size))
  assert f.bytes_read == 2 * a.nbytes
  assert len(f) == len(seen)


def test_indices_subset_and_validation():
  ds, a = _dataset()
  f = fd.ChunkFeeder(ds, 'time', 2, indices=[4, 1], pin=False)
  got = [(k['time'], c['a'].values.copy()) for k, c in f]
  assert [g[0] for g in got] == [8, 2]
  np.testing.assert_array_equal(got[0][1], a[8:10])
  with pytest.raises(IndexError):
    fd.ChunkFeeder(ds, 'time', 2, indices=[6], pin=False)
  with pytest.raises(ValueError):
    fd.ChunkFeeder(ds, 'nope', 2, pin=False)
  assert list(fd.ChunkFeeder(ds, 'time', 2, indices=[], pin=False)) == []


def test_prefetch_runs_ahead_but_is_bounded():
  """While the consumer holds chunk i, chunks i+1 .. i+depth are read -- no
  more (the pinned memory of a sweep stays bounded)."""
  ds, _ = _dataset(ntime=12)
  started, lock = [], threading.Lock()

  class Slow(fd.ChunkFeeder):

    def _load(self, ci, pool):
      with lock:
        started.append(ci)
      return super()._load(ci, pool)

  f = Slow(ds, 'time', 1, depth=3, pin=False)
  it = iter(f)
  next(it)
  time.sleep(0.2)
  with lock:
    n = len(started)
  assert 3 <= n <= 4, started  # chunk 0 delivered, 3 more in flight
  list(it)
  assert sorted(started) == list(range(12))


def test_reader_exception_reaches_the_consumer():
  ds, _ = _dataset()

  class Broken(fd.ChunkFeeder):

    def _read_variable(self, da, sl):
      if sl.start >= 4:
        raise OSError('disk gone')
      return super()._read_variable(da, sl)

  with pytest.raises(OSError, match='disk gone'):
    for _ in Broken(ds, 'time', 2, pin=False):
      pass
