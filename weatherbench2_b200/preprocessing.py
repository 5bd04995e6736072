"""Device-side preprocessing that feeds the metric kernels (SURVEY.md section 8
f3): the ensemble mean of scripts/compute_ensemble_mean.py:110-141
(`xbeam.Mean(realization, skipna)`) computed in HBM, so that ensemble-mean
RMSE / ACC through K1 need no host round trip of the mean."""
from __future__ import annotations

import numpy as np

from weatherbench2_b200 import _lib
from weatherbench2_b200 import xarray_lite as xl

REALIZATION = 'realization'


def compute_ensemble_mean(dataset, realization_name: str = REALIZATION,
                          skipna: bool = False):
  """Mean over `realization_name` of every variable that has it
  (scripts/compute_ensemble_mean.py:131: xbeam.Mean(REALIZATION_NAME, skipna)).

  float32 like xarray's mean of float32 data.  CUDA-tensor variables stay on
  the device (the result is a CUDA tensor K1 / K6 read in place); NumPy
  variables are uploaded, reduced and the (M times smaller) mean comes back.
  Variables without the dimension pass through.
  """
  native = xl.is_native_xarray(dataset)
  ds = xl.from_xarray(dataset)
  ctx = _lib.default_context()
  out = xl.Dataset(attrs=ds.attrs)
  for name in ds.keys():
    v = ds[name]
    if realization_name not in v.dims:
      out[name] = v
      continue
    dims = tuple(d for d in v.dims if d != realization_name)
    work = v.transpose(realization_name, *dims)
    data = work.data
    m = work.sizes[realization_name]
    shape = tuple(work.sizes[d] for d in dims)
    cells = int(np.prod(shape)) if shape else 1
    # one flat "field" per variable: the kernel walks it with a 64-bit
    # grid-stride loop, members `cells` elements apart
    slab = cells
    off = np.zeros(1, dtype=np.int64)
    coords = {k: c for k, c in work.coords.items()
              if realization_name not in c.dims}
    if xl._is_torch(data) and data.is_cuda:  # pylint: disable=protected-access
      import torch  # pylint: disable=import-outside-toplevel
      x = data.to(torch.float32).contiguous()
      res = torch.empty(shape, device=x.device, dtype=torch.float32)
      ctx.ens_mean(x.data_ptr(), m, cells, off, slab, skipna, res.data_ptr())
      values = res
    else:
      x = np.ascontiguousarray(np.asarray(data), dtype=np.float32)
      src = ctx.to_device(x)
      dst = ctx.malloc(max(4, cells * 4))
      try:
        ctx.ens_mean(src, m, cells, off, slab, skipna, dst)
        values = ctx.from_device(dst, shape, np.float32)
      finally:
        ctx.free(src)
        ctx.free(dst)
    out[name] = xl.DataArray(values, dims, coords, name, v.attrs)
  return xl.to_xarray(out) if native else out
