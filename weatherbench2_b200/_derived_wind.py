"""Wind derived variables -- same classes as
weatherbench2/derived_variables.py:59-99 -- computed on the device
(csrc/derived.cu).  Imported into `weatherbench2_b200.derived_variables`.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from weatherbench2_b200 import _lib
from weatherbench2_b200 import derived_variables as dv
from weatherbench2_b200 import xarray_lite as xl


@dataclasses.dataclass
class _WindVariable(dv.DerivedVariable):
  """A variable derived from the U and V wind components
  (derived_variables.py:59-74)."""

  u_name: str
  v_name: str

  @property
  def base_variables(self) -> list[str]:
    return [self.u_name, self.v_name]


@dataclasses.dataclass
class WindSpeed(_WindVariable):
  """Wind speed sqrt(u**2 + v**2) (derived_variables.py:77-99).  NumPy in ->
  NumPy out; CUDA tensors in -> a CUDA tensor that the metric kernels read in
  place."""

  @property
  def core_dims(self):
    return ([], []), []

  def compute(self, dataset):
    native = xl.is_native_xarray(dataset)
    ds = xl.from_xarray(dataset)
    u, v = ds[self.u_name], ds[self.v_name]
    if u.dims != v.dims or u.shape != v.shape:
      v = v.transpose(*u.dims)
      if u.shape != v.shape:
        raise ValueError(f'{self.u_name} and {self.v_name} differ in shape')
    ctx = _lib.default_context()
    ud, vd = u.data, v.data
    if xl._is_torch(ud) and ud.is_cuda:  # pylint: disable=protected-access
      import torch  # pylint: disable=import-outside-toplevel
      a = ud.to(torch.float32).contiguous()
      b = vd.to(torch.float32).contiguous().to(a.device)
      out = torch.empty_like(a)
      torch.cuda.current_stream(a.device).synchronize()
      ctx.wind_speed(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel())
      ctx.synchronize()
      data = out
    else:
      a = np.ascontiguousarray(np.asarray(u.values), dtype=np.float32)
      b = np.ascontiguousarray(np.asarray(v.values), dtype=np.float32)
      pa, pb = ctx.to_device(a), ctx.to_device(b)
      po = ctx.malloc(max(a.nbytes, 4))
      try:
        ctx.wind_speed(pa, pb, po, a.size)
        data = ctx.from_device(po, a.shape, np.float32)
      finally:
        for p in (pa, pb, po):
          ctx.free(p)
    coords = {k: c for k, c in u.coords.items()}
    result = xl.DataArray(data, u.dims, coords)
    if native:
      return xl.to_xarray(xl.Dataset({'_': result}))['_']
    return result
