#!/usr/bin/env python
"""Runs the REFERENCE's own metric classes (/root/reference/weatherbench2/
metrics.py, regions.py, thresholds.py -- imported, not copied) on the seeded
inputs of reference_cases.py and stores what they return in
reference_run_vectors.npz.

xarray cannot be installed in this container (no network), so the reference
modules are executed on tests/golden/xarray_shim: weatherbench2_b200.xarray_lite
plus the remaining xarray calls they make, each written to xarray's documented
semantics.  The vectors are therefore "the reference's code on a re-implemented
xarray subset": they pin the REFERENCE logic that oracle/wb2_oracle.py restates
(which does not use that shim).  Only this script reads /root/reference; the
tests read the committed .npz.

Run:  python tests/golden/make_reference_vectors.py
"""
import os
import sys
import traceback
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(HERE, 'xarray_shim'), '/root/reference', HERE]

import xarray as xr  # noqa: E402  (the shim)
from weatherbench2 import (config, derived_variables, evaluation, metrics,  # noqa: E402
                           regions, regridding, thresholds)

import reference_cases as rc  # noqa: E402
import reference_eval_cases as rec  # noqa: E402


def run_evaluations(out) -> list:
  """weatherbench2.evaluation.evaluate_in_memory on the eval cases; datasets
  travel through the shim's in-memory "zarr" registry, results come back from
  the files the reference wrote."""
  import pickle
  import tempfile
  lib = types.SimpleNamespace(config=config, metrics=metrics, regions=regions,
                              Dataset=xr.Dataset)

  def source(name, dataset):
    xr.register_store(f'mem://{name}', dataset)
    return f'mem://{name}'

  failed = []
  with tempfile.TemporaryDirectory() as tmp:
    for case, data_config, eval_configs in rec.build(lib, source, tmp):
      for eval_name, eval_config in eval_configs.items():
        key = f'eval:{case}/{eval_name}'
        try:
          with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            evaluation.evaluate_in_memory(data_config,
                                          {eval_name: eval_config})
          with open(os.path.join(tmp, f'{eval_name}.nc'), 'rb') as f:
            res = pickle.load(f)
        except Exception:  # pylint: disable=broad-except
          failed.append(key)
          print('FAILED', key)
          traceback.print_exc(limit=-4)
          continue
        for var, (dims, values) in res['vars'].items():
          out[f"{key}|{var}|{','.join(dims)}"] = values
        print('ran', key, {v: d for v, (d, _) in res['vars'].items()})
  return failed


def main():
  lib = types.SimpleNamespace(metrics=metrics, regions=regions,
                              thresholds=thresholds)
  arr = rc.arrays()
  ds = rc.datasets(xr.Dataset, arr)
  out, failed = {}, []
  for case in rc.CASES:
    try:
      with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        res = rc.run_case(lib, case, ds, arr, xr.DataArray)
    except Exception:  # pylint: disable=broad-except
      failed.append(case['id'])
      print('FAILED', case['id'])
      traceback.print_exc(limit=-3)
      continue
    for var, (dims, values) in res.items():
      out[f"{case['id']}|{var}|{','.join(dims)}"] = values
  print(f'{len(rc.CASES) - len(failed)} / {len(rc.CASES)} cases ran')
  failed += run_evaluations(out)
  try:
    with warnings.catch_warnings():
      warnings.simplefilter('ignore', RuntimeWarning)
      extras = rc.run_extras(
          types.SimpleNamespace(metrics=metrics,
                                derived_variables=derived_variables,
                                regridding=regridding, regions=regions),
          xr.Dataset, arr)
    for name, (dims, values) in extras.items():
      out[f"extra:{name}||{','.join(dims)}"] = values
    print('ran', len(extras), 'extra calls')
  except Exception:  # pylint: disable=broad-except
    failed.append('extras')
    traceback.print_exc(limit=-6)
  if failed:
    sys.exit(1)
  np.savez_compressed(os.path.join(HERE, 'reference_run_vectors.npz'), **out)
  print('wrote', len(out), 'arrays')


if __name__ == '__main__':
  main()
