"""`weatherbench2` as the reference's TEST FILES import it, but with the
PRODUCT's operators behind the names under test:

  weatherbench2.metrics / regions / thresholds / derived_variables /
  regridding / config / evaluation   ->  weatherbench2_b200.*
  weatherbench2.schema / utils / test_utils (mock data, random_like, ...)
                                      ->  the reference's own files

Used by tests/test_reference_suite_on_product.py: the reference's unit tests
then exercise this repository's operators (on the NumPy stand-in context; the
datasets are the stand-in xarray's, which are xarray_lite containers).
"""
import importlib
import importlib.util
import os
import sys

_REFERENCE = os.environ.get('WB2_REFERENCE_DIR', '/root/reference')

for _name in ('metrics', 'regions', 'thresholds', 'derived_variables',
              'regridding', 'config', 'evaluation'):
  _mod = importlib.import_module(f'weatherbench2_b200.{_name}')
  sys.modules[f'weatherbench2.{_name}'] = _mod
  globals()[_name] = _mod

for _name in ('schema', 'utils', 'test_utils'):
  _path = os.path.join(_REFERENCE, 'weatherbench2', f'{_name}.py')
  _spec = importlib.util.spec_from_file_location(f'weatherbench2.{_name}',
                                                 _path)
  _mod = importlib.util.module_from_spec(_spec)
  sys.modules[f'weatherbench2.{_name}'] = _mod
  _spec.loader.exec_module(_mod)
  globals()[_name] = _mod

if os.environ.get('WB2_STANDIN_CONTEXT') == '1':
  # a test file run with `python file.py` (absltest.main) instead of pytest:
  # no plugin hook, so the NumPy stand-in context is installed here for good
  import fake_ctx  # pylint: disable=wrong-import-position
  from weatherbench2_b200 import _lib as _wb2_lib  # pylint: disable=wrong-import-position
  _STANDIN = fake_ctx.FakeContext()
  _wb2_lib.default_context = lambda device=None: _STANDIN
