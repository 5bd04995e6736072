"""Starts the reference's own test files as pytest subprocesses -- all of them
at once, the first time any result is asked for -- in the two configurations
of test_reference_suite_on_shim.py / test_reference_suite_on_product.py."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = '/root/reference'
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
FILES = ('metrics_test.py', 'regions_test.py', 'regridding_test.py',
         'derived_variables_test.py')
CONFIGS = {
    # the reference's modules on the stand-in libraries
    'shim': dict(path=[ROOT, os.path.join(GOLDEN, 'xarray_shim'), REFERENCE],
                 plugins=[]),
    # this repository's operators under the reference's module names
    'product': dict(path=[ROOT, os.path.join(ROOT, 'tests'),
                          os.path.join(GOLDEN, 'product_as_reference'),
                          os.path.join(GOLDEN, 'xarray_shim')],
                    plugins=['-p', 'standin_context_plugin']),
}
_running = {}


def available() -> bool:
  return os.path.isdir(os.path.join(REFERENCE, 'weatherbench2'))


def _start_all():
  tmp = tempfile.mkdtemp(prefix='wb2_reference_suite_')
  for config, spec in CONFIGS.items():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1',
               PYTHONPATH=os.pathsep.join(spec['path']),
               OMP_NUM_THREADS='1', OPENBLAS_NUM_THREADS='1')
    for name in FILES:
      log = open(os.path.join(tmp, f'{config}.{name}.log'), 'w+')
      proc = subprocess.Popen(
          [sys.executable, '-m', 'pytest',
           os.path.join(REFERENCE, 'weatherbench2', name), '-q', '-p',
           'no:cacheprovider'] + spec['plugins'],
          cwd=tmp, env=env, stdout=log, stderr=subprocess.STDOUT)
      _running[(config, name)] = (proc, log)
  _start_evaluation_test(tmp)


def _start_evaluation_test(tmp):
  """evaluation_test.py needs absltest's own runner (create_tempdir): run it
  with `python file.py`; the alias package installs the stand-in context."""
  spec = CONFIGS['product']
  env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1', WB2_STANDIN_CONTEXT='1',
             PYTHONPATH=os.pathsep.join(spec['path']), TEST_TMPDIR=tmp)
  log = open(os.path.join(tmp, 'product.evaluation_test.py.log'), 'w+')
  proc = subprocess.Popen(
      [sys.executable, os.path.join(REFERENCE, 'weatherbench2',
                                    'evaluation_test.py')],
      cwd=tmp, env=env, stdout=log, stderr=subprocess.STDOUT)
  _running[('product', 'evaluation_test.py')] = (proc, log)


def absltest_result(config: str, name: str):
  """(number of tests run, ok?, tail of the output) of a `python file.py` run."""
  if not _running:
    _start_all()
  proc, log = _running[(config, name)]
  proc.wait(timeout=1200)
  log.seek(0)
  out = log.read()
  ran = int((re.search(r'^Ran (\d+) test', out, re.M) or [0, 0])[1])
  ok = proc.returncode == 0 and re.search(r'^OK', out, re.M) is not None
  return ran, ok, out[-3000:]


def result(config: str, name: str):
  """(number passed, [names of failed tests], tail of the output)."""
  if not _running:
    _start_all()
  proc, log = _running[(config, name)]
  proc.wait(timeout=1200)
  log.seek(0)
  out = log.read()
  failed = re.findall(r'^FAILED \S+::(\w+)', out, re.M)
  passed = int((re.search(r'(\d+) passed', out) or [0, 0])[1])
  return passed, sorted(failed), out[-3000:]
