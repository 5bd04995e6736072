"""Import-time stand-in for apache_beam: weatherbench2/evaluation.py defines
its Beam transforms at module level; tests/golden/make_reference_vectors.py
only calls the in-memory functions of that module."""


class PTransform:
  pass


class PCollection:
  pass


def __getattr__(name):
  raise AttributeError(f'apache_beam stand-in has no {name!r}: only the '
                       'in-memory evaluation of the reference is executed')
