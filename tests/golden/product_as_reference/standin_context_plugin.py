"""pytest plugin: every test runs with the NumPy stand-in context installed
(tests/fake_ctx.py), so the product's operators need no GPU."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(
    os.path.dirname(os.path.abspath(__file__))))))
import fake_ctx  # noqa: E402  pylint: disable=wrong-import-position


@pytest.fixture(autouse=True)
def _stand_in_context():
  with fake_ctx.installed():
    yield
