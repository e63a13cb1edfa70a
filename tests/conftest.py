import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def checkers():
    """Build (if needed) and load the CPU checkers: the plain-C oracle always, the reference .so if present."""
    import cpu_checkers as cc
    cc.ensure_built()
    return cc
