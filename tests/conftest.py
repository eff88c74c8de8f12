import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a wedged kernel must fail the test, not hang the box: default per-test timeout (pytest-timeout, if installed)
    if config.pluginmanager.hasplugin("timeout") and not getattr(config.option, "timeout", None):
        config.option.timeout = 600


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def within(name, measured, limit):
    """assert measured < limit, and -- when HAV_TEST_REPORT names a file -- append 'name measured limit' to it, so that the margins
    of the stated tolerances can be read off a GPU run (profiles/*_tolerance_margins.txt)."""
    measured = float(measured)
    path = os.environ.get("HAV_TEST_REPORT")
    if path:
        with open(path, "a") as f:
            f.write("%-70s measured %.3e   limit %.1e   margin %.1fx\n" % (name, measured, limit, limit / max(measured, 1e-30)))
    assert measured < limit, (name, measured, limit)
