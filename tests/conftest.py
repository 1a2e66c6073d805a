import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_paths():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.p3cs")))


def golden_ids():
    return [os.path.splitext(os.path.basename(p))[0] for p in golden_paths()]


@pytest.fixture(scope="session")
def gpu_ctx():
    """A p3cs context on cuda:0 through the C-ABI; fails loudly if the library or GPU is missing."""
    from panda3d_b200 import _capi
    ctx = _capi.Context(0)
    yield ctx
    ctx.close()
