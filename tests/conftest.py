import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU oracle comparison")


@pytest.fixture(scope="session")
def built():
    """Make sure the CUDA library and the CPU oracle are built (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def golden_cases():
    from scripts import load_case
    gdir = os.path.join(ROOT, "tests", "golden")
    return {f[:-4]: load_case(os.path.join(gdir, f)) for f in sorted(os.listdir(gdir)) if f.endswith(".npz")}
