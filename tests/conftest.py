import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "heart-sounds-segmentation_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device")


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree native artefacts exist (cross-compiles; no GPU needed)."""
    import __graft_entry__ as g

    g.build()
    return g


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
