import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "6dof-pose-estimation-and-defect-projection_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_path.npz"))


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def built_lib():
    _ensure_built()
    from defectproj import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def ctx(built_lib):
    from defectproj import Context
    c = Context(0)            # raises without a GPU: gpu tests never fall back
    yield c
    c.close()
