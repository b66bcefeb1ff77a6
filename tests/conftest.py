import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def reference():
    import oracle_lib
    if not oracle_lib.have_reference():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return oracle_lib.Reference()


@pytest.fixture(scope="session")
def fixture_stack(golden):
    import cv2
    a = cv2.imdecode(golden["fixture_png"], cv2.IMREAD_UNCHANGED)
    assert a.shape == (2210, 640)
    return np.ascontiguousarray(a)
