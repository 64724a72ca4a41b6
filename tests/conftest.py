"""pytest configuration: `gpu` marks tests that need a B200; everything else runs on CPU."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def msb():
    import inplacemsdradixsort_b200 as m
    m.load_library()
    return m


@pytest.fixture(scope="session")
def gpu(msb):
    if msb.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests must run on a GPU box")
    return msb
