import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    import oracle_lib
    return oracle_lib.port()


@pytest.fixture(scope="session")
def ref():
    import oracle_lib
    r = oracle_lib.ref()
    if r is None:
        pytest.skip("oracle/_ref/librb_ref.so not available (needs /root/reference to build)")
    return r


@pytest.fixture(scope="session")
def gold():
    import golden_util
    return golden_util.load()
