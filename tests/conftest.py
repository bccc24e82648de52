import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: there is no CPU fallback.
    pass


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    return oracle_lib.port()


@pytest.fixture(scope="session")
def ref():
    import oracle_lib

    r = oracle_lib.ref_vecsim()
    if r is None:
        pytest.skip("oracle/_ref/libvecsim_ref.so not built (needs /root/reference at build time)")
    return r
