import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the oracle port exists (gcc is available everywhere); the CUDA library and
    oracle/_ref are prebuilt by __graft_entry__.build() and travel with the snapshot."""
    from oracle import pyoracle as po
    if not os.path.exists(po.PORT_PATH):
        po.build_port()
    return True


@pytest.fixture(scope="session")
def gpu(built):
    from c3sc_b200 import capi
    L = capi.lib()          # raises loudly if the extension is missing
    if L.c3sc_cuda_device_count() == 0:
        pytest.skip("no CUDA device")
    capi.check(L.c3sc_cuda_init(0))
    return L
