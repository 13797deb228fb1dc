import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def ref_lib_available():
    from oracle import ref
    return ref.available("parity")


@pytest.fixture(autouse=True)
def no_pending_cuda_error(request):
    """every -m gpu test must leave the CUDA runtime without a pending error (hosts that share the runtime with the library,
    PyTorch for one, would report it at their next call)"""
    yield
    if request.node.get_closest_marker("gpu") is None:
        return
    from infinitam_b200 import capi
    if capi._lib is not None:
        err = capi._lib.itm_b200_take_cuda_error()
        assert err == 0, "the test left CUDA error %d pending" % err
