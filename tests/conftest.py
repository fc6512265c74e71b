import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _cuda_ok() -> bool:
    try:
        import ctypes as C

        from frankensearch_b200 import _ffi

        n = C.c_int(0)
        return _ffi.lib().fsgpu_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda_ok():
    return _cuda_ok()


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a usable device must FAIL, not skip silently: the product has no
    # CPU fallback.  Without `-m gpu` the marker expression already deselects GPU tests.
    return


@pytest.fixture(scope="session")
def fo():
    from oracle import fs_oracle

    fs_oracle.lib()
    return fs_oracle
