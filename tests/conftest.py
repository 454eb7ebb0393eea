import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

ORACLE_SO = os.path.join(ROOT, "oracle", "libps3d_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libps3d_ref.so")
PRODUCT_SO = os.path.join(ROOT, "puresoft3d_b200", "libps3d_b200.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: minutes of CPU reference time; runs only with PS3D_SLOW=1")


@pytest.fixture(scope="session")
def built():
    """Build the native pieces once (the GPU box has no /root/reference: the prebuilt oracle/_ref .so travels)."""
    import __graft_entry__ as g
    g.build_product()   # returns at once when the .so is newer than every source; never test a stale library
    g.build_oracle()
    return g


@pytest.fixture(scope="session")
def oracle_lib(built):
    from puresoft3d_b200 import _capi
    return _capi.bind(ORACLE_SO)


@pytest.fixture(scope="session")
def ref_lib(built):
    from puresoft3d_b200 import _capi
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libps3d_ref.so not built (needs /root/reference)")
    return _capi.bind(REF_SO)


@pytest.fixture(scope="session")
def cuda_lib(built):
    from puresoft3d_b200 import _capi
    return _capi.load_product()
