"""Every entry point include/ps3d.h declares is exported by every library that implements it, and the Python
prototypes cover exactly that set. No compute call is made (no GPU needed)."""
import ctypes
import os
import re

import pytest

from conftest import ORACLE_SO, PRODUCT_SO, REF_SO, ROOT
from puresoft3d_b200 import _capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ps3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ps3d_[a-z0-9_]+)\s*\(", text)))


def test_prototypes_cover_the_header():
    assert declared_symbols() == sorted(_capi.PROTOTYPES)


@pytest.mark.parametrize("path", [PRODUCT_SO, ORACLE_SO, REF_SO], ids=["cuda", "oracle", "reference"])
def test_library_exports_every_symbol(path, built):
    if not os.path.exists(path):
        pytest.skip("%s not built here" % os.path.basename(path))
    lib = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_product_library_reports_cuda_backend(built):
    lib = _capi.bind(PRODUCT_SO)
    assert lib.ps3d_backend_name() == b"cuda-sm100a"


def test_product_has_no_cpu_fallback(built):
    """Without a CUDA device ps3d_create must fail (PS3D_ERR_DEVICE), never hand back a CPU pipe."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _capi.bind(PRODUCT_SO)
    h = ctypes.c_void_p()
    assert lib.ps3d_create(64, 64, 0, ctypes.byref(h)) == _capi.ERR_DEVICE
    assert not h.value
