"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/bfx.h declares (no compute calls without a GPU), and the product fails loudly without one."""

import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "bfx.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bfx_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from dolfinx_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(_lib.lib, n), f"libbfx.so does not export {n}"
    # and the ctypes table covers the header
    missing = [n for n in names if n not in _lib.EXPORTS]
    assert not missing, missing


def test_version_and_status_strings():
    from dolfinx_b200 import _lib

    assert _lib.lib.bfx_version() == 100
    assert _lib.lib.bfx_status_string(3).decode() == "Entry not in sparsity"
    ki = _lib.kernel_info(_lib.K_POISSON_P2_TET_A)
    assert (ki.nx, ki.nd, ki.bs, ki.rank) == (4, 10, 1, 2)
    ki = _lib.kernel_info(_lib.K_ELASTICITY_Q1_HEX_A)
    assert (ki.nx, ki.nd, ki.bs, ki.rank, ki.c_size) == (8, 8, 3, 2, 2)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly, never fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dolfinx_b200 import _lib, common, la

    n = C.c_int(0)
    st = _lib.lib.bfx_device_count(C.byref(n))
    assert st != 0 or n.value == 0
    p = C.c_void_p()
    assert _lib.lib.bfx_malloc(C.byref(p), 1024) != 0
    with pytest.raises(_lib.BfxError):
        la.Vector(common.IndexMap(common.COMM_SELF, 10), 1)


def test_host_first_touch():
    import numpy as np

    from dolfinx_b200 import _lib

    dm = np.array([[3, 1, 2], [1, 0, 3]], dtype=np.int32)
    new = _lib.host_first_touch(dm.reshape(-1), 5)
    assert new.tolist() == [3, 1, 2, 0, 4]


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under dolfinx_b200/ may reference it."""
    pkg = os.path.join(ROOT, "dolfinx_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|oracle/|oracle\.py|dlopen\([^)]*oracle)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f
