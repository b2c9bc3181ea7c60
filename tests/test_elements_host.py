"""The hand-written element kernels of the CUDA path (dolfinx_b200/csrc/elements.cuh), compiled for the HOST by g++
(tests/cpp/elements_host.cpp: the CUDA qualifiers expand to nothing outside nvcc), against the oracle's kernels, which
evaluate the same forms by numerical quadrature (oracle/oracle.c): closed-form pre-integration == quadrature on random
affine cells, entry by entry, for every local facet.  This puts the source the GPU executes under the CPU suite; the
GPU tests then only have to show that the same source gives the same numbers on the device.

Not covered here (device-only code paths, covered by tests/test_gpu_parity.py): the Q1 elasticity matrix kernels
(assemble.cu / rowgather.cu)."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from dolfinx_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# kernel ids shared by include/bfx.h and the oracle (test_kernel_id_tables_agree below)
KERNELS = [_lib.K_LAPLACE_P1_TRI_A, _lib.K_SOURCE_P1_TRI_L, _lib.K_MASS_COEFF_P1_TRI_A, _lib.K_LOAD_COEFF_P1_TRI_L,
           _lib.K_FACET_MASS_P1_TRI_A, _lib.K_FACET_CONST_P1_TRI_L, _lib.K_POISSON_P1_TET_A, _lib.K_LOAD_P1_TET_L,
           _lib.K_POISSON_P2_TET_A, _lib.K_LOAD_P2_TET_L, _lib.K_LOAD_Q1_HEX_L, _lib.K_FACET_LOAD_P1_TET_L,
           _lib.K_FACET_MASS_P1_TET_A, _lib.K_ACTION_POISSON_P1_TET_L, _lib.K_ACTION_POISSON_P2_TET_L,
           _lib.K_L2NORM2_P1_TET_M, _lib.K_COEFF2_P1_TRI_FACET_M, _lib.K_LOAD_PROD_P1_TET_L]
# interior-facet kernels: macro cells [cell0 | cell1], two local facet indices
DS_KERNELS = [_lib.K_AVG_MASS_P1_TRI_DS, _lib.K_AVG_LOAD_P1_TRI_DS_L, _lib.K_ONE_TRI_DS_M, _lib.K_AVG2_COEFF_P1_TRI_DS_M]


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("el") / "libelements_host.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-Wno-unknown-pragmas", "-ffp-contract=off", "-shared",
           "-fPIC", "-I", cuda_inc, os.path.join(ROOT, "tests", "cpp", "elements_host.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(so)
    lib.elements_host_tabulate.restype = C.c_int
    return lib


def _cell(nx, rng):
    """A random affine image of the reference cell (z = 0 for triangles)."""
    if nx == 3:
        ref = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0]])
        T = np.eye(3)
        T[:2, :2] = np.array([[1.0, 0.2], [-0.1, 0.8]]) + 0.2 * rng.standard_normal((2, 2))
        t = np.append(rng.standard_normal(2), 0.0)
    elif nx == 4:
        ref = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
        T = np.eye(3) + 0.25 * rng.standard_normal((3, 3))
        t = rng.standard_normal(3)
    else:
        ref = np.array([[(n >> 0) & 1, (n >> 1) & 1, (n >> 2) & 1] for n in range(8)], dtype=float)
        T = np.eye(3) + 0.2 * rng.standard_normal((3, 3))
        t = rng.standard_normal(3)
    if abs(np.linalg.det(T)) < 0.2:
        T = np.eye(3)
    return np.ascontiguousarray(ref @ T.T + t)


@pytest.mark.parametrize("kid", KERNELS)
def test_device_element_source_equals_oracle_quadrature(oracle, hostlib, kid):
    ki = _lib.kernel_info(kid)
    n = ki.nd * ki.bs
    nA = n * n if ki.rank == 2 else (n if ki.rank == 1 else 1)
    nfacets = {3: 3, 4: 4, 8: 6}[ki.nx] if ki.facet else 1
    rng = np.random.default_rng(1000 + kid)
    for trial in range(5):
        xc = _cell(ki.nx, rng)
        w = rng.standard_normal(max(ki.w_size, 1))
        c = 0.5 + rng.random(max(ki.c_size, 1))
        for lf in range(nfacets):
            out = np.zeros(nA)
            got = hostlib.elements_host_tabulate(C.c_int(kid), xc.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                                                 c.ctypes.data_as(C.c_void_p), C.c_int(lf), out.ctypes.data_as(C.c_void_p))
            assert got == nA
            ref = oracle.tabulate(kid, nA, xc, w, c, lf)
            scale = max(np.max(np.abs(ref)), 1e-300)
            assert np.max(np.abs(out - ref)) <= 1e-12 * scale, (kid, trial, lf)
            if ki.rank == 2 and ki.w_size == 0:
                A = out.reshape(n, n)
                assert np.max(np.abs(A - A.T)) <= 1e-14 * scale  # all bilinear benchmark forms are symmetric


def _macro_cell(rng, lf0, lf1, order1):
    """Two triangles sharing an edge: cell0 a random affine triangle whose local facet lf0 is the shared edge, cell1 the
    edge vertices (in the order `order1`) and a point on the other side placed at local index lf1."""
    c0 = _cell(3, rng)
    e = [v for v in range(3) if v != lf0]
    a, b = c0[e[0]], c0[e[1]]
    apex = a + b - c0[lf0] + 0.3 * (b - a) * rng.standard_normal()  # reflected through the edge midpoint, sheared
    c1 = np.zeros((3, 3))
    others = [v for v in range(3) if v != lf1]
    ends = (a, b) if order1 == 0 else (b, a)
    c1[others[0]], c1[others[1]], c1[lf1] = ends[0], ends[1], apex
    return np.ascontiguousarray(np.vstack([c0, c1]))


@pytest.mark.parametrize("kid", DS_KERNELS)
def test_interior_facet_element_source_equals_oracle(oracle, hostlib, kid):
    ki = _lib.kernel_info(kid)
    assert ki.nx == 6 and ki.facet
    n = ki.nd * ki.bs
    nA = n * n if ki.rank == 2 else (n if ki.rank == 1 else 1)
    rng = np.random.default_rng(2000 + kid)
    for lf0 in range(3):
        for lf1 in range(3):
            for order1 in range(2):
                xc = _macro_cell(rng, lf0, lf1, order1)
                w = rng.standard_normal(max(ki.w_size, 1))
                c = np.ones(1)
                out = np.zeros(nA)
                got = hostlib.elements_host_tabulate(C.c_int(kid), xc.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                                                     c.ctypes.data_as(C.c_void_p), C.c_int(lf0 + 8 * lf1),
                                                     out.ctypes.data_as(C.c_void_p))
                assert got == nA
                ref = oracle.tabulate(kid, nA, xc, w, c, (lf0, lf1))
                scale = max(np.max(np.abs(ref)), 1e-300)
                assert np.max(np.abs(out - ref)) <= 1e-12 * scale, (kid, lf0, lf1, order1)


def test_kernel_id_tables_agree(oracle):
    """The kernel ids of the product (include/bfx.h, dolfinx_b200._lib) and of the oracle name the same forms."""
    names = [k for k in dir(_lib) if k.startswith("K_") and isinstance(getattr(_lib, k), int)]
    assert len(names) >= 18
    for k in names:
        assert getattr(oracle, k) == getattr(_lib, k), k
