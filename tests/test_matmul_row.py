"""The per-row routine of the device SpGEMM (dolfinx_b200/csrc/matmul_row.h, shared by the CUDA kernel and this host
build) against the oracle's restatement of impl::matmul (la/matmul.h:395-536): structure and values bitwise equal, on
one rank and on simulated ranks with fetched ghost rows, including zero products and exact cancellations."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.test_oracle_golden import _random_distributed_matrices, _serial_omatrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("mm") / "libmatmul_row_host.so")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC",
           os.path.join(ROOT, "tests", "cpp", "matmul_row_host.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(so)


def _run(lib, A, B, cm, grp, gcols, gvals):
    i64, i32, f64 = np.int64, np.int32, np.float64
    mB1 = B.index_maps[1]
    if cm is None:
        cm = mB1
        remap = (mB1.size_local + np.arange(mB1.num_ghosts)).astype(i32)
    else:
        g2l = {int(g): cm.size_local + i for i, g in enumerate(cm.ghosts)}
        remap = np.array([g2l[int(g)] for g in mB1.ghosts], dtype=i32)
    n = A.index_maps[0].size_local
    arr = dict(arp=A.row_ptr.astype(i64), aod=A.off_diag_offset.astype(i64), ac=A.cols.astype(i32), av=A.data.astype(f64),
               brp=B.row_ptr.astype(i64), bc=B.cols.astype(i32), bv=B.data.astype(f64), remap=np.append(remap, 0).astype(i32),
               grp=np.asarray(grp, dtype=i64), gc=np.append(np.asarray(gcols, dtype=i32), 0).astype(i32),
               gv=np.append(np.asarray(gvals, dtype=f64), 0.0))
    cap = 1 + int(sum((B.row_ptr[j + 1] - B.row_ptr[j]) for j in A.cols[A.cols < B.index_maps[0].size_local])) + len(arr["gc"]) * len(A.cols)
    crp, cod = np.zeros(n + 1, dtype=i64), np.zeros(n, dtype=i32)
    cc, cv = np.zeros(cap, dtype=i32), np.zeros(cap)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    st = lib.matmul_rows_host(C.c_int32(n), p(arr["arp"]), p(arr["aod"]), p(arr["ac"]), p(arr["av"]), p(arr["brp"]),
                              p(arr["bc"]), p(arr["bv"]), C.c_int32(B.index_maps[0].size_local), C.c_int32(mB1.size_local),
                              p(arr["remap"]), p(arr["grp"]), p(arr["gc"]), p(arr["gv"]), C.c_int32(cm.size_local),
                              C.c_int64(cap), p(crp), p(cod), p(cc), p(cv))
    assert st == 0
    return crp, cod, cc[: crp[-1]], cv[: crp[-1]]


@pytest.mark.parametrize("shape", [(7, 7, 7), (9, 5, 11), (12, 20, 6)])
def test_matmul_row_serial(oracle, hostlib, shape):
    import scipy.sparse as sps

    n, k, m = shape
    A = sps.random(n, k, density=0.5, random_state=3, format="lil", dtype=np.float64)
    B = sps.random(k, m, density=0.5, random_state=4, format="lil", dtype=np.float64)
    A[0, 0], A[0, 1] = 2.0, -4.0  # exact cancellation in C[0, 0]
    B[0, 0], B[1, 0] = 1.0, 0.5
    for j in range(2, k):
        B[j, 0] = 0.0
    A, B = A.tocsr(), B.tocsr()
    B = sps.csr_matrix((np.where(np.arange(B.nnz) == B.nnz - 1, 0.0, B.data), B.indices, B.indptr), shape=B.shape)  # a stored zero
    oA, oB = _serial_omatrix(oracle, A), _serial_omatrix(oracle, B)
    rp, od, cols, vals = oracle.matmul_local(oA, oB)
    crp, cod, cc, cv = _run(hostlib, oA, oB, None, np.zeros(1), np.zeros(0), np.zeros(0))
    assert np.array_equal(crp, rp) and np.array_equal(cod, od) and np.array_equal(cc, cols)
    assert np.array_equal(cv, vals)  # bitwise: same order of additions
    assert 0 not in cc[crp[0]:crp[1]]


@pytest.mark.parametrize("size", [2, 3, 4])
def test_matmul_row_with_ghost_rows(oracle, hostlib, size):
    rng = np.random.default_rng(99)
    A = _random_distributed_matrices(oracle, size, (1, 1), rng)
    ncA = [a.index_maps[1].size_local for a in A]
    B = _random_distributed_matrices(oracle, size, (1, 1), rng, nr=ncA, nc=[3 + r for r in range(size)])
    Cs, ghost = oracle.matmul(A, B, return_ghost_rows=True)
    for r in range(size):
        mA1 = A[r].index_maps[1]
        fast = len(mA1.src) == 0 and len(mA1.dest) == 0
        crp, cod, cc, cv = _run(hostlib, A[r], B[r], None if fast else Cs[r].index_maps[1], *ghost[r])
        assert np.array_equal(crp, Cs[r].row_ptr) and np.array_equal(cc, Cs[r].cols)
        assert np.array_equal(crp[:-1] + cod, Cs[r].off_diag_offset)
        assert np.array_equal(cv, Cs[r].data)
