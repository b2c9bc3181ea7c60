"""MatrixCSR.matmul (la::matmul, la/matmul.h) on the GPU against the oracle and scipy, like
python/test/unit/la/test_matmul.py:21-132.

The per-row routine is also checked bitwise on the CPU (tests/test_matmul_row.py, same header)."""

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


def _explicit_matrix(la, common, torch, dense):
    """MatrixCSR with the pattern of the non-None entries of `dense` (a list of rows) and their values."""
    n0, n1 = len(dense), len(dense[0])
    im0, im1 = common.IndexMap(common.COMM_SELF, n0), common.IndexMap(common.COMM_SELF, n1)
    sp = la.SparsityPattern(common.COMM_SELF, [im0, im1], (1, 1))
    vals = []
    for i, row in enumerate(dense):
        cols = np.array([j for j, v in enumerate(row) if v is not None], dtype=np.int32)
        if cols.size:
            sp.insert(np.array([i]), cols)
        vals += [row[j] for j in cols]
    sp.finalize()
    A = la.MatrixCSR(sp)
    A.data.copy_(torch.tensor(vals, dtype=torch.float64))
    return A


def test_matmul_cancellation_and_stored_zero(oracle):
    """la/matmul.h:395-536: the structure of the product depends on the VALUES - products that are exactly zero (a
    stored zero) and sums that cancel exactly leave no entry."""
    import torch

    from dolfinx_b200 import common, la
    from tests.test_oracle_golden import _serial_omatrix

    A = _explicit_matrix(la, common, torch, [[1.0, -1.0, None], [0.0, 2.0, None], [None, None, 3.0]])
    B = _explicit_matrix(la, common, torch, [[2.0, 3.0, None], [2.0, 5.0, 1.0], [None, 0.0, 4.0]])
    Cm = A.matmul(B)
    As, Bs = A.to_scipy(), B.to_scipy()
    rp, od, cols, vals = oracle.matmul_local(_serial_omatrix(oracle, As), _serial_omatrix(oracle, Bs))
    assert np.array_equal(Cm.indptr, rp) and np.array_equal(Cm.indices, cols)
    assert np.array_equal(Cm.data.cpu().numpy(), vals)
    dense = Cm.to_scipy().toarray()
    assert np.allclose(dense, (As @ Bs).toarray(), rtol=0, atol=0)
    # row 0: (1)(2) + (-1)(2) cancels exactly -> no entry in column 0; row 1: the stored zero of A contributes nothing
    assert 0 not in Cm.indices[Cm.indptr[0]:Cm.indptr[1]]


def _random_matrix(la, common, torch, n0, n1, seed, density=0.4):
    rng = np.random.default_rng(seed)
    im0, im1 = common.IndexMap(common.COMM_SELF, n0), common.IndexMap(common.COMM_SELF, n1)
    sp = la.SparsityPattern(common.COMM_SELF, [im0, im1], (1, 1))
    for i in range(n0):
        cols = np.flatnonzero(rng.random(n1) < density)
        if cols.size:
            sp.insert(np.array([i]), cols)
    sp.finalize()
    A = la.MatrixCSR(sp)
    A.data.copy_(torch.from_numpy(rng.random(A.data.numel())))
    return A


@pytest.mark.parametrize("shape", [(17, 17, 17), (23, 11, 31), (1, 9, 4)])
def test_matmul_serial(oracle, shape):
    import torch

    from dolfinx_b200 import common, la
    from tests.test_oracle_golden import _serial_omatrix

    n, k, m = shape
    A = _random_matrix(la, common, torch, n, k, 1)
    B = _random_matrix(la, common, torch, k, m, 2)
    # exact cancellation and a stored zero, when the entries exist
    Cm = A.matmul(B)
    As, Bs = A.to_scipy(), B.to_scipy()
    ref = (As @ Bs).toarray()
    assert np.allclose(Cm.to_scipy().toarray(), ref, rtol=1e-14, atol=1e-15)
    rp, od, cols, vals = oracle.matmul_local(_serial_omatrix(oracle, As), _serial_omatrix(oracle, Bs))
    assert np.array_equal(Cm.indptr, rp) and np.array_equal(Cm.indices, cols)
    assert np.array_equal(Cm.data.cpu().numpy(), vals)  # bitwise
    assert Cm.index_map(0).size_local == n and Cm.index_map(1).size_local == m


def test_matmul_errors():
    import torch

    from dolfinx_b200 import common, la

    A = _random_matrix(la, common, torch, 5, 6, 1)
    B = _random_matrix(la, common, torch, 7, 5, 2)
    with pytest.raises(RuntimeError, match="Invalid matrix sizes for matmul."):
        A.matmul(B)
