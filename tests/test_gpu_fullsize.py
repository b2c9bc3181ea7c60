"""Parity at BASELINE.json's FULL sizes (C2 P1 256^3, C3 P2 128^3, C4 Q1 192^3) on one B200.

The oracle cannot assemble 1e8 cells in seconds, so these tests use what the domain offers:

* counts against SURVEY.md App. A (dofs, nnz closed forms) - structure at full size;
* the oracle itself on a SAMPLE of rows: all cells incident to ~1500 randomly chosen (block) rows, plus rows on the
  Dirichlet boundary, are cut out of the big mesh, renumbered compactly, assembled by the CPU oracle, and the complete
  rows (column indices bit-exact, values row-scaled 1e-12) are compared with the same rows of the 1e8-cell GPU matrix
  (with bc rows/columns zeroed and set_diagonal applied);
* null space: A 1 = 0 (Poisson, cpp/test/matrix.cpp:96-109) / A r = 0 for the 6 rigid-body modes (elasticity,
  python/test/unit/la/test_nullspace.py:86-127);
* symmetry x.(A y) = y.(A x), linearity of MatrixCSR::mult;
* the aggregated strategy (chunked / row-gather) against the RED-per-contribution kernel;
* re-assembly without zeroing doubles the matrix (python/test/unit/fem/test_assembler.py:145-165);
* load vectors: sum_i b_i = int f for f = 1 (elasticity: f = e_0).
"""

import numpy as np
import pytest

from tests import problems as P  # noqa: F401
from tests import sampled_rows

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _edges(n):
    return 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n**3


# bench config: (n, dofs(n), nnz(n)) - closed forms of SURVEY.md App. A
CASES = {
    "p1": (256, lambda n: (n + 1) ** 3, lambda n: (n + 1) ** 3 + 2 * _edges(n)),
    "p2": (128, lambda n: (n + 1) ** 3 + _edges(n), lambda n: 230 * n**3 + 138 * n**2 + 24 * n + 1),
    "q1": (192, lambda n: (n + 1) ** 3, lambda n: (3 * n + 1) ** 3),
}


@pytest.fixture(scope="module")
def env():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import bench
    from dolfinx_b200 import _lib, common, fem, la

    class NS:
        pass

    ns = NS()
    ns.K, ns.common, ns.fem, ns.la, ns.torch, ns.bench = _lib, common, fem, la, torch, bench
    return ns


@pytest.mark.parametrize("cfg", ["p1", "p2", "q1"])
def test_full_size_config(env, oracle, cfg):
    torch, fem, la, K = env.torch, env.fem, env.la, env.K
    n, f_dofs, f_nnz = CASES[cfg]
    device = torch.device("cuda", 0)
    pb = env.bench.build_problem(cfg, n, env.common.Comm(), device)
    a, L, V, bc, bs = pb["a"], pb["L"], pb["V"], pb["bc"], pb["bs"]
    sp = fem.create_sparsity_pattern(a)
    sp.finalize()
    A = la.MatrixCSR(sp)

    # ---- structure: counts of SURVEY.md App. A ------------------------------------------------------------------
    assert pb["ndofs_global"] == f_dofs(n)
    assert A._nnz == f_nnz(n)
    indptr = A.indptr
    assert indptr[0] == 0 and indptr[-1] == A._nnz and np.all(np.diff(indptr) > 0)

    # ---- without bcs: null space, symmetry, linearity, strategies, re-assembly ----------------------------------
    fem.assemble_matrix(A, a)
    amax = float(A.data.abs().max())
    nrm2 = A.squared_norm()
    x = la.Vector(A.index_map(1), bs)
    y = la.Vector(A.index_map(0), bs)
    ndofs = pb["ndofs_local"]
    if bs == 1:
        modes = [torch.ones(ndofs, dtype=torch.float64, device=device)]
    else:
        # rigid-body modes at the dof nodes: 3 translations, 3 rotations
        dmap, xd = V.dofmap.dev, pb["mesh"].x_dofmap
        dof_x = torch.empty((ndofs, 3), dtype=torch.float64, device=device)
        dof_x[dmap.reshape(-1).long()] = pb["mesh"].x[xd.reshape(-1).long()]
        z = torch.zeros(ndofs, dtype=torch.float64, device=device)
        one = torch.ones_like(z)
        X, Y, Z = dof_x[:, 0], dof_x[:, 1], dof_x[:, 2]
        modes = [torch.stack(m, dim=1).reshape(-1) for m in
                 ((one, z, z), (z, one, z), (z, z, one), (-Y, X, z), (z, -Z, Y), (Z, z, -X))]
    for m in modes:
        x.array.copy_(m)
        y.set(0.0)
        A.mult(x, y)
        assert float(y.array.abs().max()) <= 64 * TOL * amax * float(m.abs().max()), "null space"
    g = torch.Generator(device=device)
    g.manual_seed(12345)
    u = torch.rand(ndofs * bs, generator=g, device=device, dtype=torch.float64)
    v = torch.rand(ndofs * bs, generator=g, device=device, dtype=torch.float64)

    def mult(vec):
        x.array.copy_(vec)
        y.set(0.0)
        A.mult(x, y)
        return y.array.clone()

    Au, Av = mult(u), mult(v)
    s1, s2 = float(torch.dot(v, Au)), float(torch.dot(u, Av))
    assert abs(s1 - s2) <= 1e-11 * float(torch.linalg.norm(v) * torch.linalg.norm(Au)), "symmetry"
    Auv = mult(2.0 * u - 3.0 * v)
    assert float((Auv - (2.0 * Au - 3.0 * Av)).abs().max()) <= 64 * TOL * float(Au.abs().max()), "linearity of mult"
    del Au, Av, Auv

    first = A.data.clone()
    A.set_value(0.0)
    fem.assemble_matrix(A, a, strategy=K.ASM_ATOMIC)
    diff = float((A.data - first).abs().max())
    assert diff <= TOL * amax, "aggregated strategy vs RED kernel"
    assert abs(A.squared_norm() - nrm2) <= TOL * nrm2
    fem.assemble_matrix(A, a, strategy=K.ASM_ATOMIC)  # no zeroing: doubles
    assert abs(A.squared_norm() - 4 * nrm2) <= 4 * TOL * nrm2
    del first

    # ---- with bcs: sample rows against the oracle ----------------------------------------------------------------
    A.set_value(0.0)
    fem.assemble_matrix(A, a, bcs=[bc])
    fem.set_diagonal(A, V, [bc], 1.0)
    mk = fem._bc_markers(V, [bc]).cpu().numpy()
    rng = np.random.default_rng(7)
    bdofs = bc._dofs0[:: max(1, len(bc._dofs0) // 300)] // bs
    rows = np.unique(np.concatenate([rng.integers(0, ndofs, 1200), bdofs, [0, ndofs - 1]])).astype(np.int64)
    if cfg == "p1":
        kid, consts = oracle.K_POISSON_P1_TET_A, np.array([2.0])
    elif cfg == "p2":
        kid, consts = oracle.K_POISSON_P2_TET_A, np.array([2.0])
    else:
        kid, consts = oracle.K_ELASTICITY_Q1_HEX_A, np.array([1.0e9 / 2.6, 1.0e9 * 0.3 / (1.3 * 0.4)])
    ud, pat, ref, ncells_sub = sampled_rows.oracle_rows(torch, oracle, pb, kid, consts, mk, rows)
    assert ncells_sub > len(rows)
    err = sampled_rows.compare_rows(A, rows, ud, pat, ref, bs)
    assert err <= TOL, f"sampled rows vs oracle: {err}"

    # ---- load vector: sum b = int f --------------------------------------------------------------------------------
    f = L.coefficients[0]
    b = la.Vector(V.dofmap.index_map, bs)
    if bs == 1:
        f.x.array.fill_(1.0)
        fem.assemble_vector(b, L)
        assert abs(float(b.array.sum()) - 1.0) <= 1e-11
    else:
        fv = f.x.array.view(-1, 3)
        fv.zero_()
        fv[:, 0] = 1.0
        fem.assemble_vector(b, L)
        bb = b.array.view(-1, 3).sum(dim=0).cpu().numpy()
        assert np.abs(bb - np.array([1.0, 0.0, 0.0])).max() <= 1e-11
