"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same inputs.

Tolerance (BASELINE.json north_star: 1e-12 relative, fp64, summation order differs under atomics),
read as SURVEY.md §8c prescribes: |a_ij - ref_ij| <= 1e-12 * max_k |ref_ik| (row-scaled) and
||A - Aref||_F <= 1e-12 ||Aref||_F.  Structure (row_ptr, cols, off-diagonal offsets) is bit-exact.
"""

import numpy as np
import pytest

from tests import problems as P

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def bfx():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dolfinx_b200 import _lib, common, fem, la

    class NS:
        pass

    ns = NS()
    ns.lib, ns.common, ns.fem, ns.la, ns.torch = _lib, common, fem, la, torch
    return ns


def make_space(bfx, p):
    comm = bfx.common.COMM_SELF
    msh = bfx.fem.Mesh(comm, p.x, p.x_dofmap, p.cell)
    im = bfx.common.IndexMap(comm, p.ndofs)
    V = bfx.fem.FunctionSpace(msh, "Lagrange", bfx.fem.DofMap(p.dofmap, p.bs, im))
    return msh, V


def assemble_A(bfx, V, kernel, constants=(), coefficients=(), active=(), bcs=(), strategy=None):
    fem, la = bfx.fem, bfx.la
    a = fem.Form([V, V], {fem.IntegralType.cell: [(0, kernel, None, list(active))]},
                 coefficients=coefficients, constants=[fem.Constant(c) for c in constants])
    sp = fem.create_sparsity_pattern(a)
    sp.finalize()
    A = la.MatrixCSR(sp)
    fem.assemble_matrix(A, a, bcs=bcs, strategy=strategy)
    return a, sp, A


def check_matrix(A, pat, ref_data, bs2=1):
    edges, offsets = A.indices, A.indptr
    assert np.array_equal(offsets, pat.offsets), "row_ptr must be bit-exact"
    assert np.array_equal(edges, pat.edges), "column indices must be bit-exact"
    data = A.data.cpu().numpy()
    assert P.row_scaled_error(data, ref_data, pat.offsets, bs2) <= TOL
    assert np.linalg.norm(data - ref_data) <= TOL * np.linalg.norm(ref_data)


@pytest.mark.parametrize("numbering", ["lex", "first_touch", "random"])
@pytest.mark.parametrize("n", [1, 3, 8])
def test_poisson_p1_matrix(bfx, oracle, n, numbering):
    p = P.tet_p1(n, numbering=numbering, seed=n)
    msh, V = make_space(bfx, p)
    a, sp, A = assemble_A(bfx, V, bfx.lib.K_POISSON_P1_TET_A, constants=[2.0])
    pat, ref = P.oracle_assemble_matrix(oracle, p, oracle.K_POISSON_P1_TET_A, constants=np.array([2.0]))
    check_matrix(A, pat, ref)
    assert np.array_equal(sp.off_diagonal_offsets, pat.off_diagonal_offsets)
    assert A.squared_norm() == pytest.approx(float(np.sum(ref**2)), rel=1e-12)


@pytest.mark.parametrize("alt_cb", [0, 96, 192, 384, "two_stage", "two_stage_split"])
@pytest.mark.parametrize("symmetric", [True, False])
@pytest.mark.parametrize("case", ["p1_lex", "p1_random", "p2", "tri"])
def test_chunked_and_atomic_strategies(bfx, oracle, case, symmetric, alt_cb, monkeypatch):
    """Both scatter-add strategies against the oracle on meshes spanning many chunks: the
    chunk-aggregated kernel (default; complete destinations by plain update, chunk-boundary ones by
    RED) and the plain fp64-RED kernel; then re-assembly without zeroing (values += , the
    "does not zero" contract of fem/assembler.h:497-498, test_assembler.py:157-165)."""
    fem, K = bfx.fem, bfx.lib
    # symmetric plan: upper triangle staged, (i,j)/(j,i) entries updated from one sum; general plan otherwise
    monkeypatch.setattr(fem, "CHUNKS_SYMMETRIC", symmetric)
    # cells per chunk (BFX_CHUNKS_CB): 96 / 128 / 192 / 384 for the P1 kernels, 64 / 96 for symmetric P2, else the default
    # two_stage: write-back of the chunk sums in address order (symmetric P1 plans of 256-cell chunks)
    two_stage = {"two_stage": 1, "two_stage_split": 2}.get(alt_cb, 0)
    alt_cb = 0 if two_stage else alt_cb
    monkeypatch.setattr(fem, "CHUNKS_TWO_STAGE", two_stage)
    monkeypatch.setattr(fem, "CHUNKS_CB", alt_cb)
    if case.startswith("p1"):
        p = P.tet_p1(13, numbering="lex" if case == "p1_lex" else "random", seed=5)
        kern, okern, consts = K.K_POISSON_P1_TET_A, oracle.K_POISSON_P1_TET_A, [2.0]
    elif case == "p2":
        p = P.tet_p2(6)
        kern, okern, consts = K.K_POISSON_P2_TET_A, oracle.K_POISSON_P2_TET_A, [2.0]
    else:
        p = P.tri_p1(40, 37)
        kern, okern, consts = K.K_LAPLACE_P1_TRI_A, oracle.K_LAPLACE_P1_TRI_A, []
    msh, V = make_space(bfx, p)
    pat, ref = P.oracle_assemble_matrix(oracle, p, okern, constants=np.array(consts))
    for strategy in (K.ASM_CHUNKED, K.ASM_ATOMIC):
        a, sp, A = assemble_A(bfx, V, kern, constants=consts, strategy=strategy)
        check_matrix(A, pat, ref)
        fem.assemble_matrix(A, a, strategy=strategy)  # A is not zero any more: add mode
        assert np.linalg.norm(A.data.cpu().numpy() - 2 * ref) <= TOL * np.linalg.norm(2 * ref)
        if strategy == K.ASM_CHUNKED:
            nchunks, ndest, nsrc, nbytes = fem.chunk_stats(a, A)
            assert nchunks > 4 and nsrc >= ref.size // 2 and ndest >= (len(pat.edges) // 2 if symmetric else len(pat.edges))
            if case == "p2":
                cb = (alt_cb if alt_cb == 96 else 128) if symmetric else 64
            else:
                # symmetric P1-sized plans default to the lean kernel's 384-cell chunks (fem.CHUNK_LEAN)
                cb = alt_cb if alt_cb else (384 if (symmetric and fem.CHUNK_LEAN and not two_stage) else 256)
            # (the lean default is a preference: a mesh without complete warp tables keeps the classic 256 cells)
            assert nchunks == -(-len(p.dofmap) // cb) or (cb == 384 and not alt_cb and nchunks == -(-len(p.dofmap) // 256))
            if case.startswith("p1"):  # (triangles: more destinations per chunk than the two-stage kernel holds)
                assert fem.chunk_two_stage(a, A) == bool(two_stage and symmetric)
    # default strategy = chunk-aggregated for the P1 kernels
    a, sp, A = assemble_A(bfx, V, kern, constants=consts)
    if kern in K.CHUNKED_KERNELS:
        assert fem.chunk_stats(a, A)[0] > 4
    check_matrix(A, pat, ref)


def test_poisson_p1_32_config1(bfx, oracle):
    """BASELINE configs[0]: P1 Laplacian on create_box 32^3 tets — matrix + vector into MatrixCSR."""
    fem, la = bfx.fem, bfx.la
    p = P.tet_p1(32, numbering="first_touch")
    msh, V = make_space(bfx, p)
    a, sp, A = assemble_A(bfx, V, bfx.lib.K_POISSON_P1_TET_A, constants=[2.0])
    pat, ref = P.oracle_assemble_matrix(oracle, p, oracle.K_POISSON_P1_TET_A, constants=np.array([2.0]))
    assert len(pat.edges) == 513313 and p.ndofs == 35937
    check_matrix(A, pat, ref)
    f = fem.Function(V)
    fh = P.source_f(p.dof_coords)
    f.x.array.copy_(bfx.torch.from_numpy(fh))
    L = fem.Form([V], {fem.IntegralType.cell: [(0, bfx.lib.K_LOAD_P1_TET_L, None, [0])]}, coefficients=[f])
    b = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b, L)
    bref = P.oracle_assemble_vector(oracle, p, oracle.K_LOAD_P1_TET_L, coeff=(fh, p.dofmap, 1))
    assert np.max(np.abs(b.array.cpu().numpy() - bref)) <= TOL * np.max(np.abs(bref))
    assert la.norm(b) == pytest.approx(np.linalg.norm(bref), rel=1e-12)


def test_golden_scalars_on_gpu(bfx):
    """The reference's golden numbers reproduced by the CUDA path itself
    (test_custom_jit_kernels.py:115-116, test_ghost_mesh_assembly.py:64-66)."""
    fem, la, K = bfx.fem, bfx.la, bfx.lib
    from dolfinx_b200 import mesh as M

    p = P.tri_p1(13, 13)
    msh, V = make_space(bfx, p)
    a, sp, A = assemble_A(bfx, V, K.K_LAPLACE_P1_TRI_A)
    A.scatter_reverse()
    L = fem.Form([V], {fem.IntegralType.cell: [(0, K.K_SOURCE_P1_TRI_L, None, [])]})
    b = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b, L)
    b.scatter_reverse(la.InsertMode.add)
    assert np.isclose(np.sqrt(A.squared_norm()), 56.124860801609124, rtol=1e-13)
    assert np.isclose(la.norm(b), 0.0739710713711999, rtol=1e-13)

    p = P.tri_p1(12, 12)
    msh, V = make_space(bfx, p)
    f = fem.Function(V)
    f.x.array.fill_(10.0)
    ents = M.exterior_facets(p.x_dofmap, M.TRI_FACETS)
    a = fem.Form([V, V], {fem.IntegralType.cell: [(0, K.K_MASS_COEFF_P1_TRI_A, None, [0])],
                          fem.IntegralType.exterior_facet: [(0, K.K_FACET_MASS_P1_TRI_A, ents, [])]}, coefficients=[f])
    L = fem.Form([V], {fem.IntegralType.cell: [(0, K.K_LOAD_COEFF_P1_TRI_L, None, [0])],
                       fem.IntegralType.exterior_facet: [(0, K.K_FACET_CONST_P1_TRI_L, ents, [])]},
                 coefficients=[f], constants=[fem.Constant(2.0)])
    sp = fem.create_sparsity_pattern(a)
    sp.finalize()
    A = la.MatrixCSR(sp)
    fem.assemble_matrix(A, a)
    A.scatter_reverse()
    b = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b, L)
    b.scatter_reverse(la.InsertMode.add)
    assert np.sqrt(A.squared_norm()) == pytest.approx(0.6713621455570528, rel=1e-12)
    assert la.norm(b) == pytest.approx(1.582294032953906, rel=1e-12)


@pytest.mark.parametrize("n", [2, 6])
def test_poisson_p2_matrix_and_vector(bfx, oracle, n):
    fem, la = bfx.fem, bfx.la
    p = P.tet_p2(n)
    msh, V = make_space(bfx, p)
    a, sp, A = assemble_A(bfx, V, bfx.lib.K_POISSON_P2_TET_A, constants=[2.0])
    pat, ref = P.oracle_assemble_matrix(oracle, p, oracle.K_POISSON_P2_TET_A, constants=np.array([2.0]))
    check_matrix(A, pat, ref)
    # cpp/test/matrix.cpp:66-120: A.1 = 0
    x = la.Vector(A.index_map(1), 1)
    y = la.Vector(A.index_map(0), 1)
    x.set(1.0)
    A.mult(x, y)
    assert float(y.array.abs().max()) < 1e-13 * float(np.max(np.abs(ref))) * 10
    f = fem.Function(V)
    fh = P.source_f(p.dof_coords)
    f.x.array.copy_(bfx.torch.from_numpy(fh))
    L = fem.Form([V], {fem.IntegralType.cell: [(0, bfx.lib.K_LOAD_P2_TET_L, None, [0])]}, coefficients=[f])
    b = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b, L)
    bref = P.oracle_assemble_vector(oracle, p, oracle.K_LOAD_P2_TET_L, coeff=(fh, p.dofmap, 1))
    assert np.max(np.abs(b.array.cpu().numpy() - bref)) <= TOL * np.max(np.abs(bref))


@pytest.mark.parametrize("skew", [False, True])
def test_elasticity_q1_matrix_and_vector(bfx, oracle, skew):
    fem, la = bfx.fem, bfx.la
    p = P.hex_q1(4, numbering="random", seed=5, skew=skew)
    msh, V = make_space(bfx, p)
    E, nu = 1.0e9, 0.3
    mu, lmbda = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    a, sp, A = assemble_A(bfx, V, bfx.lib.K_ELASTICITY_Q1_HEX_A, constants=[[mu, lmbda]])
    pat, ref = P.oracle_assemble_matrix(oracle, p, oracle.K_ELASTICITY_Q1_HEX_A, constants=np.array([mu, lmbda]))
    check_matrix(A, pat, ref, bs2=9)
    # f = rho omega^2 (x0, x1, 0), python/demo/demo_elasticity.py:123-125
    fh = np.zeros((p.ndofs, 3))
    fh[:, 0], fh[:, 1] = 10 * 300**2 * p.dof_coords[:, 0], 10 * 300**2 * p.dof_coords[:, 1]
    f = fem.Function(V)
    f.x.array.copy_(bfx.torch.from_numpy(fh.reshape(-1)))
    L = fem.Form([V], {fem.IntegralType.cell: [(0, bfx.lib.K_LOAD_Q1_HEX_L, None, [0])]}, coefficients=[f])
    b = la.Vector(V.dofmap.index_map, 3)
    fem.assemble_vector(b, L)
    bref = P.oracle_assemble_vector(oracle, p, oracle.K_LOAD_Q1_HEX_L, coeff=(fh.reshape(-1), p.dofmap, 3))
    assert np.max(np.abs(b.array.cpu().numpy() - bref)) <= TOL * np.max(np.abs(bref))


@pytest.mark.parametrize("n,skew,perturb", [(7, False, 0.0), (6, True, 0.0), (5, False, 0.3)])
def test_q1_rowgather_and_atomic_strategies(bfx, oracle, n, skew, perturb):
    """The row-gather elasticity kernel (default for Q1: every CSR value written once, no atomics) and
    the cell-parallel fp64-RED kernel against the oracle: affine, skewed (parallelepiped) and partly
    non-affine meshes (the non-affine remainder goes through the RED kernel), several 32-row tiles;
    re-assembly adds; the row-gather result is bitwise reproducible."""
    fem, K = bfx.fem, bfx.lib
    p = P.hex_q1(n, numbering="random", seed=n, skew=skew, perturb=perturb)
    if perturb:
        keep = p.x[:, 2] < 0.5
        p.x[keep] = P.hex_q1(n, numbering="random", seed=n).x[keep]
    msh, V = make_space(bfx, p)
    consts = [[1.0, 1.5]]
    okern = oracle.K_ELASTICITY_Q1_HEX_A_G2 if perturb else oracle.K_ELASTICITY_Q1_HEX_A
    pat, ref = P.oracle_assemble_matrix(oracle, p, okern, constants=np.array([1.0, 1.5]))
    runs = []
    for strategy in (K.ASM_ROWGATHER, K.ASM_ROWGATHER, K.ASM_ATOMIC):
        a, sp, A = assemble_A(bfx, V, K.K_ELASTICITY_Q1_HEX_A, constants=consts, strategy=strategy)
        check_matrix(A, pat, ref, bs2=9)
        runs.append(A.data.cpu().numpy().copy())
        fem.assemble_matrix(A, a, strategy=strategy)  # add mode
        assert np.linalg.norm(A.data.cpu().numpy() - 2 * ref) <= TOL * np.linalg.norm(2 * ref)
    if not perturb:
        assert np.array_equal(runs[0], runs[1]), "row-gather assembly must be bitwise reproducible"


def test_elasticity_q1_nonaffine_cells(bfx, oracle):
    """General trilinear hexahedra take the 2x2x2 Gauss path of the CUDA kernel: compare with the oracle's
    2x2x2 variant to 1e-12 and with its 3x3x3 rule to quadrature accuracy; a mesh mixing affine and
    perturbed cells exercises the per-cell switch."""
    fem, la = bfx.fem, bfx.la
    p = P.hex_q1(4, numbering="random", seed=6, perturb=0.3)
    half = p.x[:, 2] < 0.4  # keep the lower layers affine
    p.x[half] = P.hex_q1(4, numbering="random", seed=6).x[half]
    msh, V = make_space(bfx, p)
    a, sp, A = assemble_A(bfx, V, bfx.lib.K_ELASTICITY_Q1_HEX_A, constants=[[1.0, 1.5]])
    pat, ref2 = P.oracle_assemble_matrix(oracle, p, oracle.K_ELASTICITY_Q1_HEX_A_G2, constants=np.array([1.0, 1.5]))
    check_matrix(A, pat, ref2, bs2=9)
    pat, ref3 = P.oracle_assemble_matrix(oracle, p, oracle.K_ELASTICITY_Q1_HEX_A, constants=np.array([1.0, 1.5]))
    data = A.data.cpu().numpy()
    assert np.linalg.norm(data - ref3) <= 2e-2 * np.linalg.norm(ref3)
    assert np.linalg.norm(data - ref3) > 1e-9 * np.linalg.norm(ref3)  # the mesh really is non-affine


@pytest.mark.parametrize("case", ["p1", "q1"])
def test_bcs_lifting_set_diagonal(bfx, oracle, case):
    """bc row/col zeroing, set_diagonal, apply_lifting, set_bc against the oracle
    (fem/assemble_matrix_impl.h:161-196, assembler.h:336-493,644-686, DirichletBC.h:495-578)."""
    fem, la, O = bfx.fem, bfx.la, oracle
    torch = bfx.torch
    if case == "p1":
        p = P.tet_p1(5, numbering="random", seed=2)
        kA, consts = bfx.lib.K_POISSON_P1_TET_A, [2.0]
        okA = O.K_POISSON_P1_TET_A
    else:
        p = P.hex_q1(4, numbering="first_touch")
        kA, consts = bfx.lib.K_ELASTICITY_Q1_HEX_A, [[1.0, 1.5]]
        okA = O.K_ELASTICITY_Q1_HEX_A
    bs = p.bs
    msh, V = make_space(bfx, p)
    bnodes = np.flatnonzero(np.isclose(p.dof_coords[:, 0], 0.0) | np.isclose(p.dof_coords[:, 1], 1.0)).astype(np.int32)
    g = fem.Function(V)
    gh = np.random.default_rng(0).random(p.ndofs * bs)
    g.x.array.copy_(torch.from_numpy(gh))
    bc = fem.DirichletBC(g, bnodes)
    dofs_unrolled, n_owned = bc.dof_indices()
    assert np.array_equal(dofs_unrolled, O.unroll_dofs(bnodes, bs)) and n_owned == len(dofs_unrolled)
    a, sp, A = assemble_A(bfx, V, kA, constants=consts, bcs=[bc])
    fem.set_diagonal(A, V, [bc], 1.0)
    markers = np.zeros(p.ndofs * bs, dtype=np.int8)
    O.bc_mark(markers, dofs_unrolled)
    cc = np.asarray(consts, dtype=np.float64).reshape(-1)
    pat, ref = P.oracle_assemble_matrix(O, p, okA, constants=cc, bc=markers)
    O.set_diagonal(ref, pat.edges, pat.offsets, bs, bs, dofs_unrolled, 1.0)
    check_matrix(A, pat, ref, bs2=bs * bs)
    # lifting with x0 and alpha
    x0h = np.random.default_rng(1).random(p.ndofs * bs)
    x0 = torch.from_numpy(x0h).cuda()
    b = la.Vector(V.dofmap.index_map, bs)
    b.array.copy_(torch.from_numpy(np.arange(p.ndofs * bs, dtype=np.float64)))
    fem.apply_lifting(b, [a], [[bc]], x0=[x0], alpha=0.5)
    bref = np.arange(p.ndofs * bs, dtype=np.float64)
    values = np.zeros(p.ndofs * bs)
    O.bc_set(values, dofs_unrolled, gh, 0, bs)
    O.lift_bc(okA, p.x_dofmap, p.x, np.arange(len(p.dofmap)), p.dofmap, bs, p.dofmap, bs, bref, values, markers,
              x0=x0h, alpha=0.5, constants=cc)
    scale = np.max(np.abs(bref - np.arange(p.ndofs * bs)))
    assert np.max(np.abs(b.array.cpu().numpy() - bref)) <= TOL * max(scale, 1.0)
    fem.set_bc(b, [bc], x0=x0, alpha=0.5)
    O.bc_set(bref, dofs_unrolled, gh, 0, bs, x0=x0h, alpha=0.5)
    assert np.max(np.abs(b.array.cpu().numpy() - bref)) <= TOL * max(scale, 1.0)
    # Constant-valued bc: value[dof % bs]
    cval = np.arange(1, bs + 1, dtype=np.float64) * 1e-3
    bcc = fem.DirichletBC(fem.Constant(cval), bnodes, V)
    xx = torch.zeros(p.ndofs * bs, dtype=torch.float64, device="cuda")
    bcc.set(xx, None, 2.0)
    xr = np.zeros(p.ndofs * bs)
    O.bc_set(xr, dofs_unrolled, cval, 1, bs, alpha=2.0)
    assert np.array_equal(xx.cpu().numpy(), xr)


def test_reassembly_and_packed_coefficients(bfx, oracle):
    """'does not zero' contract (test_assembler.py:145-165) and pre-packed coefficients (:1111-1205)."""
    fem, la = bfx.fem, bfx.la
    p = P.tet_p1(4)
    msh, V = make_space(bfx, p)
    a, sp, A = assemble_A(bfx, V, bfx.lib.K_POISSON_P1_TET_A, constants=[1.0])
    n1 = A.squared_norm()
    fem.assemble_matrix(A, a)
    assert A.squared_norm() == pytest.approx(4 * n1, rel=1e-13)
    f = fem.Function(V)
    fh = P.source_f(p.dof_coords)
    f.x.array.copy_(bfx.torch.from_numpy(fh))
    L = fem.Form([V], {fem.IntegralType.cell: [(0, bfx.lib.K_LOAD_P1_TET_L, None, [0])]}, coefficients=[f])
    b1 = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b1, L)
    packed = fem.pack_coefficients(L)
    c, cstride = packed[(fem.IntegralType.cell, 0)]
    ref = np.zeros((len(p.dofmap), 4))
    oracle.pack_coefficient(ref, 0, fh, p.dofmap, 1, cells=np.arange(len(p.dofmap)))
    assert np.array_equal(c.cpu().numpy(), ref)
    b2 = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b2, L, coeffs=packed)
    fem.assemble_vector(b2, L, coeffs=packed)
    assert np.allclose(b2.array.cpu().numpy(), 2 * b1.array.cpu().numpy(), rtol=1e-13, atol=0)


def test_exterior_facets_tet(bfx, oracle):
    fem, la, K, O = bfx.fem, bfx.la, bfx.lib, oracle
    from dolfinx_b200 import mesh as M

    p = P.tet_p1(4, numbering="random", seed=4)
    msh, V = make_space(bfx, p)
    ents = M.exterior_facets(p.x_dofmap, M.TET_FACETS)
    g = fem.Function(V)
    gh = np.sin(5 * p.dof_coords[:, 0])
    g.x.array.copy_(bfx.torch.from_numpy(gh))
    L = fem.Form([V], {fem.IntegralType.exterior_facet: [(0, K.K_FACET_LOAD_P1_TET_L, ents, [0])]}, coefficients=[g])
    b = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b, L)
    coeffs = np.zeros((len(ents), 4))
    O.pack_coefficient(coeffs, 0, gh, p.dofmap, 1, entities=ents)
    bref = np.zeros(p.ndofs)
    O.assemble_vector(O.K_FACET_LOAD_P1_TET_L, p.x_dofmap, p.x, None, p.dofmap, 1, bref, coeffs=coeffs, entities=ents)
    assert np.max(np.abs(b.array.cpu().numpy() - bref)) <= TOL * np.max(np.abs(bref))
    a = fem.Form([V, V], {fem.IntegralType.cell: [(0, K.K_POISSON_P1_TET_A, None, [])],
                          fem.IntegralType.exterior_facet: [(0, K.K_FACET_MASS_P1_TET_A, ents, [])]},
                 constants=[fem.Constant(1.0)])
    sp = fem.create_sparsity_pattern(a)
    sp.finalize()
    A = la.MatrixCSR(sp)
    fem.assemble_matrix(A, a)
    pat, ref = P.oracle_assemble_matrix(O, p, O.K_POISSON_P1_TET_A, constants=np.array([1.0]))
    O.assemble_matrix(O.K_FACET_MASS_P1_TET_A, p.x_dofmap, p.x, None, p.dofmap, 1, p.dofmap, 1, ref, pat.edges,
                      pat.offsets, entities=ents)
    check_matrix(A, pat, ref)


@pytest.mark.parametrize("bs", [(1, 1), (2, 2), (1, 2), (2, 1), (2, 3), (3, 3)])
def test_spmv_all_block_sizes(bfx, oracle, bs):
    """python/test/unit/la/test_matrix_vector.py:47-106 (float64) against scipy and the oracle."""
    import scipy.sparse as sps

    la, common, torch = bfx.la, bfx.common, bfx.torch
    n = 30
    im = common.IndexMap(common.COMM_SELF, n)
    sp = la.SparsityPattern(common.COMM_SELF, [im, im], bs)
    sp.insert(np.arange(n), np.arange(n))
    sp.finalize()
    A = la.MatrixCSR(sp)
    rng = np.random.default_rng(12345)
    A.data.copy_(torch.from_numpy(rng.random(A.data.numel())))
    As = A.to_scipy()
    b = la.Vector(im, bs[1])
    u = la.Vector(im, bs[0])
    b.array.copy_(torch.arange(b.array.numel(), dtype=torch.float64))
    A.mult(b, u)
    assert np.allclose(u.array.cpu().numpy(), As @ b.array.cpu().numpy(), rtol=1e-13)
    bt = la.Vector(im, bs[0])
    ut = la.Vector(im, bs[1])
    bt.array.copy_(torch.arange(bt.array.numel(), dtype=torch.float64))
    A.mult(bt, ut, transpose=True)
    assert np.allclose(ut.array.cpu().numpy(), As.T @ bt.array.cpu().numpy(), rtol=1e-13)
    assert np.allclose(A.to_dense(), As.toarray())


@pytest.mark.parametrize("bs", [(1, 1), (2, 2), (3, 3), (2, 3), (3, 2)])
def test_transpose(bfx, oracle, bs):
    """MatrixCSR.transpose (la::transpose, la/mattrans.h) on one rank: structure and values bit-identical to the
    oracle's restatement of impl::local_transpose, and equal to the scipy transpose - the check of
    python/test/unit/la/test_transpose.py:21-68 - for a random rectangular matrix with empty rows and columns."""
    la, common, torch, lib = bfx.la, bfx.common, bfx.torch, bfx.lib
    n0, n1 = 37, 23
    im0, im1 = common.IndexMap(common.COMM_SELF, n0), common.IndexMap(common.COMM_SELF, n1)
    rng = np.random.default_rng(12345)
    sp = la.SparsityPattern(common.COMM_SELF, [im0, im1], bs)
    for i in range(n0):
        if i % 7 == 3:
            continue  # empty row
        cols = rng.choice(n1 - 1, size=rng.integers(1, 9), replace=False)  # (column n1 - 1 stays empty)
        sp.insert(np.array([i]), np.sort(cols))
    sp.finalize()
    A = la.MatrixCSR(sp)
    A.data.copy_(torch.from_numpy(rng.random(A.data.numel())))
    AT = A.transpose()
    assert AT.block_size == [bs[1], bs[0]]
    assert AT.index_map(0).size_local == n1 and AT.index_map(1).size_local == n0
    oA = oracle.OMatrix([oracle.make_index_maps([n0], [[]], [[]])[0], oracle.make_index_maps([n1], [[]], [[]])[0]], bs,
                        A.data.cpu().numpy(), A.indices, A.indptr, A.off_diag_offset)
    c0, rp0, v0 = oracle.local_transpose(oA)
    assert np.array_equal(AT.indptr, rp0) and np.array_equal(AT.indices, c0)
    assert np.array_equal(AT.data.cpu().numpy(), v0)
    assert np.array_equal(AT.to_scipy().toarray(), A.to_scipy().toarray().T)
    # (A^T)^T == A
    ATT = AT.transpose()
    assert np.array_equal(ATT.indptr, A.indptr) and np.array_equal(ATT.indices, A.indices)
    assert np.array_equal(ATT.data.cpu().numpy(), A.data.cpu().numpy())
    # mult with the transposed matrix == multT with the original
    x = la.Vector(im0, bs[0])
    x.array.copy_(torch.arange(x.array.numel(), dtype=torch.float64))
    y1, y2 = la.Vector(im1, bs[1]), la.Vector(im1, bs[1])
    AT.mult(x, y1)
    A.mult(x, y2, transpose=True)
    assert np.allclose(y1.array.cpu().numpy(), y2.array.cpu().numpy(), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("kind", ["p1", "p2", "q1"])
def test_spmv_assembled(bfx, oracle, kind):
    la, fem, torch = bfx.la, bfx.fem, bfx.torch
    if kind == "p1":
        p, k, c = P.tet_p1(12, numbering="random", seed=9), bfx.lib.K_POISSON_P1_TET_A, [2.0]
    elif kind == "p2":
        p, k, c = P.tet_p2(6), bfx.lib.K_POISSON_P2_TET_A, [2.0]
    else:
        p, k, c = P.hex_q1(6), bfx.lib.K_ELASTICITY_Q1_HEX_A, [[1.0, 2.0]]
    msh, V = make_space(bfx, p)
    a, sp, A = assemble_A(bfx, V, k, constants=c)
    x = la.Vector(A.index_map(1), p.bs)
    y = la.Vector(A.index_map(0), p.bs)
    xh = np.random.default_rng(12345).random(x.array.numel())
    x.array.copy_(torch.from_numpy(xh))
    y0 = np.random.default_rng(1).random(y.array.numel())
    y.array.copy_(torch.from_numpy(y0))
    A.mult(x, y)  # y += A x
    data = A.data.cpu().numpy()
    yref = y0.copy()
    oracle.spmv(data, A.indptr[:-1], A.indptr[1:], A.indices, xh, yref, p.bs, p.bs)
    assert np.max(np.abs(y.array.cpu().numpy() - yref)) <= TOL * np.max(np.abs(yref))
    if p.bs in (1, 3):
        # every kernel variant (bs = 1: stream, rows, TMA-pipelined rows; bs = 3: warp-staged, TMA-fed), not only
        # the one the timing picked
        for variant in ((0, 1, 2) if p.bs == 1 else (0, 1)):
            bfx.lib.check(bfx.lib.lib.bfx_csr_set_spmv_variant(A._csr, variant))
            y.array.copy_(torch.from_numpy(y0))
            A.mult(x, y)
            assert np.max(np.abs(y.array.cpu().numpy() - yref)) <= TOL * np.max(np.abs(yref)), variant


def test_insert_set_add_and_errors(bfx, oracle):
    """python/test/unit/la/test_matrix_csr.py:36-107, 253-269."""
    la, common = bfx.la, bfx.common
    im = common.IndexMap(common.COMM_SELF, 8)
    for bs in (1, 2):
        sp = la.SparsityPattern(common.COMM_SELF, [im, im], [bs, bs])
        sp.insert(np.arange(4), np.arange(4))
        sp.finalize()
        A = la.MatrixCSR(sp)
        ref = np.zeros(A.data.numel())
        edges, offsets = sp.graph
        for dbs, kind in ((bs, "csr"), (1 if bs == 2 else 2, "nonblocked" if bs == 2 else "blocked")):
            nr, nc = 2, 2
            x = np.random.default_rng(dbs).random(nr * nc * dbs * dbs)
            if kind == "nonblocked":
                rows, cols = [0, 3], [1, 2]
            else:
                rows, cols = [0, 1], [1, 0]
            A.add(x, rows, cols, dbs)
            oracle.insert_csr(kind, ref, edges, offsets, x, rows, cols, dbs if kind != "nonblocked" else bs,
                              dbs if kind != "nonblocked" else bs, "add")
            A.set(x, rows, cols, dbs)
            oracle.insert_csr(kind, ref, edges, offsets, x, rows, cols, dbs if kind != "nonblocked" else bs,
                              dbs if kind != "nonblocked" else bs, "set")
        assert np.allclose(A.data.cpu().numpy(), ref, rtol=1e-15)
        with pytest.raises(RuntimeError, match="Entry not in sparsity"):
            A.add(np.ones(bs * bs), [0], [6], bs)


def test_not_in_sparsity_at_plan_build(bfx):
    """Assembling a form whose dofmap touches entries outside the pattern raises like insert_csr."""
    fem, la, common = bfx.fem, bfx.la, bfx.common
    p = P.tet_p1(2)
    msh, V = make_space(bfx, p)
    a = fem.Form([V, V], {fem.IntegralType.cell: [(0, bfx.lib.K_POISSON_P1_TET_A, None, [])]}, constants=[fem.Constant(1.0)])
    im = V.dofmap.index_map
    sp = la.SparsityPattern(common.COMM_SELF, [im, im], [1, 1])
    sp.insert_diagonal(np.arange(p.ndofs))
    sp.finalize()
    A = la.MatrixCSR(sp)
    with pytest.raises(RuntimeError, match="Entry not in sparsity"):
        fem.assemble_matrix(A, a)


def test_vector_reductions(bfx):
    """cpp/test/vector.cpp:22-55 flavour on one rank: norms and inner product."""
    la, common, torch = bfx.la, bfx.common, bfx.torch
    im = common.IndexMap(common.COMM_SELF, 100003)
    v = la.Vector(im, 1)
    h = np.random.default_rng(3).standard_normal(100003)
    v.array.copy_(torch.from_numpy(h))
    assert la.norm(v, la.Norm.l2) == pytest.approx(np.linalg.norm(h), rel=1e-13)
    assert la.norm(v, la.Norm.l1) == pytest.approx(np.abs(h).sum(), rel=1e-13)
    assert la.norm(v, la.Norm.linf) == np.abs(h).max()
    w = la.Vector(im, 1)
    w.set(2.0)
    assert la.inner_product(v, w) == pytest.approx(2 * h.sum(), rel=1e-11, abs=1e-9)


def test_host_buffer_entry(bfx, oracle):
    """bfx_assemble_matrix_cells_host: the drop-in call with the reference's host containers."""
    import ctypes as C

    fem, la, lib = bfx.fem, bfx.la, bfx.lib
    p = P.tet_p1(6, numbering="first_touch")
    msh, V = make_space(bfx, p)
    a, sp, A = assemble_A(bfx, V, lib.K_POISSON_P1_TET_A, constants=[2.0])
    plan = fem._asm_plan(a, a.integral(fem.IntegralType.cell, 0), fem.IntegralType.cell, A)
    out = np.zeros(A.data.numel())
    carr, nc = lib.constants_array([2.0])
    lib.check(lib.lib.bfx_assemble_matrix_cells_host(plan, lib.K_POISSON_P1_TET_A, p.x.ctypes.data, len(p.x), None, None, 0,
                                                     None, 0, 1, carr, nc, out.ctypes.data, lib.ASM_ATOMIC,
                                                     lib.current_stream()))
    pat, ref = P.oracle_assemble_matrix(oracle, p, oracle.K_POISSON_P1_TET_A, constants=np.array([2.0]))
    assert P.row_scaled_error(out, ref, pat.offsets) <= TOL


def test_host_buffer_entry_two_in_flight(bfx, oracle):
    """bfx_assemble_matrix_cells_host_begin/_end: two calls in flight on the plan's two slots, results in call
    order; a third begin and an end without a call in flight are refused."""
    fem, la, lib = bfx.fem, bfx.la, bfx.lib
    p = P.tet_p1(6, numbering="first_touch")
    msh, V = make_space(bfx, p)
    a, sp, A = assemble_A(bfx, V, lib.K_POISSON_P1_TET_A, constants=[2.0])
    plan = fem._asm_plan(a, a.integral(fem.IntegralType.cell, 0), fem.IntegralType.cell, A)
    outs = [np.zeros(A.data.numel()) for _ in range(3)]
    kappas = [2.0, 3.0, 5.0]

    def begin(k):
        carr, nc = lib.constants_array([kappas[k]])
        return lib.lib.bfx_assemble_matrix_cells_host_begin(plan, lib.K_POISSON_P1_TET_A, p.x.ctypes.data, len(p.x), None,
                                                            None, 0, None, 0, 1, carr, nc, outs[k].ctypes.data,
                                                            lib.ASM_CHUNKED)

    assert lib.lib.bfx_assemble_matrix_cells_host_end(plan) != 0
    lib.check(begin(0))
    lib.check(begin(1))
    assert begin(2) != 0  # both slots busy
    lib.check(lib.lib.bfx_assemble_matrix_cells_host_end(plan))  # call 0 complete
    lib.check(begin(2))
    lib.check(lib.lib.bfx_assemble_matrix_cells_host_end(plan))
    lib.check(lib.lib.bfx_assemble_matrix_cells_host_end(plan))
    assert lib.lib.bfx_assemble_matrix_cells_host_end(plan) != 0
    pat, ref = P.oracle_assemble_matrix(oracle, p, oracle.K_POISSON_P1_TET_A, constants=np.array([1.0]))
    for k in range(3):
        assert P.row_scaled_error(outs[k], kappas[k] * ref, pat.offsets) <= TOL


@pytest.mark.parametrize("degree", [1, 2])
def test_matrix_free_action_and_cg(bfx, oracle, degree):
    """SURVEY.md §8f rank 2, the flow of cpp/demo/poisson_matrix_free/main.cpp:150-247 on a tet box: the action
    kernel M = action(a, ui) against the oracle and against the assembled matrix times u; the lifting by
    assemble_vector(M) with ui = -g on the boundary; CG with the operator given by its action (all vectors on
    the device) reproducing u = 1 + x^2 + 2 y^2 + 3 z^2 (exactly representable in P2; P1: against a direct solve
    of the oracle-assembled system); the error functional through assemble_scalar."""
    fem, la, K, torch = bfx.fem, bfx.la, bfx.lib, bfx.torch
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    if degree == 1:
        p = P.tet_p1(6, numbering="random", seed=2)
        kA, okA, kM, okM, kL, okL = (K.K_POISSON_P1_TET_A, oracle.K_POISSON_P1_TET_A, K.K_ACTION_POISSON_P1_TET_L,
                                     oracle.K_ACTION_POISSON_P1_TET_L, K.K_LOAD_P1_TET_L, oracle.K_LOAD_P1_TET_L)
    else:
        p = P.tet_p2(4)
        kA, okA, kM, okM, kL, okL = (K.K_POISSON_P2_TET_A, oracle.K_POISSON_P2_TET_A, K.K_ACTION_POISSON_P2_TET_L,
                                     oracle.K_ACTION_POISSON_P2_TET_L, K.K_LOAD_P2_TET_L, oracle.K_LOAD_P2_TET_L)
    msh, V = make_space(bfx, p)
    dc = p.dof_coords
    uex = 1.0 + dc[:, 0] ** 2 + 2 * dc[:, 1] ** 2 + 3 * dc[:, 2] ** 2
    bdofs = np.flatnonzero(np.any((dc < 1e-12) | (dc > 1 - 1e-12), axis=1)).astype(np.int32)
    u_D = fem.Function(V)
    u_D.x.array.copy_(torch.from_numpy(uex))
    bc = fem.DirichletBC(u_D, bdofs)
    ui = fem.Function(V)
    kappa = fem.Constant(1.0)
    M = fem.Form([V], {fem.IntegralType.cell: [(0, kM, None, [0])]}, coefficients=[ui], constants=[kappa])
    f = fem.Function(V)
    f.x.array.fill_(-12.0)  # -laplace(u) = -12
    L = fem.Form([V], {fem.IntegralType.cell: [(0, kL, None, [0])]}, coefficients=[f])

    # --- action kernel: oracle, and the assembled matrix applied to the same vector
    w = np.random.default_rng(3).random(p.ndofs)
    ui.x.array.copy_(torch.from_numpy(w))
    y = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(y, M)
    yref = P.oracle_assemble_vector(oracle, p, okM, coeff=(w, p.dofmap, 1), constants=np.array([1.0]))
    assert np.max(np.abs(y.array.cpu().numpy() - yref)) <= TOL * np.max(np.abs(yref))
    pat, Aref = P.oracle_assemble_matrix(oracle, p, okA, constants=np.array([1.0]))
    Asp = sp.csr_matrix((Aref, pat.edges, pat.offsets), shape=(p.ndofs, p.ndofs))
    assert np.max(np.abs(yref - Asp @ w)) <= 1e-11 * np.max(np.abs(yref))

    # --- rhs with lifting through the action (main.cpp:185-200)
    b = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b, L)
    ui.x.array.zero_()
    bc.set(ui.x.array, None, -1.0)
    fem.assemble_vector(b, M)
    b.scatter_reverse(la.InsertMode.add)
    bc.set(b.array, None, 0.0)
    b.scatter_forward()
    bref = P.oracle_assemble_vector(oracle, p, okL, coeff=(np.full(p.ndofs, -12.0), p.dofmap, 1))
    g = np.zeros(p.ndofs)
    g[bdofs] = uex[bdofs]
    bref = bref - Asp @ g
    bref[bdofs] = 0.0
    assert np.max(np.abs(b.array.cpu().numpy() - bref)) <= 1e-11 * np.max(np.abs(bref))

    # --- operator action (main.cpp:207-226) and CG (:84-132), all on the device
    def action(xv, yv):
        yv.array.zero_()
        ui.x.array.copy_(xv.array)
        fem.assemble_vector(yv, M)
        bc.set(yv.array, None, 0.0)
        yv.scatter_reverse(la.InsertMode.add)
        yv.scatter_forward()

    u = fem.Function(V)
    its = la.cg(u.x, b, action, kmax=400, rtol=1e-10)
    assert 0 < its < 400
    bc.set(u.x.array, None, 1.0)
    uh = u.x.array.cpu().numpy()
    # direct solve of the same system from the oracle-assembled matrix
    free = np.setdiff1d(np.arange(p.ndofs), bdofs)
    uref = g.copy()
    uref[free] = spl.spsolve(Asp[free][:, free].tocsc(), bref[free])
    assert np.max(np.abs(uh - uref)) <= 1e-7 * np.max(np.abs(uref))
    if degree == 2:
        assert np.max(np.abs(uh - uex)) <= 1e-7  # the quadratic is in the P2 space

    # --- error functional E = (usol - uexact)^2 dx (main.cpp:232-247), P1 only (the functional kernel is P1)
    if degree == 1:
        d = fem.Function(V)
        diff = uh - uex
        d.x.array.copy_(torch.from_numpy(diff))
        E = fem.Form([], {fem.IntegralType.cell: [(0, K.K_L2NORM2_P1_TET_M, None, [0])]}, coefficients=[d], mesh=msh)
        err2 = fem.assemble_scalar(E)
        cells = np.arange(len(p.dofmap), dtype=np.int32)
        coeffs = np.zeros((len(cells), 4))
        oracle.pack_coefficient(coeffs, 0, diff, p.dofmap, 1, cells=cells)
        ref2 = oracle.assemble_scalar(oracle.K_L2NORM2_P1_TET_M, p.x_dofmap, p.x, cells, coeffs=coeffs)
        assert err2 == pytest.approx(ref2, rel=1e-11) and err2 > 0


@pytest.mark.parametrize("grouped", [True, False])
def test_vector_assembly_grouped_and_atomic(bfx, oracle, grouped, monkeypatch):
    """assemble_vector of the P1 load (fused coefficient gather) with the grouped kernel (one RED per distinct
    dof of 32 Morton-ordered cells) and with the cell-parallel kernel, on a mesh of many groups with a random
    numbering, a tail group, and a cell subset; accumulation into a non-zero vector."""
    fem, la, K, torch = bfx.fem, bfx.la, bfx.lib, bfx.torch
    monkeypatch.setattr(fem, "GROUPED_VECTORS", grouped)
    p = P.tet_p1(9, numbering="random", seed=4)
    msh, V = make_space(bfx, p)
    fh = P.source_f(p.dof_coords) + 0.25
    f = fem.Function(V)
    f.x.array.copy_(torch.from_numpy(fh))
    L = fem.Form([V], {fem.IntegralType.cell: [(0, K.K_LOAD_P1_TET_L, None, [0])]}, coefficients=[f])
    b = la.Vector(V.dofmap.index_map, 1)
    b.array.fill_(0.5)
    fem.assemble_vector(b, L)
    bref = P.oracle_assemble_vector(oracle, p, oracle.K_LOAD_P1_TET_L, coeff=(fh, p.dofmap, 1), b=np.full(p.ndofs, 0.5))
    assert np.max(np.abs(b.array.cpu().numpy() - bref)) <= TOL * np.max(np.abs(bref))
    # a subset of the cells (every third cell)
    sub = np.arange(0, len(p.dofmap), 3, dtype=np.int32)
    Ls = fem.Form([V], {fem.IntegralType.cell: [(0, K.K_LOAD_P1_TET_L, sub, [0])]}, coefficients=[f])
    bs_ = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(bs_, Ls)
    coeffs = np.zeros((len(sub), 4))
    oracle.pack_coefficient(coeffs, 0, fh, p.dofmap, 1, cells=sub)
    bsref = np.zeros(p.ndofs)
    oracle.assemble_vector(oracle.K_LOAD_P1_TET_L, p.x_dofmap, p.x, sub, p.dofmap, 1, bsref, coeffs=coeffs)
    assert np.max(np.abs(bs_.array.cpu().numpy() - bsref)) <= TOL * np.max(np.abs(bsref))


def test_interior_facets_dS(bfx, oracle):
    """a = inner(avg(u), avg(v))*dS on the unit square 12 x 12, P1 — the reference's golden
    norm 2.1834054713561906 (test_ghost_mesh_assembly.py:104-122) from the CUDA path, pattern and values against
    the oracle, with and without Dirichlet rows/columns."""
    fem, la, K = bfx.fem, bfx.la, bfx.lib
    from dolfinx_b200 import mesh as M

    p = P.tri_p1(12, 12)
    msh, V = make_space(bfx, p)
    facets = M.interior_facets(p.x_dofmap, M.TRI_FACETS)
    a = fem.Form([V, V], {fem.IntegralType.interior_facet: [(0, K.K_AVG_MASS_P1_TRI_DS, facets, [])]})
    sp = fem.create_sparsity_pattern(a)
    sp.finalize()
    A = la.MatrixCSR(sp)
    fem.assemble_matrix(A, a)
    A.scatter_reverse()
    assert np.sqrt(A.squared_norm()) == pytest.approx(2.1834054713561906, rel=1e-12)
    maps = oracle.make_index_maps([p.ndofs], [[]], [[]])
    r, c = oracle.sparsity_insert_interior_facets(facets, p.dofmap, p.dofmap)
    pat = oracle.sparsity_finalize(maps, maps, (1, 1), [r], [c])[0]
    ref = np.zeros(len(pat.edges))
    oracle.assemble_matrix_interior_facets(oracle.K_AVG_MASS_P1_TRI_DS, p.x_dofmap, p.x, facets, p.dofmap, 1, p.dofmap, 1,
                                           ref, pat.edges, pat.offsets)
    check_matrix(A, pat, ref)
    # Dirichlet rows / columns zeroed in the macro element (fem/assemble_matrix_impl.h:613-645)
    bdofs = np.flatnonzero((p.dof_coords[:, 0] < 1e-12) | (p.dof_coords[:, 1] > 1 - 1e-12)).astype(np.int32)
    bc = fem.DirichletBC(fem.Constant(0.0), bdofs, V)
    A2 = la.MatrixCSR(sp)
    fem.assemble_matrix(A2, a, bcs=[bc])
    mk = np.zeros(p.ndofs, dtype=np.int8)
    mk[bdofs] = 1
    ref2 = np.zeros(len(pat.edges))
    oracle.assemble_matrix_interior_facets(oracle.K_AVG_MASS_P1_TRI_DS, p.x_dofmap, p.x, facets, p.dofmap, 1, p.dofmap, 1,
                                           ref2, pat.edges, pat.offsets, bc0=mk, bc1=mk)
    check_matrix(A2, pat, ref2)


def test_block_mode_expanded(bfx, oracle):
    """MatrixCSR(pattern, BlockMode.expanded) (la/MatrixCSR.h:638-694; python/test/unit/la/test_matrix_csr.py:70-107):
    the structure against a literal restatement of the reference loop, blocked data added to the compact and the
    expanded matrix give the same dense matrix, and both multiply alike."""
    la, common, torch = bfx.la, bfx.common, bfx.torch
    n, bs0, bs1 = 9, 2, 3
    im = common.IndexMap(common.COMM_SELF, n)
    sp = la.SparsityPattern(common.COMM_SELF, [im, im], [bs0, bs1])
    rng = np.random.default_rng(7)
    for r in range(n):
        sp.insert([r], np.unique(rng.integers(0, n, size=4)))
    sp.insert_diagonal(np.arange(n))
    sp.finalize()
    A = la.MatrixCSR(sp)
    B = la.MatrixCSR(sp, la.BlockMode.expanded)
    assert B.bs == (1, 1) and B.index_map(0).size_local == n * bs0 and B.index_map(1).size_local == n * bs1
    edges, offsets = sp.graph
    cols_ref, ptr_ref = [], [0]
    for i in range(n):  # the reference loop, :671-687
        for q0 in range(bs0):
            for j in range(offsets[i], offsets[i + 1]):
                for q1 in range(bs1):
                    cols_ref.append(edges[j] * bs1 + q1)
            ptr_ref.append(len(cols_ref))
    assert np.array_equal(B.indices, np.array(cols_ref, dtype=np.int32)) and np.array_equal(B.indptr, np.array(ptr_ref))
    assert np.array_equal(B.off_diag_offset, np.array(ptr_ref[1:]))  # one rank: every column is owned
    # the same dense matrix stored in both: block values into the compact matrix, scalars into the expanded one
    dense = rng.random((n * bs0, n * bs1))
    mask = np.zeros_like(dense, dtype=bool)
    for r in range(n):
        for c in edges[offsets[r]:offsets[r + 1]]:
            mask[r * bs0:(r + 1) * bs0, c * bs1:(c + 1) * bs1] = True
    dense[~mask] = 0.0
    A.set_value(0.0)
    vals = np.zeros(A.data.numel())
    for r in range(n):
        for k, c in enumerate(edges[offsets[r]:offsets[r + 1]]):
            vals[(offsets[r] + k) * bs0 * bs1:(offsets[r] + k + 1) * bs0 * bs1] = dense[r * bs0:(r + 1) * bs0, c * bs1:(c + 1) * bs1].reshape(-1)
    A.data.copy_(torch.from_numpy(vals))
    for R in range(n * bs0):  # scalar entries into the expanded matrix (insert_csr with bs = 1)
        cols = B.indices[B.indptr[R]:B.indptr[R + 1]]
        B.add(dense[R, cols], [R], cols, 1)
    assert np.array_equal(A.to_dense(), dense) and np.array_equal(B.to_dense(), dense)
    x = la.Vector(A.index_map(1), bs1)
    xe = la.Vector(B.index_map(1), 1)
    xh = rng.random(n * bs1)
    x.array.copy_(torch.from_numpy(xh))
    xe.array.copy_(torch.from_numpy(xh))
    y = la.Vector(A.index_map(0), bs0)
    ye = la.Vector(B.index_map(0), 1)
    A.mult(x, y)
    B.mult(xe, ye)
    assert np.allclose(y.array.cpu().numpy(), dense @ xh, rtol=1e-14) and np.allclose(ye.array.cpu().numpy(), dense @ xh, rtol=1e-14)


def test_facet_functionals_and_interior_facet_vector(bfx, oracle):
    """Row f3 of SURVEY.md §8: interior-facet linear forms (fem/assemble_vector_impl.h:249-339) and the facet branches of
    assemble_scalar (fem/assemble_scalar_impl.h:78-168) on the CUDA path: the reference's expectations
    (test_assemble_domains.py:203-240) and the oracle, fused and packed coefficients, additivity of a form with
    cell + exterior + interior integrals."""
    fem, la, K, torch = bfx.fem, bfx.la, bfx.lib, bfx.torch
    from dolfinx_b200 import mesh as M

    N = 10
    p = P.tri_p1(N, N)
    msh, V = make_space(bfx, p)
    facets = M.interior_facets(p.x_dofmap, M.TRI_FACETS)
    ext = M.exterior_facets(p.x_dofmap, M.TRI_FACETS)
    length = 2 * (N - 1) + N * np.sqrt(2.0)
    one = fem.Form([], {fem.IntegralType.interior_facet: [(0, K.K_ONE_TRI_DS_M, facets, [])]}, mesh=msh)
    assert fem.assemble_scalar(one) == pytest.approx(length, rel=1e-13)

    f = fem.Function(V)
    fv = p.dof_coords[:, 0] + 2 * p.dof_coords[:, 1] ** 2
    f.x.array.copy_(torch.from_numpy(fv).to(f.x.array.device))
    j3 = fem.Form([], {fem.IntegralType.interior_facet: [(0, K.K_AVG2_COEFF_P1_TRI_DS_M, facets, [0])]}, coefficients=[f], mesh=msh)
    w3 = np.concatenate([fv[p.dofmap[facets[:, 0, 0]]], fv[p.dofmap[facets[:, 1, 0]]]], axis=1)
    ref3 = oracle.assemble_scalar_interior_facets(oracle.K_AVG2_COEFF_P1_TRI_DS_M, p.x_dofmap, p.x, facets, coeffs=w3)
    got3 = fem.assemble_scalar(j3)
    assert got3 == pytest.approx(ref3, rel=1e-12)
    assert fem.assemble_scalar(j3, coeffs=fem.pack_coefficients(j3)) == pytest.approx(ref3, rel=1e-12)  # packed layout
    j2 = fem.Form([], {fem.IntegralType.exterior_facet: [(0, K.K_COEFF2_P1_TRI_FACET_M, ext, [0])]}, coefficients=[f], mesh=msh)
    ref2 = oracle.assemble_scalar_facets(oracle.K_COEFF2_P1_TRI_FACET_M, p.x_dofmap, p.x, ext, coeffs=fv[p.dofmap[ext[:, 0]]])
    got2 = fem.assemble_scalar(j2)
    assert got2 == pytest.approx(ref2, rel=1e-12)
    # additivity (test_assemble_domains.py:213-240): one form with both facet integrals
    j23 = fem.Form([], {fem.IntegralType.exterior_facet: [(0, K.K_COEFF2_P1_TRI_FACET_M, ext, [0])],
                        fem.IntegralType.interior_facet: [(0, K.K_AVG2_COEFF_P1_TRI_DS_M, facets, [0])]},
                   coefficients=[f], mesh=msh)
    assert fem.assemble_scalar(j23) == pytest.approx(got2 + got3, rel=1e-13)
    # f = 2 / 3: the reference's constants
    f.x.array.fill_(3.0)
    assert fem.assemble_scalar(j3) == pytest.approx(9.0 * length, rel=1e-13)
    f.x.array.fill_(2.0)
    assert fem.assemble_scalar(j2) == pytest.approx(16.0, rel=1e-13)

    # linear form conj(avg(v))*dS (test_assembler.py:1003)
    L = fem.Form([V], {fem.IntegralType.interior_facet: [(0, K.K_AVG_LOAD_P1_TRI_DS_L, facets, [])]})
    b = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b, L)
    ref = np.zeros(p.ndofs)
    oracle.assemble_vector_interior_facets(oracle.K_AVG_LOAD_P1_TRI_DS_L, p.x_dofmap, p.x, facets, p.dofmap, 1, ref)
    got = b.array.cpu().numpy()
    assert np.max(np.abs(got - ref)) <= TOL * np.max(np.abs(ref))
    assert got.sum() == pytest.approx(length, rel=1e-13)
    fem.assemble_vector(b, L)  # accumulates (fem/assembler.h:230-257)
    assert np.max(np.abs(b.array.cpu().numpy() - 2 * ref)) <= 2 * TOL * np.max(np.abs(ref))


def test_two_fused_coefficients(bfx, oracle):
    """An integral with TWO active coefficients (L = f g v dx): the fused gather takes both from their dof vectors, the
    packed path reads the reference layout [f | g] at the form's coefficient offsets; both equal the oracle."""
    fem, la, K, torch = bfx.fem, bfx.la, bfx.lib, bfx.torch
    p = P.tet_p1(5, numbering="random")
    msh, V = make_space(bfx, p)
    dc = p.dof_coords
    fv, gv = 1 + dc[:, 0] + np.sin(3 * dc[:, 1]), 2 - dc[:, 1] + dc[:, 2] ** 2
    f, g = fem.Function(V), fem.Function(V)
    f.x.array.copy_(torch.from_numpy(fv).to(f.x.array.device))
    g.x.array.copy_(torch.from_numpy(gv).to(g.x.array.device))
    L = fem.Form([V], {fem.IntegralType.cell: [(0, K.K_LOAD_PROD_P1_TET_L, None, [0, 1])]}, coefficients=[f, g])
    cells = np.arange(len(p.dofmap), dtype=np.int32)
    ref = np.zeros(p.ndofs)
    oracle.assemble_vector(oracle.K_LOAD_PROD_P1_TET_L, p.x_dofmap, p.x, cells, p.dofmap, 1, ref,
                           coeffs=np.concatenate([fv[p.dofmap], gv[p.dofmap]], axis=1))
    b = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b, L)
    assert np.max(np.abs(b.array.cpu().numpy() - ref)) <= TOL * np.max(np.abs(ref))
    b2 = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b2, L, coeffs=fem.pack_coefficients(L))
    assert np.max(np.abs(b2.array.cpu().numpy() - ref)) <= TOL * np.max(np.abs(ref))
    # wrong number of coefficients for the kernel: refused by the library, not silently mis-read
    L1 = fem.Form([V], {fem.IntegralType.cell: [(0, K.K_LOAD_PROD_P1_TET_L, None, [0])]}, coefficients=[f, g])
    with pytest.raises(K.BfxError):
        fem.assemble_vector(la.Vector(V.dofmap.index_map, 1), L1)


def test_empty_inputs_on_gpu(bfx, oracle):
    """Edge cases through the C-ABI (reference behaviour: an empty cell loop assembles nothing,
    fem/assemble_matrix_impl.h:127; an empty bc list marks nothing, fem/assembler.h:558-577; set_diagonal over no
    owned bc rows returns, :644-686): zero cells, zero facets, empty bc lists, zero-length entity lists of every
    integral type, a vector of length zero."""
    fem, la, K, torch, common = bfx.fem, bfx.la, bfx.lib, bfx.torch, bfx.common
    from dolfinx_b200 import mesh as M

    p = P.tet_p1(3)
    msh, V = make_space(bfx, p)
    none = np.zeros(0, dtype=np.int32)
    a_all = fem.Form([V, V], {fem.IntegralType.cell: [(0, K.K_POISSON_P1_TET_A, None, [])]}, constants=[fem.Constant(2.0)])
    sp = fem.create_sparsity_pattern(a_all)
    sp.finalize()
    # (1) a cell integral over NO cells on the full pattern: A stays zero, every strategy
    for strat in (None, K.ASM_ATOMIC):
        a0 = fem.Form([V, V], {fem.IntegralType.cell: [(0, K.K_POISSON_P1_TET_A, none, [])]}, constants=[fem.Constant(2.0)])
        A = la.MatrixCSR(sp)
        fem.assemble_matrix(A, a0, strategy=strat)
        A.scatter_reverse()
        assert A.squared_norm() == 0.0
    # (2) empty bc list and a bc without dofs: same matrix as no bc at all
    A1, A2, A3 = la.MatrixCSR(sp), la.MatrixCSR(sp), la.MatrixCSR(sp)
    fem.assemble_matrix(A1, a_all)
    fem.assemble_matrix(A2, a_all, bcs=[])
    bc_empty = fem.DirichletBC(fem.Constant(0.0), none, V)
    fem.assemble_matrix(A3, a_all, bcs=[bc_empty])
    fem.set_diagonal(A3, V, [bc_empty], 1.0)
    assert torch.equal(A1.data, A2.data) and torch.equal(A1.data, A3.data)
    b = la.Vector(V.dofmap.index_map, 1)
    fem.apply_lifting(b, [a_all], [[bc_empty]])
    fem.apply_lifting(b, [a_all], [[]])
    fem.set_bc(b, [bc_empty])
    assert float(b.array.abs().max()) == 0.0
    # (3) linear form and functional over no cells / no facets
    f = fem.Function(V)
    f.x.array.fill_(1.0)
    L0 = fem.Form([V], {fem.IntegralType.cell: [(0, K.K_LOAD_P1_TET_L, none, [0])],
                        fem.IntegralType.exterior_facet: [(0, K.K_FACET_LOAD_P1_TET_L, np.zeros((0, 2), dtype=np.int32), [0])]},
                  coefficients=[f])
    fem.assemble_vector(b, L0)
    assert float(b.array.abs().max()) == 0.0
    M0 = fem.Form([], {fem.IntegralType.cell: [(0, K.K_L2NORM2_P1_TET_M, none, [0])]}, coefficients=[f], mesh=msh)
    assert fem.assemble_scalar(M0) == 0.0
    pt = P.tri_p1(3, 3)
    msht, Vt = make_space(bfx, pt)
    nofacets = np.zeros((0, 2, 2), dtype=np.int32)
    one = fem.Form([], {fem.IntegralType.interior_facet: [(0, K.K_ONE_TRI_DS_M, nofacets, [])]}, mesh=msht)
    assert fem.assemble_scalar(one) == 0.0
    Ls = fem.Form([Vt], {fem.IntegralType.interior_facet: [(0, K.K_AVG_LOAD_P1_TRI_DS_L, nofacets, [])]})
    bt = la.Vector(Vt.dofmap.index_map, 1)
    fem.assemble_vector(bt, Ls)
    assert float(bt.array.abs().max()) == 0.0
    # (4) zero-length vectors and an empty matrix (a rank that owns nothing): reductions, mult, scatter are no-ops
    im0 = common.IndexMap(common.COMM_SELF, 0)
    z = la.Vector(im0, 1)
    assert la.norm(z) == 0.0 and la.inner_product(z, z) == 0.0
    z.scatter_forward()
    z.scatter_reverse()
    sp0 = la.SparsityPattern(common.COMM_SELF, [im0, im0], [1, 1])
    sp0.finalize()
    A0 = la.MatrixCSR(sp0)
    assert A0.squared_norm() == 0.0 and A0.data.numel() == 0
    A0.mult(z, z)
    A0.scatter_reverse()


@pytest.mark.parametrize("perturb", [0.0, 0.3])
def test_q1_rowgather_row_ranges(bfx, oracle, perturb):
    """bfx_assemble_matrix_rows: the row-gather kernel on two disjoint row ranges (tail first, as the distributed
    overlap does with its ghost rows) reproduces the one-call result bit for bit on affine cells, and to 1e-12 with
    the non-affine remainder (REDs); rows outside a call's range are not touched; badly cut ranges are refused."""
    import ctypes as C

    fem, la, K, torch = bfx.fem, bfx.la, bfx.lib, bfx.torch
    n = 6
    p = P.hex_q1(n, numbering="random", seed=3, perturb=perturb)
    if perturb:
        keep = p.x[:, 2] < 0.5
        p.x[keep] = P.hex_q1(n, numbering="random", seed=3).x[keep]
    msh, V = make_space(bfx, p)
    bdofs = np.flatnonzero(p.dof_coords[:, 0] < 1e-12).astype(np.int32)
    bc = fem.DirichletBC(fem.Constant(np.zeros(3)), bdofs, V)
    a, sp, A = assemble_A(bfx, V, K.K_ELASTICITY_Q1_HEX_A, constants=[[1.0, 1.5]], bcs=[bc], strategy=K.ASM_ROWGATHER)
    whole = A.data.clone()
    integ = a.integral(fem.IntegralType.cell, 0)
    plan = fem._asm_plan(a, integ, fem.IntegralType.cell, A)
    tr = C.c_int(0)
    K.check(K.lib.bfx_asm_rowgather_tile_rows(plan, C.byref(tr)))
    assert tr.value > 0
    n_all = A.num_all_rows()
    split = (n_all * 2 // 3 // tr.value) * tr.value
    mk = fem._bc_markers(V, [bc])
    carr, nc = K.constants_array(fem.pack_constants(a))
    cf = K.make_coeffs()
    vals = torch.full_like(whole, 7.0)  # (sentinel: rows outside the range must keep it)

    def rows(r0, r1, reuse):
        return K.lib.bfx_assemble_matrix_rows(plan, integ.kernel, msh.x.data_ptr(), mk.data_ptr(), mk.data_ptr(), C.byref(cf),
                                              carr, nc, vals.data_ptr(), r0, r1, reuse, K.current_stream())

    K.check(rows(split, n_all, 0))
    cut = int(A.indptr[split]) * 9
    assert bool((vals[:cut] == 7.0).all()), "rows before the range were written"
    K.check(rows(0, split, 1))
    if perturb:
        assert float((vals - whole).abs().max()) <= TOL * float(whole.abs().max())
    else:
        assert torch.equal(vals, whole)
    assert rows(1, n_all, 0) != K.OK  # not cut at a tile boundary


@pytest.mark.parametrize("case", ["p1", "p2"])
def test_chunk_plan_partition(bfx, oracle, case):
    """bfx_asm_chunk_partition / bfx_assemble_matrix_cells_part: the chunks of ONE plan (p1: the lean plan, whose flagged
    chunks move to the front; p2: the classic plan, chunk list + skip flags) split at a row threshold
    (the distributed overlap's ghost rows) - part 1 then part 2 give the one-launch matrix (overwrite mode on zeros
    and add mode), part 1 alone writes nothing into chunks without a row beyond the threshold, and a threshold beyond
    the last row leaves part 1 empty."""
    import ctypes as C

    fem, la, K, torch = bfx.fem, bfx.la, bfx.lib, bfx.torch
    if case == "p1":
        p = P.tet_p1(16, numbering="first_touch")  # (2^k cubes per edge: whole-cube chunks, complete warp tables)
        kid, okid = K.K_POISSON_P1_TET_A, oracle.K_POISSON_P1_TET_A
    else:
        p = P.tet_p2(7)
        kid, okid = K.K_POISSON_P2_TET_A, oracle.K_POISSON_P2_TET_A
    msh, V = make_space(bfx, p)
    bdofs = np.flatnonzero(p.dof_coords[:, 0] < 1e-12).astype(np.int32)
    bc = fem.DirichletBC(fem.Constant(0.0), bdofs, V)
    a, sp, A = assemble_A(bfx, V, kid, constants=[2.0], bcs=[bc])
    whole = A.data.clone()
    pat, ref = P.oracle_assemble_matrix(oracle, p, okid, constants=np.array([2.0]),
                                        bc=(np.isin(np.arange(p.ndofs), bdofs)).astype(np.int8))
    check_matrix(A, pat, ref)
    integ = a.integral(fem.IntegralType.cell, 0)
    plan = fem._asm_plan(a, integ, fem.IntegralType.cell, A)
    mk = fem._bc_markers(V, [bc])
    carr, nc = K.constants_array(fem.pack_constants(a))
    cf = K.make_coeffs()
    n1 = C.c_int64(-1)
    thr = (2 * p.ndofs) // 3
    K.check(K.lib.bfx_asm_chunk_partition(plan, thr, C.byref(n1)))
    nch = fem.chunk_stats(a, A)[0]
    assert 0 < n1.value < nch

    def part(vals, k, mode):
        K.check(K.lib.bfx_assemble_matrix_cells_part(plan, integ.kernel, msh.x.data_ptr(), mk.data_ptr(), mk.data_ptr(),
                                                     C.byref(cf), carr, nc, vals.data_ptr(), mode, k, K.current_stream()))

    vals = torch.zeros_like(whole)
    part(vals, 1, K.VALUES_OVERWRITE)
    first = vals.clone()
    assert 0 < int((first != 0).sum()) < int((whole != 0).sum())
    part(vals, 2, K.VALUES_OVERWRITE)
    assert float((vals - whole).abs().max()) <= TOL * float(whole.abs().max())
    part(vals, 1, K.VALUES_ADD)
    part(vals, 2, K.VALUES_ADD)
    assert float((vals - 2 * whole).abs().max()) <= 2 * TOL * float(whole.abs().max())
    K.check(K.lib.bfx_asm_chunk_partition(plan, p.ndofs, C.byref(n1)))
    assert n1.value == 0
    vals.zero_()
    part(vals, 1, K.VALUES_OVERWRITE)
    assert float(vals.abs().max()) == 0.0
    part(vals, 2, K.VALUES_OVERWRITE)
    assert float((vals - whole).abs().max()) <= TOL * float(whole.abs().max())
