"""Pin the CPU oracle: reference golden scalars, known answers, and the reference's own
la/matrix_csr_impl.h (compiled from /root/reference when present, oracle/_ref)."""

import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp

from tests import problems as P

REF_SO = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libref_csr.so")


def test_golden_p1_laplace_unit_square_13(oracle):
    """python/test/unit/fem/test_custom_jit_kernels.py:115-116 (and :290-294, frozen FFCx C)."""
    O = oracle
    p = P.tri_p1(13, 13)
    pat, data = P.oracle_assemble_matrix(O, p, O.K_LAPLACE_P1_TRI_A)
    b = P.oracle_assemble_vector(O, p, O.K_SOURCE_P1_TRI_L)
    assert np.isclose(np.sqrt(np.sum(data**2)), 56.124860801609124, rtol=1e-14, atol=0)
    assert np.isclose(np.linalg.norm(b), 0.0739710713711999, rtol=1e-14, atol=0)


def test_golden_ghost_mesh_assembly(oracle):
    """python/test/unit/fem/test_ghost_mesh_assembly.py:43-66: a = f u v dx + u v ds, L = f v dx + 2 v ds."""
    O = oracle
    from dolfinx_b200 import mesh as M

    p = P.tri_p1(12, 12)
    f = np.full(p.ndofs, 10.0)
    pat, data = P.oracle_assemble_matrix(O, p, O.K_MASS_COEFF_P1_TRI_A, coeff=(f, p.dofmap, 1))
    ents = M.exterior_facets(p.x_dofmap, M.TRI_FACETS)
    assert len(ents) == 48
    O.assemble_matrix(O.K_FACET_MASS_P1_TRI_A, p.x_dofmap, p.x, None, p.dofmap, 1, p.dofmap, 1, data, pat.edges,
                      pat.offsets, entities=ents)
    b = P.oracle_assemble_vector(O, p, O.K_LOAD_COEFF_P1_TRI_L, coeff=(f, p.dofmap, 1))
    O.assemble_vector(O.K_FACET_CONST_P1_TRI_L, p.x_dofmap, p.x, None, p.dofmap, 1, b, constants=np.array([2.0]),
                      entities=ents)
    assert np.sqrt(np.sum(data**2)) == pytest.approx(0.6713621455570528, rel=1e-12)
    assert np.linalg.norm(b) == pytest.approx(1.582294032953906, rel=1e-12)


def test_boundary_measure(oracle):
    """python/test/unit/fem/test_assembler.py:83-89 flavour: sum of int 1*v ds over the unit cube = 6."""
    O = oracle
    from dolfinx_b200 import mesh as M

    p = P.tet_p1(3)
    ents = M.exterior_facets(p.x_dofmap, M.TET_FACETS)
    assert len(ents) == 12 * 9
    b = np.zeros(p.ndofs)
    coeffs = np.ones((len(ents), 4))
    O.assemble_vector(O.K_FACET_LOAD_P1_TET_L, p.x_dofmap, p.x, None, p.dofmap, 1, b, coeffs=coeffs, entities=ents)
    assert b.sum() == pytest.approx(6.0, rel=1e-13)


def test_p2_poisson_box12_known_answers(oracle):
    """cpp/test/matrix.cpp:31-120: P2 Poisson kappa=2 on create_box 12^3 tets, A.1 = 0 to 1e-13;
    sizes from SURVEY.md Appendix A."""
    O = oracle
    p = P.tet_p2(12)
    assert p.ndofs == 15625
    pat, data = P.oracle_assemble_matrix(O, p, O.K_POISSON_P2_TET_A, constants=np.array([2.0]))
    assert len(pat.edges) == 417601
    y = np.zeros(p.ndofs)
    O.spmv(data, pat.offsets[:-1], pat.offsets[1:], pat.edges, np.ones(p.ndofs), y, 1, 1)
    assert np.max(np.abs(y)) < 1e-13
    A = sp.csr_matrix((data, pat.edges, pat.offsets))
    assert abs(A - A.T).max() < 1e-13


def test_p1_sizes_and_row_sums(oracle):
    O = oracle
    p = P.tet_p1(12, numbering="first_touch")
    pat, data = P.oracle_assemble_matrix(O, p, O.K_POISSON_P1_TET_A, constants=np.array([2.0]))
    assert p.ndofs == 2197 and len(pat.edges) == 29053
    A = sp.csr_matrix((data, pat.edges, pat.offsets))
    assert np.max(np.abs(A @ np.ones(p.ndofs))) < 1e-13
    # exact energy of u = x0: kappa * |grad u|^2 * volume = 2
    u = p.dof_coords[:, 0]
    assert u @ (A @ u) == pytest.approx(2.0, rel=1e-13)


def test_p1_load_integrates_exactly(oracle):
    O = oracle
    p = P.tet_p1(4, numbering="random", seed=3)
    f = 1.0 + 2.0 * p.dof_coords[:, 0] - p.dof_coords[:, 2]  # linear -> exact
    b = P.oracle_assemble_vector(O, p, O.K_LOAD_P1_TET_L, coeff=(f, p.dofmap, 1))
    assert b.sum() == pytest.approx(1.0 + 1.0 - 0.5, rel=1e-13)
    p2 = P.tet_p2(3)
    f2 = p2.dof_coords[:, 0] ** 2 + p2.dof_coords[:, 1]  # quadratic, in P2
    b2 = P.oracle_assemble_vector(O, p2, O.K_LOAD_P2_TET_L, coeff=(f2, p2.dofmap, 1))
    assert b2.sum() == pytest.approx(1.0 / 3.0 + 0.5, rel=1e-13)


def test_elasticity_rigid_body_modes(oracle):
    """Rigid-body null space (cf. python/test/unit/la/test_nullspace.py:86-127) + symmetry."""
    O = oracle
    p = P.hex_q1(3, skew=True)
    E, nu = 1.0e9, 0.3
    mu, lmbda = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    pat, data = P.oracle_assemble_matrix(O, p, O.K_ELASTICITY_Q1_HEX_A, constants=np.array([mu, lmbda]))
    A = sp.bsr_matrix((data.reshape(-1, 3, 3), pat.edges, pat.offsets)).tocsr()
    X = p.dof_coords
    modes = []
    for k in range(3):
        t = np.zeros_like(X)
        t[:, k] = 1
        modes.append(t.reshape(-1))
    for a, b in [(0, 1), (1, 2), (0, 2)]:
        r = np.zeros_like(X)
        r[:, a], r[:, b] = -X[:, b], X[:, a]
        modes.append(r.reshape(-1))
    scale = abs(A).max()
    for m in modes:
        assert np.max(np.abs(A @ m)) < 1e-12 * scale * np.max(np.abs(m))
    assert abs(A - A.T).max() < 1e-12 * scale


def test_reassembly_doubles(oracle):
    """python/test/unit/fem/test_assembler.py:145-165: assembling twice without zeroing doubles A and b."""
    O = oracle
    p = P.tet_p1(3)
    pat, data = P.oracle_assemble_matrix(O, p, O.K_POISSON_P1_TET_A, constants=np.array([1.0]))
    n1 = np.sum(data**2)
    P.oracle_assemble_matrix(O, p, O.K_POISSON_P1_TET_A, pat=pat, constants=np.array([1.0]), data=data)
    assert np.sum(data**2) == pytest.approx(4 * n1, rel=1e-14)


def test_lifting_identity(oracle):
    """python/test/unit/fem/test_assembler.py:276-294: b - A g (then set_bc) == assemble + apply_lifting + set_bc."""
    O = oracle
    p = P.tet_p1(4, numbering="random", seed=1)
    bdofs = np.flatnonzero(np.isclose(p.dof_coords[:, 0], 0.0) | np.isclose(p.dof_coords[:, 0], 1.0)).astype(np.int32)
    g = 1.0 + p.dof_coords[:, 1] * 3.0
    f = P.source_f(p.dof_coords)
    kappa = np.array([2.0])
    # unconstrained A, then b - A g
    pat, A0 = P.oracle_assemble_matrix(O, p, O.K_POISSON_P1_TET_A, constants=kappa)
    b0 = P.oracle_assemble_vector(O, p, O.K_LOAD_P1_TET_L, coeff=(f, p.dofmap, 1))
    gvec = np.zeros(p.ndofs)
    gvec[bdofs] = g[bdofs]
    ref = b0 - sp.csr_matrix((A0, pat.edges, pat.offsets)) @ gvec
    ref[bdofs] = g[bdofs]
    # lifting path
    b = P.oracle_assemble_vector(O, p, O.K_LOAD_P1_TET_L, coeff=(f, p.dofmap, 1))
    markers = np.zeros(p.ndofs, dtype=np.int8)
    values = np.zeros(p.ndofs)
    O.bc_mark(markers, bdofs)
    O.bc_set(values, bdofs, g, 0, 1)
    O.lift_bc(O.K_POISSON_P1_TET_A, p.x_dofmap, p.x, np.arange(len(p.dofmap)), p.dofmap, 1, p.dofmap, 1, b, values,
              markers, constants=kappa)
    O.bc_set(b, bdofs, g, 0, 1)
    assert np.allclose(b, ref, rtol=1e-12, atol=1e-14)
    # bc rows/cols zeroed + unit diagonal gives the same solution on the boundary rows
    pat, A1 = P.oracle_assemble_matrix(O, p, O.K_POISSON_P1_TET_A, constants=kappa, bc=markers)
    O.set_diagonal(A1, pat.edges, pat.offsets, 1, 1, bdofs, 1.0)
    A1m = sp.csr_matrix((A1, pat.edges, pat.offsets))
    assert np.allclose((A1m @ gvec)[bdofs], g[bdofs])
    assert abs(A1m - A1m.T).max() < 1e-14


@pytest.mark.parametrize("bs", [(1, 1), (2, 2), (1, 2), (2, 1), (2, 3), (3, 3)])
def test_spmv_vs_scipy(oracle, bs):
    """python/test/unit/la/test_matrix_vector.py:47-106 with the fixture data of la/conftest.py:44-65."""
    O = oracle
    p = P.tri_p1(5, 4)
    maps = O.make_index_maps([p.ndofs], [[]], [[]])
    rows = np.repeat(np.arange(p.ndofs, dtype=np.int32), p.ndofs)
    cols = np.tile(np.arange(p.ndofs, dtype=np.int32), p.ndofs)
    pat = O.sparsity_finalize(maps, maps, bs, [rows], [cols])[0]
    rng = np.random.default_rng(12345)
    data = rng.random(len(pat.edges) * bs[0] * bs[1])
    A = sp.bsr_matrix((data.reshape(-1, bs[0], bs[1]), pat.edges, pat.offsets))
    x = np.arange(p.ndofs * bs[1], dtype=np.float64)
    y = np.zeros(p.ndofs * bs[0])
    O.spmv(data, pat.offsets[:-1], pat.offsets[1:], pat.edges, x, y, bs[0], bs[1])
    assert np.allclose(y, A @ x)
    xt = np.arange(p.ndofs * bs[0], dtype=np.float64)
    yt = np.zeros(p.ndofs * bs[1])
    O.spmv(data, pat.offsets[:-1], pat.offsets[1:], pat.edges, xt, yt, bs[0], bs[1], transpose=True)
    assert np.allclose(yt, A.T @ xt)


def test_insert_paths(oracle):
    """python/test/unit/la/test_matrix_csr.py:36-107, 253-269."""
    O = oracle
    n = 12
    maps = O.make_index_maps([n], [[]], [[]])
    rows = np.array([0, 1, 2, 3, 4, 5], dtype=np.int32)
    r = np.repeat(rows, 6)
    c = np.tile(rows, 6)
    p1 = O.sparsity_finalize(maps, maps, (1, 1), [r], [c])[0]
    data = np.zeros(len(p1.edges))
    O.insert_csr("csr", data, p1.edges, p1.offsets, np.arange(4.0), [0, 1], [2, 3], 1, 1, "add")
    O.insert_csr("blocked", data, p1.edges, p1.offsets, np.arange(16.0), [0], [2], 2, 2, "add")  # bs=2 data into bs=1
    with pytest.raises(RuntimeError, match="Entry not in sparsity"):
        O.insert_csr("csr", data, p1.edges, p1.offsets, [1.0], [0], [7], 1, 1, "add")
    maps2 = O.make_index_maps([6], [[]], [[]])
    p2 = O.sparsity_finalize(maps2, maps2, (2, 2), [np.repeat(np.arange(3, dtype=np.int32), 3)], [np.tile(np.arange(3, dtype=np.int32), 3)])[0]
    d2 = np.zeros(len(p2.edges) * 4)
    O.insert_csr("nonblocked", d2, p2.edges, p2.offsets, np.arange(16.0), [0, 1, 2, 3], [0, 1, 2, 3], 2, 2, "set")
    dense = sp.bsr_matrix((d2.reshape(-1, 2, 2), p2.edges, p2.offsets), shape=(12, 12)).toarray()
    assert np.array_equal(dense[:4, :4], np.arange(16.0).reshape(4, 4))


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="reference shim only exists in the build container")
@pytest.mark.parametrize("bs", [(1, 1), (2, 2), (3, 3), (2, 3)])
def test_restatement_vs_reference_matrix_csr_impl(oracle, bs):
    """Differential test against the REFERENCE's la/matrix_csr_impl.h compiled in place (oracle/_ref)."""
    O = oracle
    ref = C.CDLL(REF_SO)
    rng = np.random.default_rng(7)
    nrow, ncol = 40, 50
    dense = rng.random((nrow, ncol)) < 0.3
    dense[np.arange(nrow), np.arange(nrow)] = True
    row_ptr = np.concatenate([[0], np.cumsum(dense.sum(1))]).astype(np.int64)
    cols = np.concatenate([np.flatnonzero(dense[i]) for i in range(nrow)]).astype(np.int32)
    bs2 = bs[0] * bs[1]
    for trial in range(20):
        i = int(rng.integers(nrow))
        avail = np.flatnonzero(dense[i])
        xr = np.array([i], dtype=np.int32)
        xc = rng.choice(avail, size=min(3, len(avail)), replace=False).astype(np.int32)
        x = rng.random(len(xr) * len(xc) * bs2)
        for op in (0, 1):
            d_a = rng.random(len(cols) * bs2)
            d_b = d_a.copy()
            O.insert_csr("csr", d_a, cols, row_ptr, x, xr, xc, bs[0], bs[1], "add" if op else "set")
            err = ref.ref_insert(0, bs[0], bs[1], d_b.ctypes.data_as(C.c_void_p), C.c_size_t(len(d_b)),
                                 cols.ctypes.data_as(C.c_void_p), C.c_size_t(len(cols)), row_ptr.ctypes.data_as(C.c_void_p),
                                 C.c_size_t(len(row_ptr)), x.ctypes.data_as(C.c_void_p), xr.ctypes.data_as(C.c_void_p),
                                 len(xr), xc.ctypes.data_as(C.c_void_p), len(xc), op)
            assert err == 0
            assert np.array_equal(d_a, d_b)
    # spmv / spmvT bit-exact
    vals = rng.random(len(cols) * bs2)
    x = rng.random(ncol * bs[1])
    y_a = rng.random(nrow * bs[0])
    y_b = y_a.copy()
    O.spmv(vals, row_ptr[:-1], row_ptr[1:], cols, x, y_a, bs[0], bs[1])
    ref.ref_spmv.restype = None
    rb, re = np.ascontiguousarray(row_ptr[:-1]), np.ascontiguousarray(row_ptr[1:])
    ref.ref_spmv(0, vals.ctypes.data_as(C.c_void_p), C.c_size_t(len(vals)), rb.ctypes.data_as(C.c_void_p),
                 re.ctypes.data_as(C.c_void_p), C.c_size_t(nrow), cols.ctypes.data_as(C.c_void_p), C.c_size_t(len(cols)),
                 x.ctypes.data_as(C.c_void_p), C.c_size_t(len(x)), y_b.ctypes.data_as(C.c_void_p), C.c_size_t(len(y_b)),
                 bs[0], bs[1])
    assert np.array_equal(y_a, y_b)
    xt = rng.random(nrow * bs[0])
    yt_a = np.zeros(ncol * bs[1])
    yt_b = yt_a.copy()
    O.spmv(vals, row_ptr[:-1], row_ptr[1:], cols, xt, yt_a, bs[0], bs[1], transpose=True)
    ref.ref_spmv(1, vals.ctypes.data_as(C.c_void_p), C.c_size_t(len(vals)), rb.ctypes.data_as(C.c_void_p),
                 re.ctypes.data_as(C.c_void_p), C.c_size_t(nrow), cols.ctypes.data_as(C.c_void_p), C.c_size_t(len(cols)),
                 xt.ctypes.data_as(C.c_void_p), C.c_size_t(len(xt)), yt_b.ctypes.data_as(C.c_void_p),
                 C.c_size_t(len(yt_b)), bs[0], bs[1])
    assert np.array_equal(yt_a, yt_b)
    # out-of-pattern entry raises in both
    d = np.zeros(len(cols) * bs2)
    missing = int(np.flatnonzero(~dense[0])[0])
    with pytest.raises(RuntimeError):
        O.insert_csr("csr", d, cols, row_ptr, np.zeros(bs2), [0], [missing], bs[0], bs[1], "add")
    xr, xc, x = np.array([0], dtype=np.int32), np.array([missing], dtype=np.int32), np.zeros(bs2)
    assert ref.ref_insert(0, bs[0], bs[1], d.ctypes.data_as(C.c_void_p), C.c_size_t(len(d)), cols.ctypes.data_as(C.c_void_p),
                          C.c_size_t(len(cols)), row_ptr.ctypes.data_as(C.c_void_p), C.c_size_t(len(row_ptr)),
                          x.ctypes.data_as(C.c_void_p), xr.ctypes.data_as(C.c_void_p), 1, xc.ctypes.data_as(C.c_void_p), 1, 1) == -1


GOLDEN_NPZ = os.path.join(os.path.dirname(__file__), "golden", "ref_csr_vectors.npz")


@pytest.mark.parametrize("bs", [(1, 1), (2, 2), (3, 3), (2, 3)])
def test_oracle_vs_committed_reference_vectors(oracle, bs):
    """The oracle against outputs of the REFERENCE's la/matrix_csr_impl.h (insert_csr :67-109, insert_blocked_csr
    :112-156, insert_nonblocked_csr :159-201, spmv / spmvT :204-281) recorded by tests/golden/make_ref_csr_vectors.py
    from the library compiled in place (oracle/_ref).  Runs wherever the repo is, also without /root/reference."""
    O = oracle
    g = np.load(GOLDEN_NPZ)
    tag = f"bs{bs[0]}{bs[1]}"
    row_ptr, cols = g["row_ptr"], g["cols"]
    for name in ("csr", "blocked"):
        rp, cl = (row_ptr, cols) if name == "csr" else (g[f"{tag}_scalar_row_ptr"], g[f"{tag}_scalar_cols"])
        data = g[f"{tag}_{name}_data0"].copy()
        for k in range(int(g[f"{tag}_{name}_nops"])):
            O.insert_csr(name, data, cl, rp, g[f"{tag}_{name}_op{k}_x"], g[f"{tag}_{name}_op{k}_rows"],
                         g[f"{tag}_{name}_op{k}_cols"], bs[0], bs[1], "add" if int(g[f"{tag}_{name}_op{k}_add"]) else "set")
        assert np.array_equal(data, g[f"{tag}_{name}_data1"]), name
    data = g[f"{tag}_nonblocked_data0"].copy()
    O.insert_csr("nonblocked", data, cols, row_ptr, g[f"{tag}_nonblocked_x"], g[f"{tag}_nonblocked_rows"],
                 g[f"{tag}_nonblocked_cols"], bs[0], bs[1], "add")
    assert np.array_equal(data, g[f"{tag}_nonblocked_data1"])
    y = g[f"{tag}_spmv_y0"].copy()
    O.spmv(g[f"{tag}_spmv_vals"], row_ptr[:-1], row_ptr[1:], cols, g[f"{tag}_spmv_x"], y, bs[0], bs[1])
    assert np.array_equal(y, g[f"{tag}_spmv_y1"])
    yt = g[f"{tag}_spmvT_y0"].copy()
    O.spmv(g[f"{tag}_spmv_vals"], row_ptr[:-1], row_ptr[1:], cols, g[f"{tag}_spmvT_x"], yt, bs[0], bs[1], transpose=True)
    assert np.array_equal(yt, g[f"{tag}_spmvT_y1"])


@pytest.mark.parametrize("degree", [1, 2])
def test_action_and_functional_kernels(oracle, degree):
    """The matrix-free kernels of SURVEY.md §8f (cpp/demo/poisson_matrix_free/poisson.py): action(a, ui) equals the
    assembled matrix times ui; it annihilates constants; the functional inner(w, w)*dx integrates the P1
    interpolant of x0 over the unit cube to 1/3 (and a constant c to c^2)."""
    O = oracle
    p = P.tet_p1(4, numbering="random", seed=1) if degree == 1 else P.tet_p2(3)
    kA, kM = ((O.K_POISSON_P1_TET_A, O.K_ACTION_POISSON_P1_TET_L) if degree == 1
              else (O.K_POISSON_P2_TET_A, O.K_ACTION_POISSON_P2_TET_L))
    pat, A = P.oracle_assemble_matrix(O, p, kA, constants=np.array([2.0]))
    Asp = sp.csr_matrix((A, pat.edges, pat.offsets), shape=(p.ndofs, p.ndofs))
    w = np.random.default_rng(0).random(p.ndofs)
    y = P.oracle_assemble_vector(O, p, kM, coeff=(w, p.dofmap, 1), constants=np.array([2.0]))
    assert np.max(np.abs(y - Asp @ w)) <= 1e-12 * np.max(np.abs(y))
    y1 = P.oracle_assemble_vector(O, p, kM, coeff=(np.ones(p.ndofs), p.dofmap, 1), constants=np.array([2.0]))
    assert np.max(np.abs(y1)) <= 1e-12 * np.max(np.abs(A))
    if degree == 1:
        cells = np.arange(len(p.dofmap), dtype=np.int32)
        for vals, expect in ((p.dof_coords[:, 0], 1.0 / 3.0), (np.full(p.ndofs, 1.5), 2.25)):
            coeffs = np.zeros((len(cells), 4))
            O.pack_coefficient(coeffs, 0, vals, p.dofmap, 1, cells=cells)
            m = O.assemble_scalar(O.K_L2NORM2_P1_TET_M, p.x_dofmap, p.x, cells, coeffs=coeffs)
            assert m == pytest.approx(expect, rel=1e-13)


def test_golden_interior_facets_dS(oracle):
    """The reference's golden norm of a = inner(avg(u), avg(v))*dS on create_unit_square(12, 12), P1
    (python/test/unit/fem/test_ghost_mesh_assembly.py:104-122): pins impl::assemble_interior_facets
    (fem/assemble_matrix_impl.h:442-667), sparsitybuild::interior_facets and the macro-element kernel."""
    from dolfinx_b200 import mesh as M

    O = oracle
    p = P.tri_p1(12, 12)
    facets = M.interior_facets(p.x_dofmap, M.TRI_FACETS)
    assert facets.shape == (3 * 12 * 12 - 2 * 12, 2, 2) and np.all(facets[:, 0, 0] < facets[:, 1, 0])
    maps = O.make_index_maps([p.ndofs], [[]], [[]])
    r, c = O.sparsity_insert_interior_facets(facets, p.dofmap, p.dofmap)
    pat = O.sparsity_finalize(maps, maps, (1, 1), [r], [c])[0]
    data = np.zeros(len(pat.edges))
    O.assemble_matrix_interior_facets(O.K_AVG_MASS_P1_TRI_DS, p.x_dofmap, p.x, facets, p.dofmap, 1, p.dofmap, 1, data,
                                      pat.edges, pat.offsets)
    assert np.sqrt(np.sum(data**2)) == pytest.approx(2.1834054713561906, rel=1e-12)
    # total of all entries = int_dS (avg 1)(avg 1) = total length of the interior edges
    length = 2 * 11 * 1.0 + 144 * np.sqrt(2.0) / 12.0
    assert np.sum(data) == pytest.approx(length, rel=1e-13)


def _random_distributed_matrices(O, size, bs, rng, nr=None, nc=None):
    """A rectangular matrix on `size` simulated ranks with random rows, ghost columns and values."""
    nr = [5 + r for r in range(size)] if nr is None else nr
    nc = [4 + 2 * r for r in range(size)] if nc is None else nc
    coff = np.concatenate([[0], np.cumsum(nc)])
    ghosts, owners = [], []
    for r in range(size):
        other = [g for g in range(coff[-1]) if not (coff[r] <= g < coff[r + 1])]
        gs = sorted(rng.choice(other, size=min(len(other), 3 + r), replace=False).tolist()) if other else []
        ghosts.append(gs)
        owners.append([int(np.searchsorted(coff, g, side="right") - 1) for g in gs])
    m0 = O.make_index_maps(nr, [[] for _ in nr], [[] for _ in nr])
    m1 = O.make_index_maps(nc, ghosts, owners)
    mats = []
    for r in range(size):
        rp, cols, od = [0], [], []
        for _ in range(nr[r]):
            dc = sorted(rng.choice(nc[r], size=rng.integers(0, nc[r] + 1), replace=False).tolist())
            ng = len(ghosts[r])
            gc = sorted((nc[r] + rng.choice(ng, size=rng.integers(0, ng + 1), replace=False)).tolist()) if ng else []
            cols += dc + gc
            od.append(rp[-1] + len(dc))
            rp.append(len(cols))
        data = rng.random(len(cols) * bs[0] * bs[1])
        mats.append(O.OMatrix([m0[r], m1[r]], bs, data, np.array(cols, dtype=np.int32), np.array(rp, dtype=np.int64),
                              np.array(od, dtype=np.int64)))
    return mats


def _gather(mats):
    """mat_gather of python/test/unit/la/conftest.py:20-43 on simulated ranks."""
    import scipy.sparse as sps

    bs0, bs1 = mats[0].bs
    vals, cols, ptr = [], [], [np.zeros(1, dtype=np.int64)]
    for A in mats:
        nr = A.index_maps[0].size_local
        n = int(A.row_ptr[nr])
        vals.append(A.data[: n * bs0 * bs1])
        cols.append(A.index_maps[1].local_to_global(A.cols[:n]))
        ptr.append(A.row_ptr[1 : nr + 1] + ptr[-1][-1])
    shape = (mats[0].index_maps[0].size_global * bs0, mats[0].index_maps[1].size_global * bs1)
    return sps.bsr_matrix((np.concatenate(vals).reshape(-1, bs0, bs1), np.concatenate(cols), np.concatenate(ptr)), shape=shape)


@pytest.mark.parametrize("bs", [(1, 1), (2, 2), (3, 3), (2, 3), (3, 2)])
@pytest.mark.parametrize("size", [1, 2, 3, 4])
def test_transpose_vs_scipy(oracle, size, bs):
    """la::transpose restated (oracle.transpose, la/mattrans.h) against the gathered scipy transpose, the check of
    python/test/unit/la/test_transpose.py:21-68, on 1-4 simulated ranks."""
    rng = np.random.default_rng(12345)
    mats = _random_distributed_matrices(oracle, size, bs, rng)
    G = _gather(mats).toarray()
    T = oracle.transpose(mats)
    GT = _gather(T).toarray()
    assert GT.shape == G.T.shape and np.array_equal(GT, G.T)
    for A, AT in zip(mats, T):
        assert AT.index_maps[0].num_ghosts == 0 and AT.bs == (bs[1], bs[0])
        assert AT.index_maps[0].size_local == A.index_maps[1].size_local


def test_empty_and_ragged_inputs(oracle):
    """Edge cases of the path on the oracle: an empty cell list assembles nothing, an empty bc list marks nothing,
    SpMV over a matrix with empty rows, a ragged (subset) cell list equals the sum of its parts."""
    O = oracle
    p = P.tet_p1(3, numbering="random", seed=1)
    pat = P.oracle_pattern(O, p)
    data = np.zeros(len(pat.edges))
    none = np.zeros(0, dtype=np.int32)
    O.assemble_matrix(O.K_POISSON_P1_TET_A, p.x_dofmap, p.x, none, p.dofmap, 1, p.dofmap, 1, data, pat.edges, pat.offsets,
                      constants=np.array([2.0]))
    assert not data.any()
    b = np.zeros(p.ndofs)
    O.assemble_vector(O.K_LOAD_P1_TET_L, p.x_dofmap, p.x, none, p.dofmap, 1, b, coeffs=np.zeros((0, 4)))
    assert not b.any()
    markers = np.zeros(p.ndofs, dtype=np.int8)
    O.bc_mark(markers, none)
    assert not markers.any()
    O.set_diagonal(data, pat.edges, pat.offsets, 1, 1, none, 1.0)
    assert not data.any()
    # ragged split of the cell list: A(cells0) + A(cells1) == A(all)
    cells = np.arange(len(p.dofmap), dtype=np.int32)
    rng = np.random.default_rng(0)
    mask = rng.random(len(cells)) < 0.37
    parts = []
    for sub in (cells[mask], cells[~mask], cells):
        d = np.zeros(len(pat.edges))
        O.assemble_matrix(O.K_POISSON_P1_TET_A, p.x_dofmap, p.x, sub, p.dofmap, 1, p.dofmap, 1, d, pat.edges, pat.offsets,
                          constants=np.array([2.0]))
        parts.append(d)
    assert np.max(np.abs(parts[0] + parts[1] - parts[2])) <= 1e-13 * np.max(np.abs(parts[2]))
    # SpMV with empty rows: y untouched there
    row_ptr = np.array([0, 0, 2, 2, 3], dtype=np.int64)
    cols = np.array([0, 3, 1], dtype=np.int32)
    vals = np.array([1.0, 2.0, 3.0])
    x = np.array([1.0, 10.0, 100.0, 1000.0])
    y = np.full(4, 7.0)
    O.spmv(vals, row_ptr[:-1], row_ptr[1:], cols, x, y, 1, 1)
    assert np.array_equal(y, np.array([7.0, 7.0 + 1.0 + 2000.0, 7.0, 7.0 + 30.0]))


def _serial_omatrix(O, S):
    """A scipy CSR matrix (sorted columns, explicit zeros kept) as a one-rank OMatrix."""
    S = S.tocsr()
    S.sort_indices()
    m0 = O.make_index_maps([S.shape[0]], [[]], [[]])[0]
    m1 = O.make_index_maps([S.shape[1]], [[]], [[]])[0]
    return O.OMatrix([m0, m1], (1, 1), S.data.astype(np.float64), S.indices.astype(np.int32), S.indptr.astype(np.int64),
                     S.indptr[1:].astype(np.int64))


@pytest.mark.parametrize("shape", [(7, 7, 7), (9, 5, 11), (1, 6, 3)])
def test_matmul_local_vs_scipy(oracle, shape):
    """impl::matmul restated (oracle.matmul_local, la/matmul.h:395-536) against scipy, the check of
    python/test/unit/la/test_matmul.py:21-118: values, sorted columns, and no stored zeros - neither from zero
    factors nor from exact cancellation."""
    import scipy.sparse as sps

    rng = np.random.default_rng(12345)
    n, k, m = shape
    A = sps.random(n, k, density=0.5, random_state=1, format="csr", dtype=np.float64)
    B = sps.random(k, m, density=0.5, random_state=2, format="csr", dtype=np.float64)
    if k >= 2 and n >= 1 and A.nnz and B.nnz:
        # exact cancellation: A[0, 0] B[0, 0] + A[0, 1] B[1, 0] = 0; and an explicit zero in B
        A = A.tolil()
        B = B.tolil()
        A[0, 0], A[0, 1] = 2.0, -4.0
        B[0, 0], B[1, 0] = 1.0, 0.5
        for j in range(2, k):
            B[j, 0] = 0.0
        A, B = A.tocsr(), B.tocsr()
        # the last stored entry of B becomes an explicit zero
        B = sps.csr_matrix((np.where(np.arange(B.nnz) == B.nnz - 1, 0.0, B.data), B.indices, B.indptr), shape=B.shape)
    rp, od, cols, vals = oracle.matmul_local(_serial_omatrix(oracle, A), _serial_omatrix(oracle, B))
    C = sps.csr_matrix((vals, cols, rp), shape=(n, m))
    ref = (A @ B).toarray()
    assert np.allclose(C.toarray(), ref, rtol=1e-14, atol=1e-15)
    assert np.all(vals != 0.0)
    for i in range(n):
        assert np.all(np.diff(cols[rp[i]:rp[i + 1]]) > 0)
    assert np.array_equal(od, np.diff(rp))  # every column is owned on one rank
    if k >= 2 and A.nnz and B.nnz:
        assert ref[0, 0] == 0.0 and 0 not in cols[rp[0]:rp[1]]
    with pytest.raises(RuntimeError):
        bad = _serial_omatrix(oracle, A)
        bad.bs = (2, 2)
        oracle.matmul_local(bad, _serial_omatrix(oracle, B))


@pytest.mark.parametrize("size", [1, 2, 3, 4])
def test_matmul_distributed_vs_scipy(oracle, size):
    """la::matmul restated on simulated ranks (oracle.matmul: fetch_ghost_rows + impl::matmul, la/matmul.h) against the
    product of the gathered scipy matrices - the check of python/test/unit/la/test_matmul.py:21-80 (square and
    rectangular): values, sorted columns, diagonal block first, no ghost rows."""
    rng = np.random.default_rng(12345)
    A = _random_distributed_matrices(oracle, size, (1, 1), rng)
    ncA = [m.index_maps[1].size_local for m in A]
    B = _random_distributed_matrices(oracle, size, (1, 1), rng, nr=ncA, nc=[3 + r for r in range(size)])
    C = oracle.matmul(A, B)
    G = (_gather(A).tocsr() @ _gather(B).tocsr()).toarray()
    GC = _gather(C).toarray()
    assert GC.shape == G.shape and np.allclose(GC, G, rtol=1e-13, atol=1e-14)
    for a, c in zip(A, C):
        assert c.index_maps[0].num_ghosts == 0 and c.index_maps[0].size_local == a.index_maps[0].size_local
        nl = c.index_maps[1].size_local
        for i in range(c.index_maps[0].size_local):
            row = c.cols[c.row_ptr[i]:c.row_ptr[i + 1]]
            assert np.all(np.diff(row) > 0)
            assert np.all(row[: c.off_diag_offset[i] - c.row_ptr[i]] < nl) and np.all(row[c.off_diag_offset[i] - c.row_ptr[i]:] >= nl)
        assert np.all(c.data != 0.0)


def test_facet_functionals_and_interior_facet_vector_golden(oracle):
    """Facet branches of fem::assemble_scalar / assemble_vector pinned on the reference's own expectations:
    assemble_scalar(1*dS) on the N x N unit square = 2 (N - 1) + N sqrt(2) (python/test/unit/fem/
    test_assemble_domains.py:203-210); inner(f2, f2)*ds = 4 * 4 and inner(avg(f3), avg(f3))*dS = 9 * that length for
    f2 = 2, f3 = 3 (test_additivity, :213-240); conj(avg(v))*dS (test_assembler.py:1003) sums to the same length,
    and every entry equals the hat-function integrals over the interior edges at the vertex (brute force)."""
    from dolfinx_b200 import mesh as M

    N = 10
    p = P.tri_p1(N, N)
    facets = M.interior_facets(p.x_dofmap, M.TRI_FACETS)
    ext = M.exterior_facets(p.x_dofmap, M.TRI_FACETS)
    length = 2 * (N - 1) + N * np.sqrt(2.0)
    val = oracle.assemble_scalar_interior_facets(oracle.K_ONE_TRI_DS_M, p.x_dofmap, p.x, facets)
    assert val == pytest.approx(length, rel=1e-13)
    f3 = np.full(p.ndofs, 3.0)
    w3 = np.concatenate([f3[p.dofmap[facets[:, 0, 0]]], f3[p.dofmap[facets[:, 1, 0]]]], axis=1)
    j3 = oracle.assemble_scalar_interior_facets(oracle.K_AVG2_COEFF_P1_TRI_DS_M, p.x_dofmap, p.x, facets, coeffs=w3)
    assert j3 == pytest.approx(9.0 * length, rel=1e-13)
    f2 = np.full(p.ndofs, 2.0)
    j2 = oracle.assemble_scalar_facets(oracle.K_COEFF2_P1_TRI_FACET_M, p.x_dofmap, p.x, ext, coeffs=f2[p.dofmap[ext[:, 0]]])
    assert j2 == pytest.approx(16.0, rel=1e-13)
    # a non-constant coefficient: f = x + 2 y, int_{boundary} f^2 ds by hand = sum over the four sides
    f = p.dof_coords[:, 0] + 2 * p.dof_coords[:, 1]
    jf = oracle.assemble_scalar_facets(oracle.K_COEFF2_P1_TRI_FACET_M, p.x_dofmap, p.x, ext, coeffs=f[p.dofmap[ext[:, 0]]])
    # y=0: int x^2 = 1/3; y=1: int (x+2)^2 = 19/3; x=0: int 4y^2 = 4/3; x=1: int (1+2y)^2 = 13/3
    assert jf == pytest.approx((1 + 19 + 4 + 13) / 3.0, rel=1e-13)
    b = np.zeros(p.ndofs)
    oracle.assemble_vector_interior_facets(oracle.K_AVG_LOAD_P1_TRI_DS_L, p.x_dofmap, p.x, facets, p.dofmap, 1, b)
    assert b.sum() == pytest.approx(length, rel=1e-13)
    brute = np.zeros(p.ndofs)
    for (c0, l0), (c1, l1) in facets:
        vs = [v for k, v in enumerate(p.x_dofmap[c0]) if k != l0]
        ln = np.linalg.norm(p.x[vs[0]] - p.x[vs[1]])
        for v in vs:  # continuous P1: avg(v) = v on the edge, int phi = len / 2
            brute[p.dofmap[c0][list(p.x_dofmap[c0]).index(v)]] += 0.5 * ln
    assert np.max(np.abs(b - brute)) <= 1e-13


def test_two_coefficient_kernel_golden(oracle):
    """L = f g v dx with two P1 coefficients in ONE integral: sum_i b_i = int f g, exact for the quadrature used
    (f = 1 + x, g = 2 - y + z on the unit cube: int = 3/2 * ... computed by hand = 3.0)."""
    p = P.tet_p1(4)
    dc = p.dof_coords
    f, g = 1 + dc[:, 0], 2 - dc[:, 1] + dc[:, 2]
    cells = np.arange(len(p.dofmap), dtype=np.int32)
    w = np.concatenate([f[p.dofmap], g[p.dofmap]], axis=1)
    b = np.zeros(p.ndofs)
    oracle.assemble_vector(oracle.K_LOAD_PROD_P1_TET_L, p.x_dofmap, p.x, cells, p.dofmap, 1, b, coeffs=w)
    # int (1 + x)(2 - y + z) over the unit cube = (3/2) * 2 = 3 (y and z terms cancel)
    assert b.sum() == pytest.approx(3.0, rel=1e-13)
