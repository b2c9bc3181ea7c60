"""Host-side logic of the Python mirror that needs no device: DirichletBC dof handling (fem/DirichletBC.h:262-281,
357-361, 465-468), Form bookkeeping (fem/Form.h:52-87, 593-604), pack_constants (fem/pack.h:578-619), IndexMap
local/global maps (common/IndexMap.cpp:957-974), BlockMode::expanded maps, mesh fixture conventions."""

import numpy as np
import pytest

from dolfinx_b200 import common, fem, mesh as M


class _V:
    """The attributes of a FunctionSpace the host logic reads."""

    def __init__(self, n_owned, n_ghost, bs, nd=4):
        comm = common.COMM_SELF
        im = common.IndexMap(comm, n_owned)
        if n_ghost:  # ghosts without a communicator: only counts matter here
            im.ghosts = np.arange(n_ghost, dtype=np.int64)
            im.owners = np.zeros(n_ghost, dtype=np.int32)
        self.dofmap = fem.DofMap(np.zeros((1, nd), dtype=np.int32), bs, im)
        self.mesh = None
        self.element = "P1"

    @property
    def space_dimension(self):
        return self.dofmap.shape[1] * self.dofmap.bs

    def contains(self, V):
        return V is self


def test_dirichletbc_unrolls_and_counts_owned():
    # block size 3: dofs are BLOCK indices, unrolled by the block size (DirichletBC.h:357-361);
    # _owned_indices0 = position of the first ghost dof in the sorted list (:262-271)
    V = _V(n_owned=10, n_ghost=4, bs=3)
    bc = fem.DirichletBC(fem.Constant([1.0, 2.0, 3.0]), np.array([0, 4, 9, 11], dtype=np.int32), V)
    dofs, n_owned = bc.dof_indices()
    assert np.array_equal(dofs, [0, 1, 2, 12, 13, 14, 27, 28, 29, 33, 34, 35])
    assert n_owned == 9  # 3 owned blocks x 3
    # scalar space: list unchanged
    V1 = _V(n_owned=10, n_ghost=0, bs=1)
    bc1 = fem.dirichletbc(fem.Constant(0.0), np.array([1, 5], dtype=np.int32), V1)
    assert np.array_equal(bc1.dof_indices()[0], [1, 5]) and bc1.dof_indices()[1] == 2
    # empty bc
    bc0 = fem.DirichletBC(fem.Constant(0.0), np.zeros(0, dtype=np.int32), V1)
    assert bc0.dof_indices()[0].size == 0 and bc0.dof_indices()[1] == 0


def test_dirichletbc_errors():
    V = _V(n_owned=10, n_ghost=0, bs=3)
    with pytest.raises(RuntimeError, match="Constant size is not equal to the block size"):
        fem.DirichletBC(fem.Constant(1.0), np.array([0], dtype=np.int32), V)  # DirichletBC.h:330-338
    with pytest.raises(RuntimeError, match="needs the function space"):
        fem.DirichletBC(fem.Constant(1.0), np.array([0], dtype=np.int32))


def test_form_bookkeeping_and_constants():
    V = _V(n_owned=8, n_ghost=0, bs=1)
    a = fem.Form([V, V], {fem.IntegralType.cell: [(0, 7, None, []), (3, 9, np.arange(2), [0])],
                          fem.IntegralType.exterior_facet: [(1, 11, np.zeros((1, 2), dtype=np.int32), [])]},
                 constants=[fem.Constant(2.0), fem.Constant([[1.0, 2.0], [3.0, 4.0]])], mesh="m")
    assert a.rank == 2 and a.mesh == "m"
    assert a.integral_ids(fem.IntegralType.cell) == [0, 3]
    assert a.integral_ids(fem.IntegralType.exterior_facet) == [1]
    assert a.integral(fem.IntegralType.cell, 3).kernel == 9 and a.integral(fem.IntegralType.cell, 3).coeffs == [0]
    # pack_constants: flattened row-major, in order (pack.h:578-619)
    assert np.array_equal(fem.pack_constants(a), [2.0, 1.0, 2.0, 3.0, 4.0])
    assert fem.pack_constants(fem.Form([V], {}, mesh="m")).size == 0

    class _F:
        function_space = V

    # coefficient_offsets: cumulative space dimensions of ALL coefficients (Form.h:593-604)
    L = fem.Form([V], {fem.IntegralType.cell: [(0, 1, None, [1])]}, coefficients=[_F(), _F()], mesh="m")
    assert L.coefficient_offsets() == [0, 4, 8]


def test_index_map_local_to_global():
    im = common.IndexMap(common.COMM_SELF, 5)
    assert im.local_range == (0, 5) and im.size_global == 5 and im.num_ghosts == 0
    assert np.array_equal(im.local_to_global(np.arange(5)), np.arange(5))
    assert len(im.src) == 0 and len(im.dest) == 0


@pytest.mark.parametrize("n", [1, 2, 3, 5])
def test_box_fixture_counts(n):
    """vertices (n+1)^3, tets 6 n^3, edges 3n(n+1)^2 + 3n^2(n+1) + n^3, P2 dofs = V + E = (2n+1)^3,
    positive orientation of every Kuhn tetrahedron's volume sum, hex box (n+1)^3 nodes / n^3 cells."""
    x = M.box_vertices((n, n, n))
    tets = M.box_tets((n, n, n))
    assert len(x) == (n + 1) ** 3 and len(tets) == 6 * n**3
    eids, nedges = M.tet_edge_ids(tets, len(x))
    assert nedges == 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n**3
    dm, ndofs = M.p2_tet_dofmap(tets, len(x))
    assert ndofs == len(x) + nedges == (2 * n + 1) ** 3
    assert dm.shape == (6 * n**3, 10) and len(np.unique(dm)) == ndofs
    p = x[tets]
    vol = np.abs(np.linalg.det(p[:, 1:] - p[:, :1])) / 6.0
    assert vol.sum() == pytest.approx(1.0, rel=1e-13) and np.allclose(vol, 1.0 / (6 * n**3))
    hexes = M.box_hexes((n, n, n))
    assert hexes.shape == (n**3, 8) and len(np.unique(hexes)) == (n + 1) ** 3
    # first-touch numbering is a permutation that numbers the dofs of cell 0 first
    new = M.first_touch_numbering(tets, len(x))
    assert np.array_equal(np.sort(new), np.arange(len(x))) and np.array_equal(new[tets[0]], np.arange(4))
