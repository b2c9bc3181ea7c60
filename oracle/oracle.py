"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

Python/numpy half of the CPU oracle: (i) ctypes bindings of ``oracle.c`` (the
arithmetic loops) and (ii) a line-by-line restatement of the reference's
index-plan construction (IndexMap, Scatterer, SparsityPattern::finalize,
MatrixCSR ghost-row plan, Vector/MatrixCSR ghost exchange) executed for N
*simulated* ranks inside one process: every MPI neighbourhood collective of the
reference is replaced by :func:`neighbor_alltoallv`, which routes the per-rank
send buffers exactly as ``MPI_Neighbor_alltoallv`` on a
``MPI_Dist_graph_create_adjacent`` communicator would.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Citations are relative to
/root/reference/cpp/dolfinx.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# Kernel ids (shared with include/bfx.h BFX_K_*)
K_LAPLACE_P1_TRI_A = 0
K_SOURCE_P1_TRI_L = 1
K_MASS_COEFF_P1_TRI_A = 2
K_LOAD_COEFF_P1_TRI_L = 3
K_FACET_MASS_P1_TRI_A = 4
K_FACET_CONST_P1_TRI_L = 5
K_POISSON_P1_TET_A = 6
K_LOAD_P1_TET_L = 7
K_POISSON_P2_TET_A = 8
K_LOAD_P2_TET_L = 9
K_ELASTICITY_Q1_HEX_A = 10
K_LOAD_Q1_HEX_L = 11
K_FACET_LOAD_P1_TET_L = 12
K_FACET_MASS_P1_TET_A = 13
K_ELASTICITY_Q1_HEX_A_G2 = 14  # oracle-only: 2x2x2 Gauss variant
K_ACTION_POISSON_P1_TET_L = 15
K_ACTION_POISSON_P2_TET_L = 16
K_L2NORM2_P1_TET_M = 17
K_AVG_MASS_P1_TRI_DS = 18
K_AVG_LOAD_P1_TRI_DS_L = 19
K_ONE_TRI_DS_M = 20
K_AVG2_COEFF_P1_TRI_DS_M = 21
K_COEFF2_P1_TRI_FACET_M = 22
K_LOAD_PROD_P1_TET_L = 23
ORACLE_MASS_P1_TET_A = 24  # oracle-only forms of the kernel plug point test
ORACLE_SOURCE_CONST_P1_TET_L = 25
ORACLE_VOLUME_TET_M = 26


def build(fast: bool = False) -> str:
    """Compile oracle.c if the shared library is missing; returns its path."""
    name = "liboracle_fast.so" if fast else "liboracle.so"
    path = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        import subprocess

        subprocess.check_call(["make", "-C", _HERE, name])
    return path


_libs = {}


def lib(fast: bool = False):
    if fast not in _libs:
        _libs[fast] = C.CDLL(build(fast))
    return _libs[fast]


def _p(a, ct):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ct))


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int8)


def tabulate(kernel_id, nA, xc, w=None, c=None, local_entity=0):
    A = np.zeros(nA)
    w, c, xc = _f64(w), _f64(c), _f64(xc)
    if isinstance(local_entity, (tuple, list)):  # interior facet: the local facet index in each of the two cells
        err = lib().orc_tabulate2(
            C.c_int(kernel_id), _p(A, C.c_double), C.c_int(nA), _p(w, C.c_double), _p(c, C.c_double), _p(xc, C.c_double),
            C.c_int(local_entity[0]), C.c_int(local_entity[1]))
        assert err == 0
        return A
    err = lib().orc_tabulate(
        C.c_int(kernel_id), _p(A, C.c_double), C.c_int(nA), _p(w, C.c_double), _p(c, C.c_double), _p(xc, C.c_double),
        C.c_int(local_entity),
    )
    assert err == 0
    return A


def insert_csr(kind, data, cols, row_ptr, x, xrows, xcols, bs0, bs1, op):
    """kind: 'csr' | 'blocked' | 'nonblocked'; op: 'set' | 'add'. Raises RuntimeError like the reference."""
    fn = {"csr": lib().orc_insert_csr, "blocked": lib().orc_insert_blocked_csr, "nonblocked": lib().orc_insert_nonblocked_csr}[kind]
    cols, xrows, xcols, x = _i32(cols), _i32(xrows), _i32(xcols), _f64(x)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    err = fn(
        _p(data, C.c_double), _p(cols, C.c_int32), _p(row_ptr, C.c_int64), _p(x, C.c_double), _p(xrows, C.c_int32),
        C.c_int(len(xrows)), _p(xcols, C.c_int32), C.c_int(len(xcols)), C.c_int(bs0), C.c_int(bs1),
        C.c_int(1 if op == "add" else 0),
    )
    if err:
        raise RuntimeError("Entry not in sparsity")


def spmv(values, row_begin, row_end, indices, x, y, bs0, bs1, transpose=False, fast=False):
    fn = lib(fast).orc_spmvT if transpose else lib(fast).orc_spmv
    fn.restype = None
    values, x, indices = _f64(values), _f64(x), _i32(indices)
    rb = np.ascontiguousarray(row_begin, dtype=np.int64)
    re = np.ascontiguousarray(row_end, dtype=np.int64)
    fn(_p(values, C.c_double), _p(rb, C.c_int64), _p(re, C.c_int64), _p(indices, C.c_int32), _p(x, C.c_double),
       _p(y, C.c_double), C.c_int(bs0), C.c_int(bs1), C.c_int64(len(rb)))


def assemble_matrix(kernel_id, x_dofmap, x, cells, dmap0, bs0, dmap1, bs1, data, cols, row_ptr, bc0=None, bc1=None,
                    coeffs=None, constants=None, entities=None, fast=False):
    """fem::assemble_matrix of one integral into CSR ``data`` (in place, +=)."""
    x_dofmap, dmap0, dmap1 = _i32(x_dofmap), _i32(dmap0), _i32(dmap1)
    cells, entities = _i32(cells), _i32(entities)
    n = len(entities) if entities is not None else len(cells)
    coeffs, constants, x = _f64(coeffs), _f64(constants), _f64(x)
    bc0, bc1 = _i8(bc0), _i8(bc1)
    cols = _i32(cols)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    cstride = 0 if coeffs is None else coeffs.shape[1]
    err = lib(fast).orc_assemble_matrix(
        C.c_int(kernel_id), _p(x_dofmap, C.c_int32), C.c_int(x_dofmap.shape[1]), _p(x, C.c_double),
        _p(cells, C.c_int32), _p(entities, C.c_int32), C.c_int64(n), _p(dmap0, C.c_int32), C.c_int(dmap0.shape[1]),
        C.c_int(bs0), _p(dmap1, C.c_int32), C.c_int(dmap1.shape[1]), C.c_int(bs1), _p(bc0, C.c_int8),
        _p(bc1, C.c_int8), _p(coeffs, C.c_double), C.c_int(cstride), _p(constants, C.c_double), _p(data, C.c_double),
        _p(cols, C.c_int32), _p(row_ptr, C.c_int64),
    )
    if err == -1:
        raise RuntimeError("Entry not in sparsity")
    assert err == 0


def lift_bc(kernel_id, x_dofmap, x, cells, dmap0, bs0, dmap1, bs1, b, bc_values1, bc_markers1, x0=None, alpha=1.0,
            coeffs=None, constants=None, entities=None):
    x_dofmap, dmap0, dmap1 = _i32(x_dofmap), _i32(dmap0), _i32(dmap1)
    cells, entities = _i32(cells), _i32(entities)
    n = len(entities) if entities is not None else len(cells)
    coeffs, constants, x = _f64(coeffs), _f64(constants), _f64(x)
    bc_values1, x0, bc_markers1 = _f64(bc_values1), _f64(x0), _i8(bc_markers1)
    cstride = 0 if coeffs is None else coeffs.shape[1]
    err = lib().orc_lift_bc(
        C.c_int(kernel_id), _p(x_dofmap, C.c_int32), C.c_int(x_dofmap.shape[1]), _p(x, C.c_double),
        _p(cells, C.c_int32), _p(entities, C.c_int32), C.c_int64(n), _p(dmap0, C.c_int32), C.c_int(dmap0.shape[1]),
        C.c_int(bs0), _p(dmap1, C.c_int32), C.c_int(dmap1.shape[1]), C.c_int(bs1), _p(coeffs, C.c_double),
        C.c_int(cstride), _p(constants, C.c_double), _p(b, C.c_double), _p(bc_values1, C.c_double),
        _p(bc_markers1, C.c_int8), _p(x0, C.c_double), C.c_double(alpha),
    )
    assert err == 0


def assemble_vector(kernel_id, x_dofmap, x, cells, dmap, bs, b, coeffs=None, constants=None, entities=None, fast=False):
    x_dofmap, dmap = _i32(x_dofmap), _i32(dmap)
    cells, entities = _i32(cells), _i32(entities)
    n = len(entities) if entities is not None else len(cells)
    coeffs, constants, x = _f64(coeffs), _f64(constants), _f64(x)
    cstride = 0 if coeffs is None else coeffs.shape[1]
    err = lib(fast).orc_assemble_vector(
        C.c_int(kernel_id), _p(x_dofmap, C.c_int32), C.c_int(x_dofmap.shape[1]), _p(x, C.c_double),
        _p(cells, C.c_int32), _p(entities, C.c_int32), C.c_int64(n), _p(dmap, C.c_int32), C.c_int(dmap.shape[1]),
        C.c_int(bs), _p(coeffs, C.c_double), C.c_int(cstride), _p(constants, C.c_double), _p(b, C.c_double),
    )
    assert err == 0


def sparsity_insert_interior_facets(facets, dofmap0, dofmap1):
    """fem::sparsitybuild::interior_facets (fem/sparsitybuild.h:52-85): for every facet insert the joint dofs
    [dofs(cell0), dofs(cell1)] x [dofs(cell0), dofs(cell1)]."""
    f = np.asarray(facets).reshape(-1, 2, 2)
    j0 = np.concatenate([np.asarray(dofmap0)[f[:, 0, 0]], np.asarray(dofmap0)[f[:, 1, 0]]], axis=1)
    j1 = np.concatenate([np.asarray(dofmap1)[f[:, 0, 0]], np.asarray(dofmap1)[f[:, 1, 0]]], axis=1)
    n0, n1 = j0.shape[1], j1.shape[1]
    rows = np.repeat(j0[:, :, None], n1, axis=2).reshape(-1)
    cols = np.repeat(j1[:, None, :], n0, axis=1).reshape(-1)
    return rows.astype(np.int32), cols.astype(np.int32)


def assemble_matrix_interior_facets(kernel_id, x_dofmap, x, facets, dmap0, bs0, dmap1, bs1, data, cols, row_ptr,
                                    bc0=None, bc1=None, constants=None):
    """impl::assemble_interior_facets (fem/assemble_matrix_impl.h:442-667), cells on both sides of every facet."""
    x_dofmap, dmap0, dmap1 = _i32(x_dofmap), _i32(dmap0), _i32(dmap1)
    facets = _i32(np.asarray(facets).reshape(-1, 4))
    cols, constants, x = _i32(cols), _f64(constants), _f64(x)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    bc0 = None if bc0 is None else np.ascontiguousarray(bc0, dtype=np.int8)
    bc1 = None if bc1 is None else np.ascontiguousarray(bc1, dtype=np.int8)
    err = lib().orc_assemble_matrix_interior_facets(
        C.c_int(kernel_id), _p(x_dofmap, C.c_int32), C.c_int(x_dofmap.shape[1]), _p(x, C.c_double),
        _p(facets, C.c_int32), C.c_int64(len(facets)), _p(dmap0, C.c_int32), C.c_int(dmap0.shape[1]), C.c_int(bs0),
        _p(dmap1, C.c_int32), C.c_int(dmap1.shape[1]), C.c_int(bs1), _p(bc0, C.c_int8), _p(bc1, C.c_int8),
        _p(constants, C.c_double), _p(data, C.c_double), _p(cols, C.c_int32), _p(row_ptr, C.c_int64))
    if err:
        raise RuntimeError("Entry not in sparsity")


def assemble_vector_interior_facets(kernel_id, x_dofmap, x, facets, dmap, bs, b, coeffs=None, constants=None):
    """impl::assemble_interior_facets of a linear form (fem/assemble_vector_impl.h:249-339); coeffs (nf, 2*cstride)."""
    x_dofmap, dmap = _i32(x_dofmap), _i32(dmap)
    facets = _i32(np.asarray(facets).reshape(-1, 4))
    coeffs, constants, x = _f64(coeffs), _f64(constants), _f64(x)
    cstride = 0 if coeffs is None else coeffs.shape[1] // 2
    err = lib().orc_assemble_vector_interior_facets(
        C.c_int(kernel_id), _p(x_dofmap, C.c_int32), C.c_int(x_dofmap.shape[1]), _p(x, C.c_double),
        _p(facets, C.c_int32), C.c_int64(len(facets)), _p(dmap, C.c_int32), C.c_int(dmap.shape[1]), C.c_int(bs),
        _p(coeffs, C.c_double), C.c_int(cstride), _p(constants, C.c_double), _p(b, C.c_double))
    assert err == 0
    return b


def assemble_scalar_facets(kernel_id, x_dofmap, x, entities, coeffs=None, constants=None):
    """fem::assemble_scalar over exterior facets (fem/assemble_scalar_impl.h:78-113); entities (n, 2)."""
    x_dofmap = _i32(x_dofmap)
    entities = _i32(np.asarray(entities).reshape(-1, 2))
    coeffs, constants, x = _f64(coeffs), _f64(constants), _f64(x)
    cstride = 0 if coeffs is None else coeffs.shape[1]
    out = C.c_double(0.0)
    err = lib().orc_assemble_scalar_facets(
        C.c_int(kernel_id), _p(x_dofmap, C.c_int32), C.c_int(x_dofmap.shape[1]), _p(x, C.c_double),
        _p(entities, C.c_int32), C.c_int64(len(entities)), _p(coeffs, C.c_double), C.c_int(cstride),
        _p(constants, C.c_double), C.byref(out))
    assert err == 0
    return out.value


def assemble_scalar_interior_facets(kernel_id, x_dofmap, x, facets, coeffs=None, constants=None):
    """fem::assemble_scalar over interior facets (fem/assemble_scalar_impl.h:122-168); coeffs (nf, 2*cstride)."""
    x_dofmap = _i32(x_dofmap)
    facets = _i32(np.asarray(facets).reshape(-1, 4))
    coeffs, constants, x = _f64(coeffs), _f64(constants), _f64(x)
    cstride = 0 if coeffs is None else coeffs.shape[1] // 2
    out = C.c_double(0.0)
    err = lib().orc_assemble_scalar_interior_facets(
        C.c_int(kernel_id), _p(x_dofmap, C.c_int32), C.c_int(x_dofmap.shape[1]), _p(x, C.c_double),
        _p(facets, C.c_int32), C.c_int64(len(facets)), _p(coeffs, C.c_double), C.c_int(cstride),
        _p(constants, C.c_double), C.byref(out))
    assert err == 0
    return out.value


def assemble_scalar(kernel_id, x_dofmap, x, cells, coeffs=None, constants=None):
    """fem::assemble_scalar over cells (fem/assemble_scalar_impl.h:32-60)."""
    x_dofmap, cells = _i32(x_dofmap), _i32(cells)
    coeffs, constants, x = _f64(coeffs), _f64(constants), _f64(x)
    cstride = 0 if coeffs is None else coeffs.shape[1]
    out = C.c_double(0.0)
    err = lib().orc_assemble_scalar(
        C.c_int(kernel_id), _p(x_dofmap, C.c_int32), C.c_int(x_dofmap.shape[1]), _p(x, C.c_double),
        _p(cells, C.c_int32), C.c_int64(len(cells)), _p(coeffs, C.c_double), C.c_int(cstride),
        _p(constants, C.c_double), C.byref(out))
    assert err == 0
    return out.value


def pack_coefficient(coeffs, offset, v, dofmap, bs, cells=None, entities=None):
    dofmap, cells, entities, v = _i32(dofmap), _i32(cells), _i32(entities), _f64(v)
    n = len(entities) if entities is not None else len(cells)
    f = lib().orc_pack_coefficient
    f.restype = None
    f(_p(coeffs, C.c_double), C.c_int(coeffs.shape[1]), C.c_int(offset), _p(v, C.c_double), _p(dofmap, C.c_int32),
      C.c_int(dofmap.shape[1]), C.c_int(bs), _p(cells, C.c_int32), _p(entities, C.c_int32), C.c_int64(n))


def bc_mark(markers, dofs0):
    dofs0 = _i32(dofs0)
    f = lib().orc_bc_mark
    f.restype = None
    f(_p(markers, C.c_int8), _p(dofs0, C.c_int32), C.c_int64(len(dofs0)))


def bc_set(x, dofs0, g, g_kind, bs, x0=None, alpha=1.0, dofs_g=None):
    dofs0, dofs_g, g, x0 = _i32(dofs0), _i32(dofs_g), _f64(g), _f64(x0)
    f = lib().orc_bc_set
    f.restype = None
    f(_p(x, C.c_double), C.c_int32(len(x)), _p(dofs0, C.c_int32), _p(dofs_g, C.c_int32), C.c_int64(len(dofs0)),
      _p(g, C.c_double), C.c_int(g_kind), C.c_int(bs), _p(x0, C.c_double), C.c_double(alpha))


def set_diagonal(data, cols, row_ptr, bs0, bs1, rows, diagonal=1.0):
    rows, cols = _i32(rows), _i32(cols)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    err = lib().orc_set_diagonal(_p(data, C.c_double), _p(cols, C.c_int32), _p(row_ptr, C.c_int64), C.c_int(bs0),
                                 C.c_int(bs1), _p(rows, C.c_int32), C.c_int64(len(rows)), C.c_double(diagonal))
    if err:
        raise RuntimeError("Entry not in sparsity")


def unroll_dofs(dofs, bs):
    """DirichletBC::unroll_dofs (fem/DirichletBC.h:273-281)."""
    dofs = np.asarray(dofs, dtype=np.int32)
    return (bs * dofs[:, None] + np.arange(bs, dtype=np.int32)[None, :]).reshape(-1)


# ---------------------------------------------------------------------------
# Simulated MPI world
# ---------------------------------------------------------------------------


def neighbor_alltoallv(out_edges, in_edges, send):
    """MPI_Neighbor_alltoallv on a dist-graph communicator.

    out_edges[r] / in_edges[r]: destination / source rank lists of rank r.
    send[r][i]: array rank r sends to out_edges[r][i].
    Returns recv[r][j] = array received from in_edges[r][j].
    """
    size = len(out_edges)
    recv = []
    for r in range(size):
        rr = []
        for s in in_edges[r]:
            i = list(out_edges[s]).index(r)
            rr.append(np.array(send[s][i], copy=True))
        recv.append(rr)
    return recv


@dataclass
class OIndexMap:
    """common::IndexMap accessor subset (common/IndexMap.h:178-257)."""

    rank: int
    size: int
    local_range: tuple
    ghosts: np.ndarray
    owners: np.ndarray
    src: np.ndarray
    dest: np.ndarray
    size_global: int

    @property
    def size_local(self):
        return self.local_range[1] - self.local_range[0]

    @property
    def num_ghosts(self):
        return len(self.ghosts)

    def local_to_global(self, local):
        """common/IndexMap.cpp:957-974"""
        local = np.asarray(local)
        out = np.empty(local.shape, dtype=np.int64)
        m = local < self.size_local
        out[m] = self.local_range[0] + local[m]
        out[~m] = self.ghosts[local[~m] - self.size_local]
        return out


def make_index_maps(size_locals, ghosts, owners):
    """IndexMap(comm, local_size, ghosts, owners) on every rank (common/IndexMap.cpp:888-932).

    offset = MPI_Exscan, size_global = MPI_Allreduce; src = sorted unique owners;
    dest = ranks that ghost my indices (build_src_dest -> NBX), sorted.
    """
    size = len(size_locals)
    offs = np.concatenate([[0], np.cumsum(size_locals)]).astype(np.int64)
    srcs = [np.unique(np.asarray(owners[r], dtype=np.int32)) for r in range(size)]
    dests = [np.array(sorted(s for s in range(size) if r in srcs[s]), dtype=np.int32) for r in range(size)]
    return [
        OIndexMap(r, size, (int(offs[r]), int(offs[r + 1])), np.asarray(ghosts[r], dtype=np.int64),
                  np.asarray(owners[r], dtype=np.int32), srcs[r], dests[r], int(offs[-1]))
        for r in range(size)
    ]


@dataclass
class OScatterer:
    """common::Scatterer plan (common/Scatterer.h:65-198)."""

    src: np.ndarray
    dest: np.ndarray
    local_inds: np.ndarray
    remote_inds: np.ndarray
    sizes_local: np.ndarray
    displs_local: np.ndarray
    sizes_remote: np.ndarray
    displs_remote: np.ndarray


def make_scatterers(maps, bs):
    size = len(maps)
    if size == 1:  # Scatterer.h:71-72
        m = maps[0]
        z = np.zeros(0, dtype=np.int32)
        return [OScatterer(m.src, m.dest, z, z, z, np.zeros(1, np.int32), z, np.zeros(1, np.int32))]
    perms, ghosts_sorted, sizes_remote, displs_remote = [], [], [], []
    for m in maps:
        # stable sort of ghost positions by owner (radix sort, Scatterer.h:98-101)
        perm = np.argsort(m.owners, kind="stable").astype(np.int32)
        owners_sorted = m.owners[perm]
        gs = m.ghosts[perm]
        sr = np.array([np.count_nonzero(owners_sorted == s) for s in m.src], dtype=np.int32)
        dr = np.concatenate([[0], np.cumsum(sr)]).astype(np.int32)
        perms.append(perm)
        ghosts_sorted.append(gs)
        sizes_remote.append(sr)
        displs_remote.append(dr)
    # comm1: ghost -> owner. out-edges = src, in-edges = dest (Scatterer.h:89-95)
    send = [[ghosts_sorted[r][displs_remote[r][i]:displs_remote[r][i + 1]] for i in range(len(maps[r].src))] for r in range(size)]
    recv = neighbor_alltoallv([m.src for m in maps], [m.dest for m in maps], send)
    out = []
    for r, m in enumerate(maps):
        sl = np.array([len(a) for a in recv[r]], dtype=np.int32)
        dl = np.concatenate([[0], np.cumsum(sl)]).astype(np.int32)
        recv_buffer = np.concatenate(recv[r]) if recv[r] else np.zeros(0, dtype=np.int64)
        assert np.all((recv_buffer >= m.local_range[0]) & (recv_buffer < m.local_range[1]))
        k = np.arange(bs, dtype=np.int64)
        local_inds = ((recv_buffer[:, None] * bs + k[None, :]) - m.local_range[0] * bs).reshape(-1).astype(np.int32)
        remote_inds = (perms[r].astype(np.int64)[:, None] * bs + k[None, :]).reshape(-1).astype(np.int32)
        out.append(OScatterer(m.src, m.dest, local_inds, remote_inds, sl * bs, dl * bs, sizes_remote[r] * bs,
                              displs_remote[r] * bs))
    return out


def vector_scatter_fwd(maps, scs, bs, xs):
    """la::Vector::scatter_fwd (la/Vector.h:219-281): owner -> ghost, in place on xs[r]."""
    size = len(maps)
    if size == 1:
        return
    # pack: buf_local[i] = x[local_inds[i]]; comm0 out-edges = dest, in-edges = src
    send = []
    for r in range(size):
        buf = xs[r][scs[r].local_inds]
        send.append([buf[scs[r].displs_local[i]:scs[r].displs_local[i + 1]] for i in range(len(scs[r].dest))])
    recv = neighbor_alltoallv([s.dest for s in scs], [s.src for s in scs], send)
    for r in range(size):
        buf_remote = np.concatenate(recv[r]) if recv[r] else np.zeros(0)
        n0 = bs * maps[r].size_local
        xs[r][n0 + scs[r].remote_inds] = buf_remote  # unpack (Vector.h:75-88)


def vector_scatter_rev(maps, scs, bs, xs, op="add"):
    """la::Vector::scatter_rev(op) (la/Vector.h:314-379): ghost -> owner."""
    size = len(maps)
    if size == 1:
        return
    send = []
    for r in range(size):
        n0 = bs * maps[r].size_local
        buf = xs[r][n0 + scs[r].remote_inds]
        send.append([buf[scs[r].displs_remote[i]:scs[r].displs_remote[i + 1]] for i in range(len(scs[r].src))])
    recv = neighbor_alltoallv([s.src for s in scs], [s.dest for s in scs], send)
    for r in range(size):
        buf_local = np.concatenate(recv[r]) if recv[r] else np.zeros(0)
        # serial unpack with op; one slot may be hit several times (Vector.h:96-114)
        for i, idx in enumerate(scs[r].local_inds):
            if op == "add":
                xs[r][idx] = xs[r][idx] + buf_local[i]
            else:
                xs[r][idx] = buf_local[i]


# ---------------------------------------------------------------------------
# la::SparsityPattern
# ---------------------------------------------------------------------------


@dataclass
class OPattern:
    """Finalised la::SparsityPattern of one rank (la/SparsityPattern.h:25-180)."""

    index_maps: list
    bs: tuple
    edges: np.ndarray  # int32 local column indices, sorted per row
    offsets: np.ndarray  # int64
    off_diagonal_offsets: np.ndarray  # int32, per row


def sparsity_insert_cells(cells, dofmap0, dofmap1):
    """fem::sparsitybuild::cells (fem/sparsitybuild.h:36-50) -> COO cache (rows, cols) in insertion order.

    SparsityPattern::insert(rows, cols) appends, for each row, all cols (SparsityPattern.cpp:207-227).
    """
    d0 = np.asarray(dofmap0)[cells]
    d1 = np.asarray(dofmap1)[cells]
    nd0, nd1 = d0.shape[1], d1.shape[1]
    rows = np.repeat(d0[:, :, None], nd1, axis=2).reshape(-1)
    cols = np.repeat(d1[:, None, :], nd0, axis=1).reshape(-1)
    return rows.astype(np.int32), cols.astype(np.int32)


def _bucket_by_row(rows, cols, num_rows):
    """bucket_by_row (la/SparsityPattern.cpp:28-44): stable counting sort by row."""
    order = np.argsort(rows, kind="stable")
    offsets = np.zeros(num_rows + 1, dtype=np.int64)
    np.add.at(offsets, np.asarray(rows, dtype=np.int64) + 1, 1)
    return np.cumsum(offsets), np.asarray(cols)[order]


def sparsity_finalize(maps0, maps1, bs, cache_rows, cache_cols):
    """SparsityPattern::finalize on every simulated rank (la/SparsityPattern.cpp:264-491)."""
    size = len(maps0)
    buckets, send, nbr = [], [], []
    for r in range(size):
        m0, m1 = maps0[r], maps1[r]
        local_size0, local_size1 = m0.size_local, m1.size_local
        num_rows0 = local_size0 + m0.num_ghosts
        off, cc = _bucket_by_row(cache_rows[r], cache_cols[r], num_rows0)
        buckets.append((off, cc))
        neighbour_rank = np.searchsorted(m0.src, m0.owners)  # :291-299
        bufs = [[] for _ in m0.src]
        for i in range(m0.num_ghosts):  # :326-351 — (global row, global col, col owner) triplets
            for k in range(off[local_size0 + i], off[local_size0 + i + 1]):
                col_local = int(cc[k])
                if col_local < local_size1:
                    trip = (int(m0.ghosts[i]), col_local + m1.local_range[0], r)
                else:
                    trip = (int(m0.ghosts[i]), int(m1.ghosts[col_local - local_size1]), int(m1.owners[col_local - local_size1]))
                bufs[neighbour_rank[i]].extend(trip)
        send.append([np.array(b, dtype=np.int64) for b in bufs])
        nbr.append(neighbour_rank)
    # :354-383 — dist graph (sources = dest0, destinations = src0): send to src0, receive from dest0
    recv = neighbor_alltoallv([m.src for m in maps0], [m.dest for m in maps0], send)
    out = []
    for r in range(size):
        m0, m1 = maps0[r], maps1[r]
        local_size0, local_size1 = m0.size_local, m1.size_local
        num_rows0 = local_size0 + m0.num_ghosts
        off, cc = buckets[r]
        col_ghosts = list(m1.ghosts)
        col_ghost_owners = list(m1.owners)
        g2l = {int(g): local_size1 + i for i, g in enumerate(col_ghosts)}
        ghost_data_in = np.concatenate(recv[r]) if recv[r] else np.zeros(0, dtype=np.int64)
        recv_rows, recv_cols = [], []
        local_i = local_size1 + len(col_ghosts)
        for i in range(0, len(ghost_data_in), 3):  # :394-423
            row_local = int(ghost_data_in[i]) - m0.local_range[0]
            col = int(ghost_data_in[i + 1])
            owner = int(ghost_data_in[i + 2])
            recv_rows.append(row_local)
            if m1.local_range[0] <= col < m1.local_range[1]:
                recv_cols.append(col - m1.local_range[0])
            else:
                if col not in g2l:
                    g2l[col] = local_i
                    col_ghosts.append(col)
                    col_ghost_owners.append(owner)
                    local_i += 1
                recv_cols.append(g2l[col])
        roff, rcc = _bucket_by_row(np.array(recv_rows, dtype=np.int64), np.array(recv_cols, dtype=np.int32), local_size0)
        # :438-478 — dedup + sort per row; vectorised (the result, sorted unique columns, is order independent)
        row_of_cache = np.repeat(np.arange(num_rows0, dtype=np.int64), np.diff(off))
        row_of_recv = np.repeat(np.arange(local_size0, dtype=np.int64), np.diff(roff))
        ncols = local_size1 + len(col_ghosts)
        keys = np.concatenate([row_of_cache * ncols + cc, row_of_recv * ncols + rcc])
        keys = np.unique(keys)
        rows_u = keys // ncols
        edges = (keys % ncols).astype(np.int32)
        offsets = np.zeros(num_rows0 + 1, dtype=np.int64)
        np.add.at(offsets, rows_u + 1, 1)
        offsets = np.cumsum(offsets)
        diag = np.zeros(num_rows0, dtype=np.int64)
        np.add.at(diag, rows_u[edges < local_size1], 1)
        new_m1 = OIndexMap(r, size, m1.local_range, np.array(col_ghosts, dtype=np.int64),
                           np.array(col_ghost_owners, dtype=np.int32), None, None, m1.size_global)
        out.append((edges, offsets, diag.astype(np.int32), new_m1))
    # :488-490 — new column IndexMap (src/dest recomputed from the extended ghost owners)
    new_maps1 = make_index_maps([m.size_local for m in maps1], [o[3].ghosts for o in out], [o[3].owners for o in out])
    return [OPattern([maps0[r], new_maps1[r]], tuple(bs), out[r][0], out[r][1], out[r][2]) for r in range(size)]


# ---------------------------------------------------------------------------
# la::MatrixCSR
# ---------------------------------------------------------------------------


@dataclass
class OMatrix:
    """la::MatrixCSR (compact block mode) of one rank: structure + ghost-row plan (la/MatrixCSR.h:628-850)."""

    index_maps: list
    bs: tuple
    data: np.ndarray
    cols: np.ndarray
    row_ptr: np.ndarray
    off_diag_offset: np.ndarray
    ghost_row_to_rank: np.ndarray = None
    val_send_disp: np.ndarray = None
    val_recv_disp: np.ndarray = None
    unpack_pos: np.ndarray = None


def make_matrices(patterns):
    size = len(patterns)
    mats, send = [], []
    for r, p in enumerate(patterns):
        m0, m1 = p.index_maps
        bs2 = p.bs[0] * p.bs[1]
        A = OMatrix(p.index_maps, p.bs, np.zeros(len(p.edges) * bs2), p.edges.copy(), p.offsets.copy(),
                    (p.off_diagonal_offsets.astype(np.int64) + p.offsets[:-1]))  # :695-703
        local_size = (m0.size_local, m1.size_local)
        A.ghost_row_to_rank = np.searchsorted(m0.src, m0.owners).astype(np.int32)  # :725-733
        data_per_proc = np.zeros(len(m0.src), dtype=np.int64)
        for i, g in enumerate(A.ghost_row_to_rank):  # :735-742
            pos = local_size[0] + i
            data_per_proc[g] += A.row_ptr[pos + 1] - A.row_ptr[pos]
        val_send_disp = np.concatenate([[0], np.cumsum(data_per_proc)]).astype(np.int64)
        bufs = [[] for _ in m0.src]
        for i, g in enumerate(A.ghost_row_to_rank):  # :750-775 — (global row, global col) pairs
            row_id = local_size[0] + i
            for j in range(A.row_ptr[row_id], A.row_ptr[row_id + 1]):
                col_local = int(A.cols[j])
                gc = col_local + m1.local_range[0] if col_local < local_size[1] else int(m1.ghosts[col_local - local_size[1]])
                bufs[g].extend((int(m0.ghosts[i]), gc))
        send.append([np.array(b, dtype=np.int64) for b in bufs])
        A.val_send_disp = (val_send_disp * bs2).astype(np.int64)  # :806-811
        mats.append(A)
    # comm: sources = dest_ranks, destinations = src_ranks (:715-721)
    recv = neighbor_alltoallv([p.index_maps[0].src for p in patterns], [p.index_maps[0].dest for p in patterns], send)
    for r, A in enumerate(mats):
        m0, m1 = A.index_maps
        bs2 = A.bs[0] * A.bs[1]
        recv_disp = np.concatenate([[0], np.cumsum([len(a) for a in recv[r]])]).astype(np.int64)
        A.val_recv_disp = bs2 * recv_disp // 2
        arr = np.concatenate(recv[r]) if recv[r] else np.zeros(0, dtype=np.int64)
        g2l = {int(g): m1.size_local + i for i, g in enumerate(m1.ghosts)}
        unpack = []
        for i in range(0, len(arr), 2):  # :820-846
            local_row = int(arr[i]) - m0.local_range[0]
            assert 0 <= local_row < m0.size_local
            local_col = int(arr[i + 1]) - m1.local_range[0]
            if local_col < 0 or local_col >= m1.size_local:
                local_col = g2l[int(arr[i + 1])]
            c0, c1 = A.row_ptr[local_row], A.row_ptr[local_row + 1]
            d = c0 + int(np.searchsorted(A.cols[c0:c1], local_col))
            assert d < c1 and A.cols[d] == local_col
            unpack.append(d)
        A.unpack_pos = np.array(unpack, dtype=np.int64)
    return mats


def matrix_scatter_rev(mats):
    """MatrixCSR::scatter_rev (la/MatrixCSR.h:399-468): ghost-row values -> owners (+=), zero ghost rows."""
    size = len(mats)
    send = []
    for A in mats:
        m0 = A.index_maps[0]
        bs2 = A.bs[0] * A.bs[1]
        bufs = [[] for _ in m0.src]
        for i, g in enumerate(A.ghost_row_to_rank):  # :406-420 pack per neighbour in ghost-row order
            r0, r1 = A.row_ptr[m0.size_local + i] * bs2, A.row_ptr[m0.size_local + i + 1] * bs2
            bufs[g].append(A.data[r0:r1])
        send.append([np.concatenate(b) if b else np.zeros(0) for b in bufs])
    recv = neighbor_alltoallv([A.index_maps[0].src for A in mats], [A.index_maps[0].dest for A in mats], send)
    for r, A in enumerate(mats):
        bs2 = A.bs[0] * A.bs[1]
        vin = np.concatenate(recv[r]) if recv[r] else np.zeros(0)
        assert len(vin) == len(A.unpack_pos) * bs2
        for i, p in enumerate(A.unpack_pos):  # :457-459 serial +=
            A.data[p * bs2:(p + 1) * bs2] += vin[i * bs2:(i + 1) * bs2]
        A.data[A.row_ptr[A.index_maps[0].size_local] * bs2:] = 0  # :465-467


def matrix_squared_norm(mats):
    """MatrixCSR::squared_norm (la/MatrixCSR.h:473-486): owned rows only, summed over ranks."""
    tot = 0.0
    for A in mats:
        bs2 = A.bs[0] * A.bs[1]
        n = A.row_ptr[A.index_maps[0].size_local] * bs2
        tot += float(np.sum(A.data[:n] ** 2))
    return tot


def matrix_mult(mats, scs_col, xs, ys):
    """MatrixCSR::mult (la/MatrixCSR.h:877-946): y += A x with the diag / off-diag split around the ghost update."""
    maps1 = [A.index_maps[1] for A in mats]
    bs1 = mats[0].bs[1]
    # x.scatter_fwd_begin(); diagonal block; x.scatter_fwd_end(); off-diagonal block
    for A, x, y in zip(mats, xs, ys):
        n = A.index_maps[0].size_local
        spmv(A.data, A.row_ptr[:n], A.off_diag_offset[:n], A.cols, x, y, A.bs[0], A.bs[1])
    vector_scatter_fwd(maps1, scs_col, bs1, xs)
    for A, x, y in zip(mats, xs, ys):
        n = A.index_maps[0].size_local
        spmv(A.data, A.off_diag_offset[:n], A.row_ptr[1:n + 1], A.cols, x, y, A.bs[0], A.bs[1])


def inner_product(maps, bs, xs, ys):
    """la::inner_product (la/Vector.h:434-460): owned entries, MPI_SUM."""
    return float(sum(np.dot(x[: bs * m.size_local], y[: bs * m.size_local]) for m, x, y in zip(maps, xs, ys)))


# ---------------------------------------------------------------------------
# la::transpose (la/mattrans.h) - SURVEY.md section 8f rank 4
# ---------------------------------------------------------------------------
def local_transpose(A: OMatrix):
    """impl::local_transpose (la/mattrans.h:47-108): transpose of the block owned rows x owned columns.

    Returns (colsT int32, row_ptrT int64, valsT); the loops are the reference's (a write cursor per column, rows in
    ascending order, every bs0 x bs1 block stored transposed)."""
    bs0, bs1 = A.bs
    n_rows, n_cols = A.index_maps[0].size_local, A.index_maps[1].size_local
    row_count = np.zeros(n_cols, dtype=np.int64)
    for i in range(n_rows):  # :64-67
        for k in range(int(A.row_ptr[i]), int(A.off_diag_offset[i])):
            row_count[A.cols[k]] += 1
    row_ptrT = np.concatenate([[0], np.cumsum(row_count)]).astype(np.int64)  # :70-72
    colsT = np.zeros(int(row_ptrT[-1]), dtype=np.int32)
    valsT = np.zeros(int(row_ptrT[-1]) * bs0 * bs1)
    cursor = row_ptrT[:-1].copy()
    blocks = A.data.reshape(-1, bs0, bs1)
    outT = valsT.reshape(-1, bs1, bs0)
    for i in range(n_rows):  # :78-104
        for k in range(int(A.row_ptr[i]), int(A.off_diag_offset[i])):
            col = int(A.cols[k])
            pos = int(cursor[col])
            cursor[col] += 1
            colsT[pos] = i
            outT[pos] = blocks[k].T
    return colsT, row_ptrT, valsT


def transpose(mats):
    """la::transpose on every (simulated) rank - la/mattrans.h:121-437.

    The entries in ghost columns travel to the column owners as (global row, global column, value) triplets over
    the neighbourhood "send to src, receive from dest" of the column map (:200-206, :290-300); the owner appends
    them to its rows in arrival order (:395-424) behind the locally transposed block; the ghost columns of the
    result are the sorted unique (global row, sender) pairs received (:335-361).  Row map of the result: the owned
    columns, no ghosts (:429)."""
    size = len(mats)
    sends = []
    for A in mats:
        m0, m1 = A.index_maps
        n_row, n_col = m0.size_local, m1.size_local
        nbs = A.bs[0] * A.bs[1]
        ghost_col_owner = np.searchsorted(m1.src, m1.owners)  # rank_to_nbr (:191-197)
        bufs = [([], [], []) for _ in m1.src]
        for i in range(n_row):  # :248-262
            for k in range(int(A.off_diag_offset[i]), int(A.row_ptr[i + 1])):
                j = int(A.cols[k]) - n_col
                b = bufs[int(ghost_col_owner[j])]
                b[0].append(m0.local_range[0] + i)
                b[1].append(int(m1.ghosts[j]))
                b[2].append(A.data[k * nbs:(k + 1) * nbs].copy())
        sends.append(bufs)
    out_edges = [A.index_maps[1].src for A in mats]
    in_edges = [A.index_maps[1].dest for A in mats]
    recv_rows = neighbor_alltoallv(out_edges, in_edges, [[np.array(b[0], dtype=np.int64) for b in s] for s in sends])
    recv_cols = neighbor_alltoallv(out_edges, in_edges, [[np.array(b[1], dtype=np.int64) for b in s] for s in sends])
    recv_vals = neighbor_alltoallv(
        out_edges, in_edges,
        [[np.concatenate(b[2]) if b[2] else np.zeros(0) for b in s] for s in sends])
    # column maps of the results: owned = rows of A, ghosts = sorted unique (global row, sender) pairs
    ghosts_T, owners_T = [], []
    for r, A in enumerate(mats):
        pairs = sorted({(int(g), int(A.index_maps[1].dest[p])) for p, arr in enumerate(recv_rows[r]) for g in arr})
        ghosts_T.append([g for g, _ in pairs])
        owners_T.append([o for _, o in pairs])
    col_maps = make_index_maps([A.index_maps[0].size_local for A in mats], ghosts_T, owners_T)
    row_maps = make_index_maps([A.index_maps[1].size_local for A in mats], [[] for _ in mats], [[] for _ in mats])
    out = []
    for r, A in enumerate(mats):
        bs0, bs1 = A.bs
        nbs = bs0 * bs1
        m1 = A.index_maps[1]
        n_row, n_col = A.index_maps[0].size_local, m1.size_local
        c0, rp0, v0 = local_transpose(A)
        count = np.diff(rp0).astype(np.int64)
        rc = np.concatenate(recv_cols[r]) if recv_cols[r] else np.zeros(0, dtype=np.int64)
        rr = np.concatenate(recv_rows[r]) if recv_rows[r] else np.zeros(0, dtype=np.int64)
        rv = np.concatenate(recv_vals[r]) if recv_vals[r] else np.zeros(0)
        for c in rc:  # :318-325
            count[int(c) - m1.local_range[0]] += 1
        row_ptr = np.concatenate([[0], np.cumsum(count)]).astype(np.int64)
        g2l = {g: n_row + i for i, g in enumerate(ghosts_T[r])}
        cols = np.zeros(int(row_ptr[-1]), dtype=np.int32)
        vals = np.zeros(int(row_ptr[-1]) * nbs)
        off = np.diff(rp0).astype(np.int64)  # off_diag_offsets (:386-393)
        cursor = row_ptr[:-1] + off
        for i in range(n_col):  # :371-381
            cols[row_ptr[i]:row_ptr[i] + off[i]] = c0[rp0[i]:rp0[i + 1]]
            vals[row_ptr[i] * nbs:(row_ptr[i] + off[i]) * nbs] = v0[rp0[i] * nbs:rp0[i + 1] * nbs]
        for k in range(len(rc)):  # :395-424
            lc = int(rc[k]) - m1.local_range[0]
            pos = int(cursor[lc])
            cursor[lc] += 1
            cols[pos] = g2l[int(rr[k])]
            vals[pos * nbs:(pos + 1) * nbs] = rv[k * nbs:(k + 1) * nbs].reshape(bs0, bs1).T.reshape(-1)
        out.append(OMatrix([row_maps[r], col_maps[r]], (bs1, bs0), vals, cols, row_ptr, row_ptr[:-1] + off))
    return out


def matmul_local(A: OMatrix, B: OMatrix, new_col_map=None, ghost_row_ptr=None, ghost_cols=None, ghost_vals=None):
    """impl::matmul (la/matmul.h:395-536), block size 1: C = A B row by row with a dense accumulator.

    Restated with the reference's order of operations, because the STRUCTURE of C depends on the values: a product
    that is exactly zero does not create an entry (:457, :477), and entries whose sum cancels to exactly zero are
    removed (:486-497) - so the summation order (entries of row i of A in storage order, for each the entries of the
    row of B in storage order) is part of the result.  Rows of B behind ghost columns of A come in through
    ``ghost_row_ptr / ghost_cols / ghost_vals`` (already in the column numbering of C), as fetch_ghost_rows (:60-390)
    delivers them; on one rank they are empty.  Returns (row_ptr int64, off_diag int32, cols int32, vals)."""
    if tuple(A.bs) != (1, 1) or tuple(B.bs) != (1, 1):
        raise RuntimeError("Currently matmul only supports block size=1")  # :549-552
    mB1 = B.index_maps[1]
    col_map_C = mB1 if new_col_map is None else new_col_map
    n_rows_A = A.index_maps[0].size_local
    n_rows_B = B.index_maps[0].size_local
    n_owned_cols_B = mB1.size_local
    n_owned_cols_C = col_map_C.size_local
    if new_col_map is None:
        remap = n_owned_cols_B + np.arange(mB1.num_ghosts)
    else:
        g2l = {int(g): n_owned_cols_C + i for i, g in enumerate(col_map_C.ghosts)}
        remap = np.array([g2l[int(g)] for g in mB1.ghosts], dtype=np.int64)
    num_cols_C = n_owned_cols_C + col_map_C.num_ghosts
    acc = np.zeros(num_cols_C)
    in_row = np.zeros(num_cols_C, dtype=bool)
    row_ptr, off_diag, cols_C, vals_C = [0], [], [], []
    for i in range(n_rows_A):
        row_cols = []
        for ka in range(int(A.row_ptr[i]), int(A.off_diag_offset[i])):  # :449-466
            j = int(A.cols[ka])
            a = A.data[ka]
            for kb in range(int(B.row_ptr[j]), int(B.row_ptr[j + 1])):
                c = int(B.cols[kb])
                k = c if c < n_owned_cols_B else int(remap[c - n_owned_cols_B])
                v = a * B.data[kb]
                if not in_row[k] and v != 0.0:
                    in_row[k] = True
                    row_cols.append(k)
                acc[k] += v
        for ka in range(int(A.off_diag_offset[i]), int(A.row_ptr[i + 1])):  # :469-484
            g = int(A.cols[ka]) - n_rows_B
            a = A.data[ka]
            for kb in range(int(ghost_row_ptr[g]), int(ghost_row_ptr[g + 1])):
                k = int(ghost_cols[kb])
                v = a * ghost_vals[kb]
                if not in_row[k] and v != 0.0:
                    in_row[k] = True
                    row_cols.append(k)
                acc[k] += v
        kept = []
        for k in row_cols:  # :486-497
            if acc[k] == 0.0:
                in_row[k] = False
            else:
                kept.append(k)
        kept.sort()  # :501
        off_diag.append(int(np.searchsorted(kept, n_owned_cols_C)))  # :505-507
        for c in kept:  # :510-516
            cols_C.append(c)
            vals_C.append(acc[c])
            acc[c] = 0.0
            in_row[c] = False
        # (columns touched only by products that were exactly zero keep acc == 0 and in_row == False)
        row_ptr.append(len(cols_C))
    return (np.array(row_ptr, dtype=np.int64), np.array(off_diag, dtype=np.int32), np.array(cols_C, dtype=np.int32),
            np.array(vals_C, dtype=np.float64))


def matmul(matsA, matsB, return_ghost_rows=False):
    """la::matmul on every (simulated) rank - la/matmul.h:538-579: C = A B for block size 1.

    fetch_ghost_rows (:79-390): every ghost column of A is a row of B on its owner; the owner sends that row (global
    columns, owner of every column, values, in its storage order); the column map of C is B's column map extended by
    the SORTED UNIQUE (global column, owner) pairs of the received entries and of B's own ghosts (:330-356) - unless
    the column map of A has no neighbours on this rank, then B's column map is kept as it is (:104-113).  The rows
    of C are then computed by impl::matmul (matmul_local).  The row map of C has no ghosts (:569-570)."""
    size = len(matsA)
    fetched, new_ghosts, new_owners = [], [], []
    for r in range(size):
        A, B = matsA[r], matsB[r]
        mA1, mB1 = A.index_maps[1], B.index_maps[1]
        if size == 1 or (len(mA1.src) == 0 and len(mA1.dest) == 0):
            fetched.append(None)
            new_ghosts.append(list(mB1.ghosts))
            new_owners.append(list(mB1.owners))
            continue
        rows = []
        for g, o in zip(mA1.ghosts, mA1.owners):
            Bo = matsB[int(o)]
            mo0, mo1 = Bo.index_maps
            lr = int(g) - mo0.local_range[0]
            assert 0 <= lr < mo0.size_local
            k0, k1 = int(Bo.row_ptr[lr]), int(Bo.row_ptr[lr + 1])
            lc = Bo.cols[k0:k1]
            gcols = mo1.local_to_global(lc)
            cown = np.where(lc < mo1.size_local, int(o), mo1.owners[np.maximum(lc - mo1.size_local, 0)] if mo1.num_ghosts
                            else int(o))
            rows.append((gcols, cown.astype(np.int64), Bo.data[k0:k1].copy()))
        lo, hi = mB1.local_range
        pairs = {(int(c), int(w)) for gc, ow, _ in rows for c, w in zip(gc, ow) if not (lo <= c < hi)}
        pairs |= {(int(g), int(o)) for g, o in zip(mB1.ghosts, mB1.owners)}
        pairs = sorted(pairs)
        fetched.append(rows)
        new_ghosts.append([g for g, _ in pairs])
        new_owners.append([o for _, o in pairs])
    col_maps = make_index_maps([B.index_maps[1].size_local for B in matsB], new_ghosts, new_owners)
    row_maps = make_index_maps([A.index_maps[0].size_local for A in matsA], [[] for _ in matsA], [[] for _ in matsA])
    out, ghost_rows = [], []
    for r in range(size):
        A, B = matsA[r], matsB[r]
        cm = col_maps[r]
        if fetched[r] is None:
            rp, od, cols, vals = matmul_local(A, B)
            ghost_rows.append((np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0)))
        else:
            g2l = {int(g): cm.size_local + i for i, g in enumerate(cm.ghosts)}
            lo = cm.local_range[0]
            grp, gcols, gvals = [0], [], []
            for gc, _, v in fetched[r]:
                gcols += [int(c) - lo if lo <= c < cm.local_range[1] else g2l[int(c)] for c in gc]
                gvals += list(v)
                grp.append(len(gcols))
            rp, od, cols, vals = matmul_local(A, B, cm, np.array(grp), np.array(gcols, dtype=np.int64), np.array(gvals))
            ghost_rows.append((np.array(grp, dtype=np.int64), np.array(gcols, dtype=np.int64), np.array(gvals)))
        out.append(OMatrix([row_maps[r], cm], (1, 1), vals, cols, rp, rp[:-1] + od.astype(np.int64)))
    return (out, ghost_rows) if return_ghost_rows else out
