"""Shared problem builders for the tests: the same arrays feed the oracle and the CUDA path."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from dolfinx_b200 import mesh as M


@dataclass
class Problem:
    x: np.ndarray  # (N,3) geometry
    x_dofmap: np.ndarray  # (C,nx) int32
    dofmap: np.ndarray  # (C,nd) int32
    ndofs: int  # number of (block) dofs
    bs: int
    cell: str
    dof_coords: np.ndarray = None  # (ndofs,3) coordinates of the dof nodes (P1/Q1/P2)


def _permute(dofmap, ndofs, rng):
    perm = rng.permutation(ndofs).astype(np.int32)
    return perm[dofmap], perm


def tet_p1(n, numbering="first_touch", seed=0, shuffle_geometry=True):
    n3 = (n, n, n) if np.isscalar(n) else tuple(n)
    x = M.box_vertices(n3)
    cells = M.box_tets(n3)
    nv = len(x)
    rng = np.random.default_rng(seed)
    # geometry numbering independent of the dof numbering (they differ in DOLFINx too)
    if shuffle_geometry:
        gperm = rng.permutation(nv).astype(np.int32)
        xg = np.empty_like(x)
        xg[gperm] = x
        x_dofmap = gperm[cells]
    else:
        xg, x_dofmap = x, cells.copy()
    if numbering == "lex":
        dofmap, new = cells.copy(), np.arange(nv, dtype=np.int32)
    elif numbering == "first_touch":
        new = M.first_touch_numbering(cells, nv)
        dofmap = new[cells]
    elif numbering == "random":
        dofmap, new = _permute(cells, nv, rng)
    else:
        raise ValueError(numbering)
    dc = np.empty_like(x)
    dc[new] = x
    return Problem(xg, np.ascontiguousarray(x_dofmap, dtype=np.int32), np.ascontiguousarray(dofmap, dtype=np.int32), nv, 1, "tetrahedron", dc)


def tet_p2(n, seed=0):
    n3 = (n, n, n) if np.isscalar(n) else tuple(n)
    x = M.box_vertices(n3)
    cells = M.box_tets(n3)
    dofmap, ndofs = M.p2_tet_dofmap(cells, len(x))
    # dof coordinates: vertices and edge midpoints
    dc = np.zeros((ndofs, 3))
    dc[dofmap[:, :4].reshape(-1)] = x[cells.reshape(-1)]
    for k, (a, b) in enumerate(M.TET_EDGES):
        dc[dofmap[:, 4 + k]] = 0.5 * (x[cells[:, a]] + x[cells[:, b]])
    return Problem(x, cells.copy(), dofmap, ndofs, 1, "tetrahedron", dc)


def hex_q1(n, bs=3, numbering="first_touch", seed=0, skew=False, perturb=0.0):
    n3 = (n, n, n) if np.isscalar(n) else tuple(n)
    x = M.box_vertices(n3)
    if skew:  # affine map: still parallelepipeds
        T = np.array([[1.0, 0.2, 0.1], [0.0, 0.9, 0.3], [0.05, 0.0, 1.1]])
        x = x @ T.T
    if perturb:  # general trilinear (non-affine) cells
        x = x + perturb / max(n3) * (np.random.default_rng(seed + 17).random(x.shape) - 0.5)
    cells = M.box_hexes(n3)
    nv = len(x)
    if numbering == "lex":
        new = np.arange(nv, dtype=np.int32)
    elif numbering == "random":
        new = np.random.default_rng(seed).permutation(nv).astype(np.int32)
    else:
        new = M.first_touch_numbering(cells, nv)
    dofmap = new[cells]
    dc = np.empty_like(x)
    dc[new] = x
    return Problem(x, cells.copy(), np.ascontiguousarray(dofmap, dtype=np.int32), nv, bs, "hexahedron", dc)


def tri_p1(nx, ny):
    x, cells = M.unit_square_tris(nx, ny)
    return Problem(x, cells.copy(), cells.copy(), len(x), 1, "triangle", x.copy())


# ---- oracle-side helpers -----------------------------------------------------------------------
def oracle_pattern(O, p: Problem, cells=None):
    maps = O.make_index_maps([p.ndofs], [[]], [[]])
    c = np.arange(len(p.dofmap)) if cells is None else cells
    r, cc = O.sparsity_insert_cells(c, p.dofmap, p.dofmap)
    return O.sparsity_finalize(maps, maps, (p.bs, p.bs), [r], [cc])[0]


def oracle_assemble_matrix(O, p: Problem, kernel_id, pat=None, constants=None, coeff=None, bc=None, data=None):
    pat = oracle_pattern(O, p) if pat is None else pat
    if data is None:
        data = np.zeros(len(pat.edges) * p.bs * p.bs)
    cells = np.arange(len(p.dofmap), dtype=np.int32)
    coeffs = None
    if coeff is not None:
        v, cdm, cbs = coeff
        coeffs = np.zeros((len(cells), cdm.shape[1] * cbs))
        O.pack_coefficient(coeffs, 0, v, cdm, cbs, cells=cells)
    O.assemble_matrix(kernel_id, p.x_dofmap, p.x, cells, p.dofmap, p.bs, p.dofmap, p.bs, data, pat.edges, pat.offsets,
                      bc0=bc, bc1=bc, coeffs=coeffs, constants=constants)
    return pat, data


def oracle_assemble_vector(O, p: Problem, kernel_id, coeff=None, constants=None, b=None):
    b = np.zeros(p.ndofs * p.bs) if b is None else b
    cells = np.arange(len(p.dofmap), dtype=np.int32)
    coeffs = None
    if coeff is not None:
        v, cdm, cbs = coeff
        coeffs = np.zeros((len(cells), cdm.shape[1] * cbs))
        O.pack_coefficient(coeffs, 0, v, cdm, cbs, cells=cells)
    O.assemble_vector(kernel_id, p.x_dofmap, p.x, cells, p.dofmap, p.bs, b, coeffs=coeffs, constants=constants)
    return b


def source_f(dc):
    """f of cpp/demo/poisson/main.cpp:177-185 sampled at dof coordinates."""
    return 10.0 * np.exp(-((dc[:, 0] - 0.5) ** 2 + (dc[:, 1] - 0.5) ** 2) / 0.02)


def row_scaled_error(a, ref, row_ptr, bs2=1):
    """max_ij |a_ij - ref_ij| / max_k |ref_ik|  (how '1e-12 relative' must be read)."""
    nrows = len(row_ptr) - 1
    lens = np.diff(row_ptr) * bs2
    rows = np.repeat(np.arange(nrows), lens)
    rowmax = np.zeros(nrows)
    np.maximum.at(rowmax, rows, np.abs(ref))
    rowmax[rowmax == 0] = 1.0
    return float(np.max(np.abs(a - ref) / rowmax[rows])) if len(ref) else 0.0
