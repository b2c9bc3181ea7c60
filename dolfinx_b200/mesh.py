"""Synthetic mesh / dofmap / partition fixtures (the INPUTS of the hot path).

The mesh, dofmap and partition builders of DOLFINx are out of scope (SURVEY.md
§2.1, §8d): they only produce the arrays the assembly path reads.  This module
restates the *conventions* of the reference's box generators so that the
fixtures have the array layout the reference would hand to the assembler:

* vertex lattice / coordinates: ``create_geom``      cpp/dolfinx/mesh/generation.h:333-375
* 6-tet split of each cube:     ``build_tet``        cpp/dolfinx/mesh/generation.h:377-427
* hexahedron node order:        ``build_hex``        cpp/dolfinx/mesh/generation.h:429-472
* unit square triangles:        ``build_tri``        cpp/dolfinx/mesh/generation.h:522-680 (diagonal "right")
* first-touch dof numbering:    ``compute_reordering_map`` cpp/dolfinx/fem/dofmapbuilder.cpp:446-459
* tetrahedron edge numbering (Basix): e0=(2,3) e1=(1,3) e2=(1,2) e3=(0,3) e4=(0,2) e5=(0,1)
  (corroborated by cpp/dolfinx/io/cells.cpp:273)

Everything here is host-side numpy; nothing in this file is on the timed path.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# Local vertex pairs of the 6 tetrahedron edges (Basix order)
TET_EDGES = np.array([[2, 3], [1, 3], [1, 2], [0, 3], [0, 2], [0, 1]], dtype=np.int64)
# Tetrahedron facet i is opposite vertex i (Basix sub-entity numbering)
TET_FACETS = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]], dtype=np.int64)
# Triangle facet (edge) i is opposite vertex i
TRI_FACETS = np.array([[1, 2], [0, 2], [0, 1]], dtype=np.int64)
# Hexahedron facets (Basix: sorted vertex tuples)
HEX_FACETS = np.array(
    [[0, 1, 2, 3], [0, 1, 4, 5], [0, 2, 4, 6], [1, 3, 5, 7], [2, 3, 6, 7], [4, 5, 6, 7]],
    dtype=np.int64,
)

# Offsets (dx, dy, dz) of the cube corners v0..v7 as named in build_tet/build_hex
_CORNER = np.array(
    [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]],
    dtype=np.int64,
)
# generation.h:418-420 — {v0,v1,v3,v7},{v0,v1,v7,v5},{v0,v5,v7,v4},{v0,v3,v2,v7},{v0,v6,v4,v7},{v0,v2,v6,v7}
_TET_TABLE = np.array(
    [[0, 1, 3, 7], [0, 1, 7, 5], [0, 5, 7, 4], [0, 3, 2, 7], [0, 6, 4, 7], [0, 2, 6, 7]],
    dtype=np.int64,
)


def box_vertices(n, p0=(0.0, 0.0, 0.0), p1=(1.0, 1.0, 1.0), origin=(0, 0, 0), ntot=None):
    """Coordinates of the (nx+1)(ny+1)(nz+1) lattice, x fastest (generation.h:333-375).

    ``origin``/``ntot`` let a rank generate the sub-lattice of a larger global
    box: vertex (ix,iy,iz) has coordinate p0 + (origin+idx) * (p1-p0)/ntot.
    """
    nx, ny, nz = n
    ntot = n if ntot is None else ntot
    ext = [(p1[i] - p0[i]) / float(ntot[i]) for i in range(3)]
    ix = np.arange(nx + 1, dtype=np.float64) + origin[0]
    iy = np.arange(ny + 1, dtype=np.float64) + origin[1]
    iz = np.arange(nz + 1, dtype=np.float64) + origin[2]
    x = np.empty(((nz + 1), (ny + 1), (nx + 1), 3), dtype=np.float64)
    x[..., 0] = (p0[0] + ix * ext[0])[None, None, :]
    x[..., 1] = (p0[1] + iy * ext[1])[None, :, None]
    x[..., 2] = (p0[2] + iz * ext[2])[:, None, None]
    return x.reshape(-1, 3)


def _cube_corner_ids(n):
    """Lexicographic vertex id of corner k of every cube; shape (ncubes, 8)."""
    nx, ny, nz = n
    iz, iy, ix = np.meshgrid(
        np.arange(nz, dtype=np.int64), np.arange(ny, dtype=np.int64), np.arange(nx, dtype=np.int64), indexing="ij"
    )
    v0 = (iz * (ny + 1) + iy) * (nx + 1) + ix
    v0 = v0.reshape(-1)
    stride = np.array([1, nx + 1, (nx + 1) * (ny + 1)], dtype=np.int64)
    return v0[:, None] + (_CORNER @ stride)[None, :]


def box_tets(n, dtype=np.int32):
    """Tetrahedra of the box in generation order: (6*ncubes, 4) lattice vertex ids."""
    c = _cube_corner_ids(n)
    cells = c[:, _TET_TABLE]  # (ncubes, 6, 4)
    return np.ascontiguousarray(cells.reshape(-1, 4).astype(dtype))


def box_hexes(n, dtype=np.int32):
    """Hexahedra (ncubes, 8) in tensor node order v0..v7 (generation.h:456-465)."""
    return np.ascontiguousarray(_cube_corner_ids(n).astype(dtype))


def unit_square_tris(nx, ny, dtype=np.int32):
    """create_unit_square(nx, ny), DiagonalType::right: {v0,v1,v3},{v0,v2,v3}.

    Returns x (N,3) with z = 0 (the assembler always sees 3 components,
    assemble_matrix_impl.h:146-148) and cells (2*nx*ny, 3).
    """
    xs = np.arange(nx + 1, dtype=np.float64) * (1.0 / nx)
    ys = np.arange(ny + 1, dtype=np.float64) * (1.0 / ny)
    x = np.zeros(((ny + 1), (nx + 1), 3))
    x[..., 0] = xs[None, :]
    x[..., 1] = ys[:, None]
    iy, ix = np.meshgrid(np.arange(ny, dtype=np.int64), np.arange(nx, dtype=np.int64), indexing="ij")
    v0 = (iy * (nx + 1) + ix).reshape(-1)
    v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
    cells = np.stack([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], 1).reshape(-1, 3)
    return x.reshape(-1, 3), np.ascontiguousarray(cells.astype(dtype))


def first_touch_numbering(dofmap: np.ndarray, ndofs: int) -> np.ndarray:
    """old index -> new index, numbering dofs in the order cells first touch them.

    Restates fem/dofmapbuilder.cpp:446-459 (serial case: every dof is owned).
    Untouched dofs are appended at the end (dofmapbuilder.cpp:461-475).
    """
    flat = dofmap.reshape(-1)
    try:  # native helper (sequential scan) when the library is built
        from . import _lib

        return _lib.host_first_touch(flat, ndofs)
    except Exception:
        pass
    uniq, first = np.unique(flat, return_index=True)
    order = uniq[np.argsort(first, kind="stable")]
    new = np.full(ndofs, -1, dtype=np.int64)
    new[order] = np.arange(order.size)
    missing = np.flatnonzero(new < 0)
    new[missing] = order.size + np.arange(missing.size)
    return new.astype(dofmap.dtype)


def tet_edge_ids(cells: np.ndarray, nverts: int):
    """Global edge id of each of the 6 local edges of every tetrahedron.

    Edges are identified by their sorted vertex pair; ids are assigned by
    first touch over cells (cell order, local edge order e0..e5).
    Returns (edge_ids (C,6) int64, num_edges).
    """
    a = cells[:, TET_EDGES[:, 0]].astype(np.int64)
    b = cells[:, TET_EDGES[:, 1]].astype(np.int64)
    key = np.minimum(a, b) * np.int64(nverts) + np.maximum(a, b)
    uniq, first, inv = np.unique(key.reshape(-1), return_index=True, return_inverse=True)
    rank_of_uniq = np.empty(uniq.size, dtype=np.int64)
    rank_of_uniq[np.argsort(first, kind="stable")] = np.arange(uniq.size)
    return rank_of_uniq[inv].reshape(key.shape), int(uniq.size)


def p2_tet_dofmap(cells: np.ndarray, nverts: int, dtype=np.int32):
    """P2 Lagrange dofmap on tetrahedra: 4 vertex dofs then 6 edge dofs (Basix order).

    Numbering is first-touch over cells (dofmapbuilder.cpp:446-459), which
    interleaves vertex and edge dofs like the reference does before its
    optional graph re-ordering.  Returns (dofmap (C,10), ndofs).
    """
    eids, nedges = tet_edge_ids(cells, nverts)
    raw = np.concatenate([cells.astype(np.int64), nverts + eids], axis=1)
    ndofs = nverts + nedges
    new = first_touch_numbering(np.ascontiguousarray(raw.astype(np.int64)), ndofs)
    return np.ascontiguousarray(new[raw].astype(dtype)), ndofs


def exterior_facets(cells: np.ndarray, facet_table: np.ndarray):
    """(cell, local_facet) pairs of facets that belong to exactly one cell.

    Mirrors mesh::exterior_facet_indices + fem get_cell_entity_pairs
    (fem/utils.h:732-748, 103-124): pairs are ordered by (cell, local facet),
    which for a box is the order the reference visits them up to its facet
    numbering (an input of the path; order does not change sums beyond rounding).
    """
    C = cells.shape[0]
    nf = facet_table.shape[0]
    fv = np.sort(cells[:, facet_table].astype(np.int64), axis=2).reshape(C * nf, -1)
    _, inv, counts = np.unique(fv, axis=0, return_inverse=True, return_counts=True)
    ext = np.flatnonzero(counts[inv.reshape(-1)] == 1)
    return np.ascontiguousarray(np.stack([ext // nf, ext % nf], 1).astype(np.int32))


def interior_facets(cells: np.ndarray, facet_table: np.ndarray):
    """(F, 2, 2) array [[cell0, local_facet0], [cell1, local_facet1]] of the facets shared by two cells, cell0 <
    cell1 — the entity layout of interior-facet integrals (fem/Form.h:52-87, fem/utils.h get_cell_facet_pairs<2>)."""
    C = cells.shape[0]
    nf = facet_table.shape[0]
    fv = np.sort(cells[:, facet_table].astype(np.int64), axis=2).reshape(C * nf, -1)
    _, inv, counts = np.unique(fv, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    both = np.flatnonzero(counts[inv] == 2)
    order = np.lexsort((both, inv[both]))  # group the two (cell, facet) slots of every shared facet, lower cell first
    pairs = both[order].reshape(-1, 2)
    out = np.stack([pairs // nf, pairs % nf], axis=2)
    return np.ascontiguousarray(out[np.argsort(out[:, 0, 0] * nf + out[:, 0, 1], kind="stable")].astype(np.int32))


# ---------------------------------------------------------------------------
# Distributed box: brick partition with analytic, exchange-free numbering
# ---------------------------------------------------------------------------


@dataclass
class BoxPartition:
    """One rank's share of a global box of ``nglob`` cubes split into ``pgrid`` bricks.

    Cells: every cube of the brick (GhostMode::none semantics — no ghost cells,
    fem/utils.h:579-588 iterates owned cells only).  Vertices on a plane shared
    with a lower-coordinate brick are ghosts owned by that brick ("lower rank
    owns"; ownership is an input).  Local numbering = owned
    lattice points lexicographically (x fastest), then ghosts lexicographically.
    Global index = rank offset + local owned index, so every rank can compute
    its ghosts' global indices and owners without communication.
    """

    rank: int
    pgrid: tuple
    nglob: tuple
    nloc: tuple = field(init=False)
    origin: tuple = field(init=False)
    lo: tuple = field(init=False)

    def __post_init__(self):
        px, py, pz = self.pgrid
        r = self.rank
        self.rc = (r % px, (r // px) % py, r // (px * py))
        self.nloc = tuple(self._split(self.nglob[d], self.pgrid[d], self.rc[d])[1] for d in range(3))
        self.origin = tuple(self._split(self.nglob[d], self.pgrid[d], self.rc[d])[0] for d in range(3))
        self.lo = tuple(1 if self.rc[d] > 0 else 0 for d in range(3))

    @staticmethod
    def _split(n, p, r):
        base, rem = divmod(n, p)
        start = r * base + min(r, rem)
        return start, base + (1 if r < rem else 0)

    @staticmethod
    def rank_of(rc, pgrid):
        return rc[0] + pgrid[0] * (rc[1] + pgrid[1] * rc[2])

    def owned_shape(self, rc=None):
        """(ox, oy, oz) extents of the owned lattice box of brick ``rc``."""
        rc = self.rc if rc is None else rc
        out = []
        for d in range(3):
            n = self._split(self.nglob[d], self.pgrid[d], rc[d])[1]
            out.append(n + 1 - (1 if rc[d] > 0 else 0))
        return tuple(out)

    def num_owned(self, rc=None):
        o = self.owned_shape(rc)
        return o[0] * o[1] * o[2]

    def offset(self, rank=None):
        """Global index of the first owned vertex of ``rank`` (exclusive scan of sizes)."""
        rank = self.rank if rank is None else rank
        px, py, pz = self.pgrid
        tot = 0
        for q in range(rank):
            rc = (q % px, (q // px) % py, q // (px * py))
            tot += self.num_owned(rc)
        return tot

    def size_global(self):
        return (self.nglob[0] + 1) * (self.nglob[1] + 1) * (self.nglob[2] + 1)

    def vertex_numbering(self):
        """Returns (local_index (nz+1,ny+1,nx+1) int64, n_owned, ghosts int64, owners int32).

        ``local_index[iz,iy,ix]`` is the local vertex (= P1 dof) number of the
        lattice point: owned points first, ghosts after.
        """
        nx, ny, nz = self.nloc
        lx, ly, lz = self.lo
        iz, iy, ix = np.meshgrid(
            np.arange(nz + 1, dtype=np.int64),
            np.arange(ny + 1, dtype=np.int64),
            np.arange(nx + 1, dtype=np.int64),
            indexing="ij",
        )
        owned = (ix >= lx) & (iy >= ly) & (iz >= lz)
        ox, oy, oz = self.owned_shape()
        local = np.empty(owned.shape, dtype=np.int64)
        local[owned] = (((iz - lz) * oy + (iy - ly)) * ox + (ix - lx))[owned]
        n_owned = ox * oy * oz
        gmask = ~owned
        n_ghost = int(gmask.sum())
        local[gmask] = n_owned + np.arange(n_ghost)
        # owners and global indices of ghosts
        gx, gy, gz = ix[gmask], iy[gmask], iz[gmask]
        mvx = (gx == 0) & (lx == 1)
        mvy = (gy == 0) & (ly == 1)
        mvz = (gz == 0) & (lz == 1)
        orx = self.rc[0] - mvx.astype(np.int64)
        ory = self.rc[1] - mvy.astype(np.int64)
        orz = self.rc[2] - mvz.astype(np.int64)
        owners = (orx + self.pgrid[0] * (ory + self.pgrid[1] * orz)).astype(np.int32)
        ghosts = np.empty(n_ghost, dtype=np.int64)
        for o in np.unique(owners):
            sel = owners == o
            rc = (int(o) % self.pgrid[0], (int(o) // self.pgrid[0]) % self.pgrid[1], int(o) // (self.pgrid[0] * self.pgrid[1]))
            on = tuple(self._split(self.nglob[d], self.pgrid[d], rc[d])[1] for d in range(3))
            olo = tuple(1 if rc[d] > 0 else 0 for d in range(3))
            osh = self.owned_shape(rc)
            # coordinates of the point in the owner's lattice
            px_ = np.where(mvx[sel], on[0], gx[sel])
            py_ = np.where(mvy[sel], on[1], gy[sel])
            pz_ = np.where(mvz[sel], on[2], gz[sel])
            ghosts[sel] = self.offset(int(o)) + ((pz_ - olo[2]) * osh[1] + (py_ - olo[1])) * osh[0] + (px_ - olo[0])
        return local, n_owned, ghosts, owners

    def coordinates(self, p0=(0.0, 0.0, 0.0), p1=(1.0, 1.0, 1.0)):
        """Geometry of the local lattice (all (nloc+1)^3 points, lexicographic)."""
        return box_vertices(self.nloc, p0, p1, origin=self.origin, ntot=self.nglob)

    def tets(self):
        """(x_dofmap, dofmap_p1, n_owned, ghosts, owners) for the brick's tetrahedra."""
        lat = box_tets(self.nloc, dtype=np.int64)
        local, n_owned, ghosts, owners = self.vertex_numbering()
        dm = local.reshape(-1)[lat]
        return lat.astype(np.int32), np.ascontiguousarray(dm.astype(np.int32)), n_owned, ghosts, owners

    def hexes(self):
        lat = box_hexes(self.nloc, dtype=np.int64)
        local, n_owned, ghosts, owners = self.vertex_numbering()
        dm = local.reshape(-1)[lat]
        return lat.astype(np.int32), np.ascontiguousarray(dm.astype(np.int32)), n_owned, ghosts, owners


@dataclass
class DoubledBoxPartition(BoxPartition):
    """The lattice of P2 dof nodes of a brick partition.

    On a Kuhn box every point of the DOUBLED lattice (2n+1)^3 carries exactly one P2 dof: even coordinates are
    vertices, one / two / three odd coordinates are the midpoints of the axis edges / the face diagonals / the body
    diagonal (each unit square and each cube has exactly one diagonal: (2n+1)^3 = V + E).  So the
    P2 dofs of a brick partition are numbered, owned and ghosted exactly like the vertices of the partition of the
    doubled lattice ("lower brick owns the shared plane"); this class is that partition: ``nglob`` holds 2 n, and the
    split of a direction is twice the split of the cube lattice."""

    @staticmethod
    def _split(n, p, r):
        start, count = BoxPartition._split(n // 2, p, r)
        return 2 * start, 2 * count


def p2_partition(part: "BoxPartition"):
    """P2 Lagrange dofs of the brick ``part`` of a tetrahedral (Kuhn) box.

    Returns (x_dofmap int32 (C,4), dofmap int32 (C,10), n_owned, ghosts int64, owners int32, dof_coords (ndofs,3)):
    local dofs = owned doubled-lattice points lexicographically, then ghosts (see :class:`DoubledBoxPartition`);
    cell dofs = 4 vertices, then the 6 edges in Basix order (:data:`TET_EDGES`)."""
    d2 = DoubledBoxPartition(part.rank, part.pgrid, tuple(2 * n for n in part.nglob))
    assert d2.nloc == tuple(2 * n for n in part.nloc) and d2.origin == tuple(2 * o for o in part.origin)
    local, n_owned, ghosts, owners = d2.vertex_numbering()
    lat = box_tets(part.nloc, dtype=np.int64)
    nx, ny, nz = part.nloc
    ix, iy, iz = lat % (nx + 1), (lat // (nx + 1)) % (ny + 1), lat // ((nx + 1) * (ny + 1))

    def flat2(x2, y2, z2):
        return (z2 * (2 * ny + 1) + y2) * (2 * nx + 1) + x2

    cols = [flat2(2 * ix[:, a], 2 * iy[:, a], 2 * iz[:, a]) for a in range(4)]
    cols += [flat2(ix[:, a] + ix[:, b], iy[:, a] + iy[:, b], iz[:, a] + iz[:, b]) for a, b in TET_EDGES]
    dm = local.reshape(-1)[np.stack(cols, axis=1)]
    # coordinates of the dof nodes (owned first, then ghosts), domain [0,1]^3
    ndofs = n_owned + len(ghosts)
    z2, y2, x2 = np.meshgrid(np.arange(2 * nz + 1), np.arange(2 * ny + 1), np.arange(2 * nx + 1), indexing="ij")
    dc = np.empty((ndofs, 3))
    idx = local.reshape(-1)
    for d, (c2, o, n) in enumerate(((x2, part.origin[0], part.nglob[0]), (y2, part.origin[1], part.nglob[1]),
                                    (z2, part.origin[2], part.nglob[2]))):
        dc[idx, d] = (c2.reshape(-1) + 2 * o) / (2.0 * n)
    return lat.astype(np.int32), np.ascontiguousarray(dm.astype(np.int32)), n_owned, ghosts, owners, dc


def _partition_numbering_torch(part: "BoxPartition", device):
    """Device version of :meth:`BoxPartition.vertex_numbering`: (local int32 (nz+1,ny+1,nx+1), n_owned,
    ghosts int64 numpy, owners int32 numpy)."""
    import torch

    nx, ny, nz = part.nloc
    lx, ly, lz = part.lo
    iz, iy, ix = torch.meshgrid(
        torch.arange(nz + 1, device=device, dtype=torch.int32),
        torch.arange(ny + 1, device=device, dtype=torch.int32),
        torch.arange(nx + 1, device=device, dtype=torch.int32),
        indexing="ij",
    )
    owned = (ix >= lx) & (iy >= ly) & (iz >= lz)
    ox, oy, oz = part.owned_shape()
    n_owned = ox * oy * oz
    local = (((iz - lz) * oy + (iy - ly)) * ox + (ix - lx)).to(torch.int32)
    gmask = ~owned
    gidx = torch.nonzero(gmask.reshape(-1)).reshape(-1)  # lexicographic order of the ghosts
    n_ghost = int(gidx.numel())
    local.reshape(-1)[gidx] = n_owned + torch.arange(n_ghost, device=device, dtype=torch.int32)
    gx = ix.reshape(-1)[gidx].to(torch.int64)
    gy = iy.reshape(-1)[gidx].to(torch.int64)
    gz = iz.reshape(-1)[gidx].to(torch.int64)
    mvx = (gx == 0) & (lx == 1)
    mvy = (gy == 0) & (ly == 1)
    mvz = (gz == 0) & (lz == 1)
    orx = part.rc[0] - mvx.to(torch.int64)
    ory = part.rc[1] - mvy.to(torch.int64)
    orz = part.rc[2] - mvz.to(torch.int64)
    owners = orx + part.pgrid[0] * (ory + part.pgrid[1] * orz)
    ghosts = torch.empty(n_ghost, device=device, dtype=torch.int64)
    for o in torch.unique(owners).tolist():
        sel = owners == o
        rc = (o % part.pgrid[0], (o // part.pgrid[0]) % part.pgrid[1], o // (part.pgrid[0] * part.pgrid[1]))
        on = tuple(part._split(part.nglob[d], part.pgrid[d], rc[d])[1] for d in range(3))
        olo = tuple(1 if rc[d] > 0 else 0 for d in range(3))
        osh = part.owned_shape(rc)
        px_ = torch.where(mvx[sel], torch.full_like(gx[sel], on[0]), gx[sel])
        py_ = torch.where(mvy[sel], torch.full_like(gy[sel], on[1]), gy[sel])
        pz_ = torch.where(mvz[sel], torch.full_like(gz[sel], on[2]), gz[sel])
        ghosts[sel] = part.offset(o) + ((pz_ - olo[2]) * osh[1] + (py_ - olo[1])) * osh[0] + (px_ - olo[0])
    return local, n_owned, ghosts.cpu().numpy(), owners.to(torch.int32).cpu().numpy()


def partition_cells_torch(part: "BoxPartition", cell: str, device):
    """(x (N,3) f64, x_dofmap int32, dofmap int32, local numbering, n_owned, ghosts, owners) on ``device``."""
    local, n_owned, ghosts, owners = _partition_numbering_torch(part, device)
    lat = box_tets_torch(part.nloc, device) if cell == "tet" else box_hexes_torch(part.nloc, device)
    dm = local.reshape(-1)[lat.long()].contiguous()
    x = box_vertices_torch(part.nloc, device, origin=part.origin, ntot=part.nglob)
    return x, lat, dm, local, n_owned, ghosts, owners


def pgrid_for(nranks: int):
    """Brick grid used by the benchmarks: 1, 2x1x1, 2x2x1, 2x2x2."""
    table = {1: (1, 1, 1), 2: (2, 1, 1), 3: (3, 1, 1), 4: (2, 2, 1), 6: (3, 2, 1), 8: (2, 2, 2)}
    if nranks in table:
        return table[nranks]
    return (nranks, 1, 1)


# ---------------------------------------------------------------------------
# Device-side generators (torch) for the large benchmark meshes
# ---------------------------------------------------------------------------
def box_tets_torch(n, device, origin_stride=None):
    """Same cells as :func:`box_tets` built on ``device`` (int32, (6*ncubes, 4))."""
    import torch

    nx, ny, nz = n
    iz, iy, ix = torch.meshgrid(
        torch.arange(nz, device=device, dtype=torch.int64),
        torch.arange(ny, device=device, dtype=torch.int64),
        torch.arange(nx, device=device, dtype=torch.int64),
        indexing="ij",
    )
    v0 = ((iz * (ny + 1) + iy) * (nx + 1) + ix).reshape(-1)
    stride = np.array([1, nx + 1, (nx + 1) * (ny + 1)], dtype=np.int64)
    offs = torch.from_numpy((_CORNER @ stride)[_TET_TABLE].astype(np.int32)).to(device)  # (6,4)
    return (v0.to(torch.int32)[:, None, None] + offs[None]).reshape(-1, 4).contiguous()


def box_hexes_torch(n, device):
    import torch

    nx, ny, nz = n
    iz, iy, ix = torch.meshgrid(
        torch.arange(nz, device=device, dtype=torch.int64),
        torch.arange(ny, device=device, dtype=torch.int64),
        torch.arange(nx, device=device, dtype=torch.int64),
        indexing="ij",
    )
    v0 = ((iz * (ny + 1) + iy) * (nx + 1) + ix).reshape(-1)
    stride = np.array([1, nx + 1, (nx + 1) * (ny + 1)], dtype=np.int64)
    corner = torch.from_numpy((_CORNER @ stride).astype(np.int32)).to(device)
    return (v0.to(torch.int32)[:, None] + corner[None]).contiguous()


def box_vertices_torch(n, device, p0=(0.0, 0.0, 0.0), p1=(1.0, 1.0, 1.0), origin=(0, 0, 0), ntot=None):
    import torch

    nx, ny, nz = n
    ntot = n if ntot is None else ntot
    ext = [(p1[i] - p0[i]) / float(ntot[i]) for i in range(3)]
    ax = [torch.arange(m + 1, device=device, dtype=torch.float64) + origin[d] for d, m in enumerate((nx, ny, nz))]
    x = torch.empty((nz + 1, ny + 1, nx + 1, 3), device=device, dtype=torch.float64)
    x[..., 0] = (p0[0] + ax[0] * ext[0])[None, None, :]
    x[..., 1] = (p0[1] + ax[1] * ext[1])[None, :, None]
    x[..., 2] = (p0[2] + ax[2] * ext[2])[:, None, None]
    return x.reshape(-1, 3)


def first_touch_numbering_torch(dofmap, ndofs):
    """Device version of :func:`first_touch_numbering` (scatter-min of first positions + sort)."""
    import torch

    flat = dofmap.reshape(-1).to(torch.int64)
    big = flat.numel()
    first = torch.full((ndofs,), big, device=flat.device, dtype=torch.int64)
    pos = torch.arange(big, device=flat.device, dtype=torch.int64)
    first.scatter_reduce_(0, flat, pos, reduce="amin", include_self=True)
    del pos
    order = torch.argsort(first, stable=True)
    new = torch.empty(ndofs, device=flat.device, dtype=torch.int32)
    new[order] = torch.arange(ndofs, device=flat.device, dtype=torch.int32)
    return new
