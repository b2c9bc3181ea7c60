"""Host mirror of dolfinx::la — SparsityPattern, MatrixCSR, Vector — over the libbfx C-ABI.

Reference: cpp/dolfinx/la/{SparsityPattern.h,SparsityPattern.cpp,MatrixCSR.h,matrix_csr_impl.h,Vector.h}
and the Python surface python/dolfinx/la/__init__.py (same method names: ``insert``, ``finalize``,
``add``, ``set``, ``scatter_reverse``, ``squared_norm``, ``mult``, ``to_dense``, ``to_scipy``,
``scatter_forward`` …).  Values live in HBM as torch tensors (torch is the allocator); every
arithmetic call goes to a CUDA kernel of libbfx.so — there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import enum
import itertools

import numpy as np

from .common import Comm, IndexMap, Scatterer, cached_scatterer, release_with


class InsertMode(enum.IntEnum):
    add = 0
    insert = 1


class Norm(enum.IntEnum):
    l1 = 0
    l2 = 1
    linf = 2
    frobenius = 3


class BlockMode(enum.IntEnum):
    compact = 0
    expanded = 1


def _torch():
    import torch

    return torch


def _device():
    torch = _torch()
    if not torch.cuda.is_available():
        from ._lib import BfxError, ERR_NO_DEVICE

        raise BfxError(ERR_NO_DEVICE, "no CUDA device: dolfinx_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# ---------------------------------------------------------------------------------------------
# SparsityPattern
# ---------------------------------------------------------------------------------------------
class SparsityPattern:
    """la::SparsityPattern (la/SparsityPattern.h:25-180).

    ``insert`` / ``insert_diagonal`` append to a COO cache like the reference
    (SparsityPattern.cpp:194-240).  ``insert_cells`` is the bulk entry used by
    ``fem.create_sparsity_pattern`` (fem::sparsitybuild::cells, fem/sparsitybuild.h:36-50): with
    device dofmaps the pattern is built on the GPU without ever materialising the COO cache.
    ``finalize`` follows SparsityPattern.cpp:264-491 bit for bit: ghost-row entries go to the row
    owners as (global row, global col, col owner) triplets, unknown columns are appended to the
    column IndexMap in arrival order, rows are de-duplicated and sorted by local column index.
    """

    def __init__(self, comm: Comm, maps, bs):
        self.comm = comm
        self._index_maps = [maps[0], maps[1]]
        self._bs = (int(bs[0]), int(bs[1]))
        self._segments = []  # ("coo", rows, cols) | ("cells", cells, dofmap0, dofmap1)
        self._finalized = False
        self._edges = self._offsets = self._off_diag = None
        self._csr = None  # device structure (bfx_csr_t*) when built natively
        self._col_ghosts = self._col_ghost_owners = None

    @classmethod
    def from_graph(cls, comm: Comm, maps, bs, edges, offsets, off_diagonal_offsets):
        """A finalised pattern from its arrays - the reference's ``la::impl::Sparsity`` aggregate, which
        ``la::transpose`` / ``la::matmul`` hand to the MatrixCSR constructor (la/mattrans.h:153-155, 431-434)."""
        sp = cls(comm, maps, bs)
        sp._edges = np.ascontiguousarray(edges, dtype=np.int32)
        sp._offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        sp._off_diag = np.ascontiguousarray(off_diagonal_offsets, dtype=np.int32)
        sp._col_ghosts = maps[1].ghosts.copy()
        sp._col_ghost_owners = maps[1].owners.copy()
        sp._finalized = True
        return sp

    # -- insertion -------------------------------------------------------------------------------
    def _check_open(self):
        if self._finalized:
            raise RuntimeError("Cannot insert into sparsity pattern. It has already been finalized")

    def insert(self, rows, cols):
        """SparsityPattern::insert(rows, cols): dense block rows x cols (SparsityPattern.cpp:207-227)."""
        self._check_open()
        rows = np.atleast_1d(np.asarray(rows, dtype=np.int32))
        cols = np.atleast_1d(np.asarray(cols, dtype=np.int32))
        self._segments.append(("coo", np.repeat(rows, cols.size), np.tile(cols, rows.size)))

    def insert_diagonal(self, rows):
        """SparsityPattern.cpp:229-240"""
        self._check_open()
        rows = np.atleast_1d(np.asarray(rows, dtype=np.int32))
        self._segments.append(("coo", rows.copy(), rows.copy()))

    def insert_cells(self, cells, dofmap0, dofmap1):
        """fem::sparsitybuild::cells: for every cell insert(dofmap0[cell], dofmap1[cell]).

        ``cells`` None means all rows of the dofmaps.  numpy inputs are expanded on the host
        (reference behaviour); torch CUDA inputs are kept for the native device build.
        """
        self._check_open()
        if isinstance(dofmap0, np.ndarray):
            d0 = dofmap0 if cells is None else dofmap0[np.asarray(cells)]
            d1 = dofmap1 if cells is None else dofmap1[np.asarray(cells)]
            nd0, nd1 = d0.shape[1], d1.shape[1]
            rows = np.repeat(d0[:, :, None], nd1, axis=2).reshape(-1).astype(np.int32)
            cols = np.repeat(d1[:, None, :], nd0, axis=1).reshape(-1).astype(np.int32)
            self._segments.append(("coo", rows, cols))
        else:
            self._segments.append(("cells", cells, dofmap0, dofmap1))

    # -- finalize --------------------------------------------------------------------------------
    def _ghost_row_entries(self, local_size0, num_rows0, ncols):
        """(ghost row index, local column) of every distinct entry inserted into a ghost row, grouped by row, the
        columns of a row in the order of their first insertion."""
        rows_all, cols_all = [], []
        for seg in self._segments:
            if seg[0] == "coo":
                _, rows, cols = seg
                sel = np.flatnonzero(rows >= local_size0)
                rows_all.append(rows[sel].astype(np.int64) - local_size0)
                cols_all.append(cols[sel].astype(np.int64))
            else:
                from . import _lib

                _, cells, dm0, dm1 = seg
                n_ghost = num_rows0 - local_size0
                ncells = dm0.shape[0] if cells is None else cells.numel()
                counts = np.zeros(n_ghost, dtype=np.int64)
                args = (num_rows0, local_size0, _lib.dptr(dm0), dm0.shape[1], _lib.dptr(dm1), dm1.shape[1],
                        _lib.dptr(cells), ncells)
                _lib.check(_lib.lib.bfx_sparsity_ghost_rows(*args, counts.ctypes.data, None, _lib.current_stream()))
                packed = np.zeros(int(counts.sum()), dtype=np.int32)
                _lib.check(_lib.lib.bfx_sparsity_ghost_rows(*args, counts.ctypes.data, packed.ctypes.data, _lib.current_stream()))
                rows_all.append(np.repeat(np.arange(n_ghost, dtype=np.int64), counts))
                cols_all.append(packed.astype(np.int64))
        if not rows_all:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
        return first_occurrences_by_row(np.concatenate(rows_all), np.concatenate(cols_all), ncols)

    def finalize(self):
        if self._finalized:
            raise RuntimeError("Sparsity pattern has already been finalised.")
        m0, m1 = self._index_maps
        local_size0, local_size1 = m0.size_local, m1.size_local
        num_rows0 = local_size0 + m0.num_ghosts
        col_ghosts = list(m1.ghosts)
        col_ghost_owners = list(m1.owners)
        recv_rows = np.zeros(0, dtype=np.int32)
        recv_cols = np.zeros(0, dtype=np.int32)

        if self.comm.size > 1:
            # (global row, global col, col owner) triplets of ghost rows -> owners (:291-383)
            grow_i, gcol_l = self._ghost_row_entries(local_size0, num_rows0, local_size1 + m1.num_ghosts)
            neighbour_rank = np.searchsorted(m0.src, m0.owners)
            owned = gcol_l < local_size1
            gi = np.where(owned, 0, gcol_l - local_size1)
            has_g = m1.num_ghosts > 0
            gcol = np.where(owned, gcol_l + m1.local_range[0], m1.ghosts[gi] if has_g else 0)
            gown = np.where(owned, self.comm.rank, m1.owners[gi].astype(np.int64) if has_g else 0)
            trip = np.stack([m0.ghosts[grow_i] if grow_i.size else np.zeros(0, dtype=np.int64), gcol, gown], axis=1)
            nb = neighbour_rank[grow_i] if grow_i.size else np.zeros(0, dtype=np.int64)
            order = np.argsort(nb, kind="stable")  # per owner: ghost rows ascending, columns in insertion order
            cnt = np.bincount(nb, minlength=len(m0.src)) if grow_i.size else np.zeros(len(m0.src), dtype=np.int64)
            disp = np.concatenate([[0], np.cumsum(cnt)])
            trip = trip[order].astype(np.int64)
            send = [trip[disp[k]:disp[k + 1]].reshape(-1) for k in range(len(m0.src))]
            recv = self.comm.neighbor_alltoallv(m0.src, m0.dest, send, dtype=np.int64)
            data_in = np.concatenate(recv) if recv else np.zeros(0, dtype=np.int64)
            # new ghost columns in arrival order (:389-423)
            in_rows, in_cols, in_own = data_in[0::3], data_in[1::3], data_in[2::3]
            recv_rows = (in_rows - m0.local_range[0]).astype(np.int32)
            recv_cols, col_ghosts, col_ghost_owners = received_columns_to_local(
                in_cols, in_own, m1.local_range, col_ghosts, col_ghost_owners)

        cell_segs = [s for s in self._segments if s[0] == "cells"]
        coo_rows = [s[1] for s in self._segments if s[0] == "coo"] + [recv_rows]
        coo_cols = [s[2] for s in self._segments if s[0] == "coo"] + [recv_cols]
        coo_rows = np.concatenate(coo_rows).astype(np.int32)
        coo_cols = np.concatenate(coo_cols).astype(np.int32)
        if len(cell_segs) == 1:
            # native device build: per-row dedup + sort (:438-478)
            from . import _lib

            _, cells, dm0, dm1 = cell_segs[0]
            ncells = dm0.shape[0] if cells is None else cells.numel()
            h = C.c_void_p()
            _lib.check(
                _lib.lib.bfx_sparsity_build(
                    C.byref(h), num_rows0, local_size0, local_size1, _lib.dptr(dm0), dm0.shape[1], _lib.dptr(dm1),
                    dm1.shape[1], _lib.dptr(cells), ncells, coo_rows.ctypes.data if coo_rows.size else None,
                    coo_cols.ctypes.data if coo_cols.size else None, coo_rows.size, self._bs[0], self._bs[1],
                    _lib.current_stream(),
                )
            )
            self._csr = h
            release_with(self, "bfx_csr_destroy", h)  # matrices on this pattern keep the pattern alive
        elif len(cell_segs) > 1:
            raise NotImplementedError("at most one device cell list per SparsityPattern")
        else:
            if coo_rows.size and (coo_rows.min() < 0 or coo_rows.max() >= num_rows0):
                raise RuntimeError("SparsityPattern: row index out of range")
            ncols = local_size1 + len(col_ghosts)
            keys = np.unique(coo_rows.astype(np.int64) * ncols + coo_cols)
            rows_u = keys // ncols
            self._edges = (keys % ncols).astype(np.int32)
            offsets = np.zeros(num_rows0 + 1, dtype=np.int64)
            np.add.at(offsets, rows_u + 1, 1)
            self._offsets = np.cumsum(offsets)
            diag = np.zeros(num_rows0, dtype=np.int64)
            np.add.at(diag, rows_u[self._edges < local_size1], 1)
            self._off_diag = diag.astype(np.int32)

        self._segments = []
        self._col_ghosts = np.asarray(col_ghosts, dtype=np.int64)
        self._col_ghost_owners = np.asarray(col_ghost_owners, dtype=np.int32)
        # new column IndexMap (:488-490)
        if self.comm.size > 1 or len(col_ghosts) != m1.num_ghosts:
            self._index_maps[1] = IndexMap(self.comm, m1.size_local, self._col_ghosts, self._col_ghost_owners)
        self._finalized = True

    # -- accessors -------------------------------------------------------------------------------
    def _need_final(self):
        if not self._finalized:
            raise RuntimeError("Sparsity pattern has not been finalised.")

    def _download(self):
        if self._edges is None:
            from . import _lib

            n = self._index_maps[0].size_local + self._index_maps[0].num_ghosts
            nnz = int(_lib.lib.bfx_csr_nnz(self._csr))
            self._offsets = np.empty(n + 1, dtype=np.int64)
            self._edges = np.empty(nnz, dtype=np.int32)
            od = np.empty(n, dtype=np.int64)
            _lib.check(_lib.lib.bfx_csr_get_structure(self._csr, self._offsets.ctypes.data, self._edges.ctypes.data, od.ctypes.data))
            self._off_diag = (od - self._offsets[:-1]).astype(np.int32)

    def index_map(self, dim):
        return self._index_maps[dim]

    def block_size(self, dim):
        return self._bs[dim]

    @property
    def num_nonzeros(self):
        self._need_final()
        if self._edges is None:
            from . import _lib

            return int(_lib.lib.bfx_csr_nnz(self._csr))
        return int(self._edges.size)

    @property
    def graph(self):
        """(edges int32, offsets int64) — SparsityPattern::graph (SparsityPattern.cpp:514-520)."""
        self._need_final()
        self._download()
        return self._edges, self._offsets

    @property
    def off_diagonal_offsets(self):
        self._need_final()
        self._download()
        return self._off_diag

    def nnz_diag(self, row):
        return int(self.off_diagonal_offsets[row])

    def nnz_off_diag(self, row):
        e, o = self.graph
        return int(o[row + 1] - o[row] - self._off_diag[row])

    def column_indices(self):
        """SparsityPattern.cpp:247-259"""
        self._need_final()
        m1 = self._index_maps[1]
        return np.concatenate([np.arange(m1.local_range[0], m1.local_range[1], dtype=np.int64), self._col_ghosts])


# ---------------------------------------------------------------------------------------------
# Vector
# ---------------------------------------------------------------------------------------------
class Vector:
    """la::Vector<double> with device storage (la/Vector.h:47-422).

    ``array`` is a torch CUDA tensor [owned*bs | ghosts*bs]; scatter_forward / scatter_reverse run the
    pack kernel, the NCCL exchange and the unpack kernel of libbfx (Vector.h:219-379).
    """

    def __init__(self, index_map: IndexMap, bs: int = 1, array=None):
        torch = _torch()
        self.index_map, self.bs = index_map, int(bs)
        n = self.bs * (index_map.size_local + index_map.num_ghosts)
        self.array = torch.zeros(n, dtype=torch.float64, device=_device()) if array is None else array
        assert self.array.numel() == n
        self._scatterer = cached_scatterer(index_map, self.bs)
        self._dev_plan = None  # this vector's own device buffers / stream / events, created by the first exchange

    def _plan(self):
        if self._dev_plan is None:
            self._dev_plan = self._scatterer.new_device_plan()
            release_with(self, "bfx_scatter_destroy", self._dev_plan)
        return self._dev_plan

    @property
    def block_size(self):
        return self.bs

    @property
    def scatterer(self) -> Scatterer:
        return self._scatterer

    def set(self, v: float):
        from . import _lib

        _lib.check(_lib.lib.bfx_fill(self.array.numel(), float(v), self.array.data_ptr(), _lib.current_stream()))

    def _n_owned(self):
        return self.bs * self.index_map.size_local

    def scatter_fwd_begin(self):
        from . import _lib

        if self.index_map.comm.size > 1:
            _lib.check(_lib.lib.bfx_scatter_fwd_begin(self._plan(), self.array.data_ptr(), _lib.current_stream()))

    def scatter_fwd_end(self):
        from . import _lib

        if self.index_map.comm.size > 1:
            _lib.check(_lib.lib.bfx_scatter_fwd_end(self._plan(), self.array.data_ptr(), self._n_owned(), _lib.current_stream()))

    def scatter_forward(self):
        """Vector::scatter_fwd — owner values to ghosts."""
        self.scatter_fwd_begin()
        self.scatter_fwd_end()

    def scatter_rev_begin(self):
        from . import _lib

        if self.index_map.comm.size > 1:
            _lib.check(_lib.lib.bfx_scatter_rev_begin(self._plan(), self.array.data_ptr(), self._n_owned(), _lib.current_stream()))

    def scatter_rev_end(self, mode: InsertMode = InsertMode.add):
        from . import _lib

        if self.index_map.comm.size > 1:
            op = 1 if mode == InsertMode.add else 0
            _lib.check(_lib.lib.bfx_scatter_rev_end(self._plan(), self.array.data_ptr(), op, _lib.current_stream()))

    def scatter_reverse(self, mode: InsertMode = InsertMode.add):
        """Vector::scatter_rev(op) — ghost values to owners (add or insert)."""
        self.scatter_rev_begin()
        self.scatter_rev_end(mode)


def inner_product(a: Vector, b: Vector) -> float:
    """la::inner_product (la/Vector.h:434-460): owned entries, summed over ranks."""
    from . import _lib

    n = a.bs * a.index_map.size_local
    if n != b.bs * b.index_map.size_local:
        raise RuntimeError("Incompatible vector sizes")
    out = C.c_double(0.0)
    _lib.check(_lib.lib.bfx_dot(n, a.array.data_ptr(), b.array.data_ptr(), C.byref(out), _lib.current_stream()))
    return a.index_map.comm.allreduce_sum(out.value)


def squared_norm(a: Vector) -> float:
    return inner_product(a, a)


def norm(x: Vector, type: Norm = Norm.l2) -> float:
    """la::norm (la/Vector.h:479-514)."""
    from . import _lib

    n = x.bs * x.index_map.size_local
    out = C.c_double(0.0)
    if type == Norm.l2:
        return float(np.sqrt(squared_norm(x)))
    if type == Norm.l1:
        _lib.check(_lib.lib.bfx_norm(n, x.array.data_ptr(), 0, C.byref(out), _lib.current_stream()))
        return x.index_map.comm.allreduce_sum(out.value)
    if type == Norm.linf:
        _lib.check(_lib.lib.bfx_norm(n, x.array.data_ptr(), 2, C.byref(out), _lib.current_stream()))
        return x.index_map.comm.allreduce_max(out.value)
    raise RuntimeError("Norm type not supported")


def axpy(r: Vector, alpha: float, x: Vector, y: Vector):
    """r = alpha x + y on the whole local array (cpp/demo/poisson_matrix_free/main.cpp:60-69)."""
    from . import _lib

    if r is y:
        _lib.check(_lib.lib.bfx_axpy(r.array.numel(), float(alpha), x.array.data_ptr(), r.array.data_ptr(), _lib.current_stream()))
    elif r is x:
        r.array.mul_(float(alpha)).add_(y.array)
    else:
        r.array.copy_(y.array)
        _lib.check(_lib.lib.bfx_axpy(r.array.numel(), float(alpha), x.array.data_ptr(), r.array.data_ptr(), _lib.current_stream()))


def cg(x: Vector, b: Vector, action, kmax: int = 50, rtol: float = 1e-8) -> int:
    """Conjugate gradients with an operator given by its action, everything on the device
    (linalg::cg of cpp/demo/poisson_matrix_free/main.cpp:84-132, same update order).
    ``action(p, y)`` computes y = A p; ghost values of x and b must be up to date.  Returns the iterations."""
    torch = _torch()
    r = Vector(b.index_map, b.bs, torch.empty_like(b.array))
    y = Vector(b.index_map, b.bs, torch.empty_like(b.array))
    action(x, y)
    axpy(r, -1.0, y, b)
    p = Vector(b.index_map, b.bs, r.array.clone())
    rnorm0 = squared_norm(r)
    rnorm = rnorm0
    if rnorm0 == 0.0:
        return 0
    k = 0
    while k < kmax:
        k += 1
        action(p, y)
        alpha = rnorm / inner_product(p, y)
        axpy(x, alpha, p, x)
        axpy(r, -alpha, y, r)
        rnorm_new = squared_norm(r)
        beta = rnorm_new / rnorm
        rnorm = rnorm_new
        if rnorm / rnorm0 < rtol * rtol:
            break
        axpy(p, beta, p, r)
    return k


# ---------------------------------------------------------------------------------------------
# MatrixCSR
# ---------------------------------------------------------------------------------------------
def received_columns_to_local(in_cols, in_own, local_range, col_ghosts, col_ghost_owners):
    """Local column index of every received (global column, owner) pair; columns this rank neither owns nor ghosts are
    appended to the ghost list in ARRIVAL order (la/SparsityPattern.cpp:389-423).
    Returns (local columns int32, ghosts list, owners list)."""
    in_cols = np.asarray(in_cols, dtype=np.int64)
    in_own = np.asarray(in_own, dtype=np.int64)
    ls1 = local_range[1] - local_range[0]
    mine = (in_cols >= local_range[0]) & (in_cols < local_range[1])
    known = np.asarray(col_ghosts, dtype=np.int64)
    cand = np.flatnonzero(~mine & ~np.isin(in_cols, known))
    col_ghosts, col_ghost_owners = list(col_ghosts), list(col_ghost_owners)
    if cand.size:
        _, first = np.unique(in_cols[cand], return_index=True)
        newg = cand[np.sort(first)]  # first arrival of every unknown column
        col_ghosts += in_cols[newg].tolist()
        col_ghost_owners += in_own[newg].tolist()
    allg = np.asarray(col_ghosts, dtype=np.int64)
    out = in_cols - local_range[0]
    if np.any(~mine):
        og = np.argsort(allg, kind="stable")
        out[~mine] = ls1 + og[np.searchsorted(allg[og], in_cols[~mine])]
    return out.astype(np.int32), col_ghosts, col_ghost_owners


def locate_entries(indptr, indices, rows, cols, ncols):
    """Position in the CSR arrays of every (rows[k], cols[k]) - the ``std::lower_bound`` per entry of
    la/MatrixCSR.h:829-845, for all entries at once: inside the rows that occur, (row rank, column) is a strictly
    increasing key, so one ``searchsorted`` over the entries of those rows does it.  None if an entry is missing."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    if rows.size == 0:
        return np.zeros(0, dtype=np.int64)
    indptr = np.asarray(indptr)
    urows, inv = np.unique(rows, return_inverse=True)
    start = indptr[urows].astype(np.int64)
    lens = indptr[urows + 1].astype(np.int64) - start
    tot = int(lens.sum())
    first = np.concatenate([[0], np.cumsum(lens)[:-1]])
    flat = np.arange(tot, dtype=np.int64) + np.repeat(start - first, lens)  # positions of the entries of those rows
    rank_of = np.repeat(np.arange(urows.size, dtype=np.int64), lens)
    keys = rank_of * np.int64(ncols) + np.asarray(indices)[flat].astype(np.int64)
    want = inv.astype(np.int64) * np.int64(ncols) + cols
    at = np.searchsorted(keys, want)
    if tot == 0 or np.any(at >= tot) or np.any(keys[np.minimum(at, tot - 1)] != want):
        return None
    return flat[at]


def first_occurrences_by_row(rows, cols, ncols):
    """(rows, cols) pairs in insertion order -> the pairs grouped by row, every row keeping its DISTINCT columns in
    the order of their first insertion (what the COO cache of la/SparsityPattern.cpp:291-330 hands to the owners).
    Returns (rows_sorted, cols_sorted)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    if rows.size == 0:
        return rows, cols
    _, first = np.unique(rows * np.int64(ncols) + cols, return_index=True)
    keep = np.sort(first)  # first occurrences, still in insertion order
    order = np.argsort(rows[keep], kind="stable")
    return rows[keep][order], cols[keep][order]


def matrix_ghost_plan(index_maps, bs, indptr, indices):
    """Ghost-row exchange plan built by the MatrixCSR constructor (la/MatrixCSR.h:705-849), host side.

    Returns the reference's members: ``ghost_row_to_rank`` (index into src), ``val_send_disp`` /
    ``val_recv_disp`` (already x bs0*bs1), ``unpack_pos`` (block positions in the owner's CSR)."""
    m0, m1 = index_maps
    comm = m0.comm
    bs2 = bs[0] * bs[1]
    ls0, ls1 = m0.size_local, m1.size_local
    ghost_row_to_rank = np.searchsorted(m0.src, m0.owners).astype(np.int32)  # :725-733
    row_len = np.diff(indptr[ls0:])
    data_per_proc = np.zeros(len(m0.src), dtype=np.int64)
    np.add.at(data_per_proc, ghost_row_to_rank, row_len)  # :735-742
    val_send_disp = np.concatenate([[0], np.cumsum(data_per_proc)]).astype(np.int64)
    # (global row, global col) of every ghost-row entry, grouped by owner in ghost-row order (:750-775)
    grows = np.repeat(np.arange(m0.num_ghosts), row_len)
    gcols_local = indices[indptr[ls0]:]
    gcol = m1.local_to_global(gcols_local)
    grow = m0.ghosts[grows]
    order = np.argsort(ghost_row_to_rank[grows], kind="stable")
    pairs = np.stack([grow[order], gcol[order]], axis=1).reshape(-1)
    send = [pairs[2 * val_send_disp[i]:2 * val_send_disp[i + 1]] for i in range(len(m0.src))]
    recv = comm.neighbor_alltoallv(m0.src, m0.dest, send, dtype=np.int64)  # :777-800
    recv_disp = np.concatenate([[0], np.cumsum([len(a) for a in recv])]).astype(np.int64)
    arr = np.concatenate(recv) if recv else np.zeros(0, dtype=np.int64)
    # positions in the owner's CSR (:813-846)
    lrow = (arr[0::2] - m0.local_range[0]).astype(np.int64)
    if lrow.size and (lrow.min() < 0 or lrow.max() >= ls0):
        raise RuntimeError("MatrixCSR: received a ghost row this rank does not own")
    gc = arr[1::2]
    lcol = gc - m1.local_range[0]
    ghost_mask = (lcol < 0) | (lcol >= ls1)
    if np.any(ghost_mask):
        order1 = np.argsort(m1.ghosts, kind="stable")
        pos = np.searchsorted(m1.ghosts[order1], gc[ghost_mask])
        lcol[ghost_mask] = ls1 + order1[pos]
    unpack = locate_entries(indptr, indices, lrow, lcol, ls1 + m1.num_ghosts)
    if unpack is None:
        raise RuntimeError("MatrixCSR: received ghost-row entry not in sparsity")
    return dict(ghost_row_to_rank=ghost_row_to_rank, val_send_disp=val_send_disp * bs2,
                val_recv_disp=bs2 * recv_disp // 2, unpack_pos=unpack, src=m0.src.copy(), dest=m0.dest.copy())



def matrix_matmul_plan(m1A: IndexMap, m0B: IndexMap, m1B: IndexMap, indptrB, indicesB, gather_values_B):
    """Host side of la::matmul: impl::fetch_ghost_rows (la/matmul.h:79-390), numpy + the neighbourhood exchanges.

    Every ghost column of A (map ``m1A``) is a row of B on its owner.  The owners send those rows (global columns, the
    owner of every column, values - ``gather_values_B(ks)`` returns the values of the entries ``ks`` of B); the column
    map of C is B's column map extended by the sorted unique (global column, owner) pairs received (:330-356), or B's
    column map itself when the column map of A has no neighbours on this rank (:104-113).
    Returns dict(col_map, ghost_row_ptr int64, ghost_cols int32 (local in col_map), ghost_vals, b_ghost_remap int32
    (B's ghost columns in col_map, la/matmul.h:419-422)); the rows are in the order of A's ghost columns."""
    comm = m1A.comm
    if comm.size == 1 or (len(m1A.src) == 0 and len(m1A.dest) == 0):
        cm = IndexMap(comm, m1B.size_local, m1B.ghosts, m1B.owners)
        return {"col_map": cm, "ghost_row_ptr": np.zeros(1, dtype=np.int64), "ghost_cols": np.zeros(0, dtype=np.int32),
                "ghost_vals": np.zeros(0), "b_ghost_remap": (m1B.size_local + np.arange(m1B.num_ghosts)).astype(np.int32)}
    indptrB = np.asarray(indptrB, dtype=np.int64)
    # the rows we need, grouped by owner in ghost order; perm[i] = position of ghost i in that order (:150-160)
    nbr = np.searchsorted(m1A.src, m1A.owners)
    order = np.argsort(nbr, kind="stable")
    perm = np.empty(len(order), dtype=np.int64)
    perm[order] = np.arange(len(order))
    cnt = np.bincount(nbr, minlength=len(m1A.src))
    disp = np.concatenate([[0], np.cumsum(cnt)])
    required = m1A.ghosts[order]
    asked = comm.neighbor_alltoallv(m1A.src, m1A.dest, [required[disp[i]:disp[i + 1]] for i in range(len(m1A.src))])
    # the rows asked of this rank: entries, global columns, owners of the columns, row sizes (:180-262)
    send_cols, send_own, send_vals, send_sizes = [], [], [], []
    for rows_g in asked:
        lr = (np.asarray(rows_g, dtype=np.int64) - m0B.local_range[0]).astype(np.int64)
        assert np.all((lr >= 0) & (lr < m0B.size_local)), "matmul: asked for a row this rank does not own"
        lens = indptrB[lr + 1] - indptrB[lr]
        first = np.concatenate([[0], np.cumsum(lens)[:-1]]) if lens.size else np.zeros(0, dtype=np.int64)
        ks = np.arange(int(lens.sum()), dtype=np.int64) + np.repeat(indptrB[lr] - first, lens)
        lc = np.asarray(indicesB)[ks].astype(np.int64)
        send_cols.append(m1B.local_to_global(lc) if lc.size else np.zeros(0, dtype=np.int64))
        own = np.full(lc.size, comm.rank, dtype=np.int64)
        gh = lc >= m1B.size_local
        if np.any(gh):
            own[gh] = m1B.owners[lc[gh] - m1B.size_local]
        send_own.append(own)
        send_vals.append(np.asarray(gather_values_B(ks), dtype=np.float64) if ks.size else np.zeros(0))
        send_sizes.append(lens.astype(np.int64))
    rc = comm.neighbor_alltoallv(m1A.dest, m1A.src, send_cols)
    ro = comm.neighbor_alltoallv(m1A.dest, m1A.src, send_own)
    rv = comm.neighbor_alltoallv(m1A.dest, m1A.src, send_vals, dtype=np.float64)
    rs = comm.neighbor_alltoallv(m1A.dest, m1A.src, send_sizes)
    rc = np.concatenate(rc) if len(rc) else np.zeros(0, dtype=np.int64)
    ro = np.concatenate(ro) if len(ro) else np.zeros(0, dtype=np.int64)
    rv = np.concatenate(rv) if len(rv) else np.zeros(0)
    rs = np.concatenate(rs) if len(rs) else np.zeros(0, dtype=np.int64)
    # column map of C (:318-356)
    lo, hi = m1B.local_range
    far = (rc < lo) | (rc >= hi)
    pairs = np.stack([np.concatenate([rc[far], m1B.ghosts]), np.concatenate([ro[far], m1B.owners.astype(np.int64)])], axis=1)
    pairs = np.unique(pairs, axis=0) if len(pairs) else np.zeros((0, 2), dtype=np.int64)
    cm = IndexMap(comm, m1B.size_local, pairs[:, 0], pairs[:, 1].astype(np.int32))

    def to_local(g):
        g = np.asarray(g, dtype=np.int64)
        out = g - lo
        f = (g < lo) | (g >= hi)
        if np.any(f):
            out[f] = m1B.size_local + np.searchsorted(pairs[:, 0], g[f])
        return out.astype(np.int32)

    cols_l = to_local(rc)
    # back to the order of A's ghost columns (:364-388)
    rp_recv = np.concatenate([[0], np.cumsum(rs)]).astype(np.int64)
    sizes = rs[perm]
    rp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    src_pos = np.arange(int(sizes.sum()), dtype=np.int64) + np.repeat(rp_recv[perm] - rp[:-1], sizes)
    return {"col_map": cm, "ghost_row_ptr": rp, "ghost_cols": cols_l[src_pos], "ghost_vals": rv[src_pos],
            "b_ghost_remap": to_local(m1B.ghosts)}


def matrix_transpose_plan(m0: IndexMap, m1: IndexMap, bs, indptr, indices, off_diag_offset, rp0, c0, gather_values):
    """Host side of the distributed la::transpose (la/mattrans.h:200-434), numpy + the neighbourhood exchange.

    ``rp0`` / ``c0``: row pointer and columns of the locally transposed block (impl::local_transpose);
    ``gather_values(ks)``: the blocks of the entries ``ks`` of A, flattened (a device gather in the product).
    Returns the maps [row map, column map] of the result, its structure, and where the local and the received blocks go
    (``local_dst``: position of every entry of the local transpose, ``recv_dst`` / ``recv_blocks``: position and
    transposed block of every received entry)."""
    comm = m0.comm
    bs0, bs1 = bs
    nbs = bs0 * bs1
    n_row, n_col = m0.size_local, m1.size_local
    n0 = int(rp0[-1])
    # entries in ghost columns, grouped by the owner of the column, rows ascending (:215-262)
    od = np.asarray(off_diag_offset[:n_row], dtype=np.int64)
    lens = np.asarray(indptr[1 : n_row + 1], dtype=np.int64) - od
    tot = int(lens.sum())
    starts = np.repeat(od - np.concatenate([[0], np.cumsum(lens)[:-1]]), lens)
    ks = (np.arange(tot, dtype=np.int64) + starts).astype(np.int64)
    rows_g = np.repeat(np.arange(n_row, dtype=np.int64) + m0.local_range[0], lens)
    jg = np.asarray(indices)[ks].astype(np.int64) - n_col
    nbr = np.searchsorted(m1.src, m1.owners)[jg] if tot else np.zeros(0, dtype=np.int64)
    order = np.argsort(nbr, kind="stable")
    counts = np.bincount(nbr, minlength=len(m1.src)) if tot else np.zeros(len(m1.src), dtype=np.int64)
    disp = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    send_rows = rows_g[order]
    send_cols = m1.ghosts[jg[order]] if tot else np.zeros(0, dtype=np.int64)
    send_vals = np.asarray(gather_values(ks[order]), dtype=np.float64) if tot else np.zeros(0)

    def split(a, w=1):
        return [a[disp[i] * w : disp[i + 1] * w] for i in range(len(m1.src))]

    # this rank sends to src[] (owners of its ghost columns) and receives from dest[] (:200-206, :290-300)
    recv_rows = comm.neighbor_alltoallv(m1.src, m1.dest, split(send_rows), dtype=np.int64)
    recv_cols = comm.neighbor_alltoallv(m1.src, m1.dest, split(send_cols), dtype=np.int64)
    recv_vals = comm.neighbor_alltoallv(m1.src, m1.dest, split(send_vals, nbs), dtype=np.float64)
    rr = np.concatenate(recv_rows) if len(recv_rows) else np.zeros(0, dtype=np.int64)
    rc = np.concatenate(recv_cols) if len(recv_cols) else np.zeros(0, dtype=np.int64)
    rv = np.concatenate(recv_vals) if len(recv_vals) else np.zeros(0)
    # column map of the result: owned = rows of A, ghosts = sorted unique (global row, sender) pairs (:335-364)
    senders = (np.repeat(np.asarray(m1.dest, dtype=np.int64), [len(a) for a in recv_rows]) if len(rr)
               else np.zeros(0, dtype=np.int64))
    pairs = np.unique(np.stack([rr, senders], axis=1), axis=0) if len(rr) else np.zeros((0, 2), dtype=np.int64)
    at_col_map = IndexMap(comm, n_row, pairs[:, 0], pairs[:, 1].astype(np.int32))
    at_row_map = IndexMap(comm, n_col)
    new_col_local = (n_row + np.searchsorted(pairs[:, 0], rr)).astype(np.int32)  # (a global row has one owner)
    # merge: the local rows first, the received entries appended in arrival order (:314-424)
    lc = (rc - m1.local_range[0]).astype(np.int64)
    assert np.all((lc >= 0) & (lc < n_col))
    off = np.diff(rp0).astype(np.int64)
    per_row = np.bincount(lc, minlength=n_col).astype(np.int64) if len(lc) else np.zeros(n_col, dtype=np.int64)
    row_ptr = np.concatenate([[0], np.cumsum(off + per_row)]).astype(np.int64)
    cols = np.zeros(int(row_ptr[-1]), dtype=np.int32)
    local_dst = np.arange(n0, dtype=np.int64) + np.repeat(row_ptr[:-1] - np.asarray(rp0[:-1], dtype=np.int64), off)
    cols[local_dst] = c0
    recv_dst = np.zeros(0, dtype=np.int64)
    if len(lc):
        o2 = np.argsort(lc, kind="stable")  # arrival rank of an entry inside its row: the reference's cursor
        first = np.concatenate([[0], np.cumsum(per_row)[:-1]])
        rank_in_row = np.empty(len(lc), dtype=np.int64)
        rank_in_row[o2] = np.arange(len(lc)) - first[lc[o2]]
        recv_dst = row_ptr[lc] + off[lc] + rank_in_row
        cols[recv_dst] = new_col_local
    blocks = np.ascontiguousarray(rv.reshape(-1, bs0, bs1).transpose(0, 2, 1)).reshape(-1, nbs)
    return {"maps": [at_row_map, at_col_map], "row_ptr": row_ptr, "cols": cols, "off_diag": off, "local_dst": local_dst,
            "recv_dst": recv_dst, "recv_blocks": blocks}


_MATRIX_SERIAL = itertools.count(1)


class MatrixCSR:
    """la::MatrixCSR<double> with device storage (la/MatrixCSR.h:67-624).

    ``data`` (torch CUDA, nnz*bs0*bs1), ``indices`` int32, ``indptr`` int64 keep the reference's
    names and widths (python/dolfinx/la/__init__.py).  The constructor builds the ghost-row
    exchange plan of la/MatrixCSR.h:705-849.
    """

    def __init__(self, pattern: SparsityPattern, block_mode: BlockMode = BlockMode.compact):
        from . import _lib

        pattern._need_final()
        torch = _torch()
        self._pattern = pattern
        self._block_mode = block_mode
        self._index_maps = [pattern.index_map(0), pattern.index_map(1)]
        self._bs = (pattern.block_size(0), pattern.block_size(1))
        self._expanded = None
        if block_mode == BlockMode.expanded and self._bs != (1, 1):
            # la/MatrixCSR.h:638-694: index maps with every block index unrolled, every block row repeated bs0 times
            # with its columns unrolled by bs1, block sizes 1
            bs0, bs1 = self._bs
            maps = []
            for im, b in zip(self._index_maps, self._bs):
                ghosts = (im.ghosts[:, None] * b + np.arange(b)[None, :]).reshape(-1)
                maps.append(IndexMap(im.comm, im.size_local * b, ghosts, np.repeat(im.owners, b), (im.src, im.dest)))
            edges, offsets = pattern.graph
            lens = np.diff(offsets)
            new_lens = np.repeat(lens * bs1, bs0)
            new_ptr = np.concatenate([[0], np.cumsum(new_lens)]).astype(np.int64)
            unrolled = (edges.astype(np.int64)[:, None] * bs1 + np.arange(bs1)[None, :]).reshape(-1)  # per block row
            blk_ptr = offsets.astype(np.int64) * bs1
            pieces = [unrolled[blk_ptr[i]:blk_ptr[i + 1]] for i in range(len(lens)) for _ in range(bs0)]
            new_cols = (np.concatenate(pieces) if pieces else np.zeros(0)).astype(np.int32)
            new_od = new_ptr[:-1] + np.repeat(pattern.off_diagonal_offsets.astype(np.int64) * bs1, bs0)
            self._index_maps = maps
            self._bs = (1, 1)
            self._expanded = (new_cols, new_ptr, new_od)
        m0 = self._index_maps[0]
        n_all = m0.size_local + m0.num_ghosts
        if self._expanded is not None:
            cols_e, ptr_e, od_e = self._expanded
            h = C.c_void_p()
            _lib.check(_lib.lib.bfx_csr_create(C.byref(h), n_all, m0.size_local, ptr_e.ctypes.data,
                                               np.ascontiguousarray(cols_e).ctypes.data, od_e.ctypes.data, 1, 1))
            self._csr = h
            release_with(self, "bfx_csr_destroy", h)
        elif pattern._csr is None:
            edges, offsets = pattern.graph
            off_diag = (pattern.off_diagonal_offsets.astype(np.int64) + offsets[:-1])  # :695-703
            h = C.c_void_p()
            _lib.check(_lib.lib.bfx_csr_create(C.byref(h), n_all, m0.size_local, offsets.ctypes.data,
                                               np.ascontiguousarray(edges).ctypes.data, off_diag.ctypes.data,
                                               self._bs[0], self._bs[1]))
            self._csr = h
            release_with(self, "bfx_csr_destroy", h)
        else:
            self._csr = pattern._csr
        self._nnz = int(_lib.lib.bfx_csr_nnz(self._csr))
        self._data = torch.zeros(self._nnz * self._bs[0] * self._bs[1], dtype=torch.float64, device=_device())
        self._is_zero = True
        self._zero_pending = False  # set_value(0) not yet written to memory (see set_value)
        self._scatter_plan = None
        # never reused, unlike id(): forms key their cached assembly plans (positions in THIS sparsity) on it
        self._serial = next(_MATRIX_SERIAL)
        self._build_ghost_plan()

    # -- ghost-row plan (la/MatrixCSR.h:705-849) -------------------------------------------------
    def _build_ghost_plan(self):
        self._plan_arrays = None
        if self._index_maps[0].comm.size > 1:
            self._plan_arrays = matrix_ghost_plan(self._index_maps, self._bs, self.indptr, self.indices)

    def _device_scatter_plan(self):
        if self._scatter_plan is None and self._plan_arrays is not None:
            from . import _lib

            p = self._plan_arrays
            m0 = self._index_maps[0]
            h = C.c_void_p()
            _lib.check(
                _lib.lib.bfx_csr_scatter_create(
                    C.byref(h), self._csr, m0.comm.nccl, p["ghost_row_to_rank"].ctypes.data, m0.num_ghosts,
                    np.ascontiguousarray(p["val_send_disp"]).ctypes.data, np.ascontiguousarray(p["src"]).ctypes.data,
                    len(p["src"]), np.ascontiguousarray(p["val_recv_disp"]).ctypes.data,
                    np.ascontiguousarray(p["dest"]).ctypes.data, len(p["dest"]), p["unpack_pos"].ctypes.data,
                )
            )
            self._scatter_plan = h
            release_with(self, "bfx_csr_scatter_destroy", h)
        return self._scatter_plan

    # -- accessors -------------------------------------------------------------------------------
    def index_map(self, i):
        return self._index_maps[i]

    @property
    def block_size(self):
        return list(self._bs)

    @property
    def bs(self):
        return self._bs

    @property
    def indices(self):
        return self._expanded[0] if self._expanded is not None else self._pattern.graph[0]

    @property
    def indptr(self):
        return self._expanded[1] if self._expanded is not None else self._pattern.graph[1]

    @property
    def off_diag_offset(self):
        if self._expanded is not None:
            return self._expanded[2]
        return self._pattern.off_diagonal_offsets.astype(np.int64) + self.indptr[:-1]

    def num_owned_rows(self):
        return self._index_maps[0].size_local

    def num_all_rows(self):
        return self._index_maps[0].size_local + self._index_maps[0].num_ghosts

    # -- value manipulation ------------------------------------------------------------------------
    def _values(self):
        """The value array for the library's own kernels (a pending zero-fill is written first)."""
        if self._zero_pending:
            from . import _lib

            # (the library's own fill: cudaMemsetAsync on the caller's stream, no framework kernel in the step)
            _lib.check(_lib.lib.bfx_fill(self._data.numel(), 0.0, self._data.data_ptr(), _lib.current_stream()))
            self._zero_pending = False
        return self._data

    @property
    def data(self):
        """The value array (python/dolfinx/la/__init__.py ``MatrixCSR.data``).  The caller may write through it, so
        the matrix no longer counts as known-zero: the next assembly accumulates (fem/assembler.h:497-498)."""
        self._is_zero = False
        return self._values()

    def set_value(self, x: float):
        """MatrixCSR::set(value) (la/MatrixCSR.h:239-241).

        ``set_value(0)`` is recorded and written to memory by the first reader of ``data`` - unless an
        assembly kernel that writes every value exactly once (BFX_ASM_ROWGATHER with
        BFX_VALUES_OVERWRITE) gets there first and fuses the zero-fill (``_take_zero_fill``)."""
        if x == 0.0:
            self._zero_pending = True
        else:
            from . import _lib

            self._zero_pending = False
            _lib.check(_lib.lib.bfx_fill(self._data.numel(), float(x), self._data.data_ptr(), _lib.current_stream()))
        self._is_zero = x == 0.0

    def _take_zero_fill(self):
        """For a kernel that overwrites EVERY value: returns the raw array and drops the pending zero-fill."""
        assert self._is_zero
        self._zero_pending = False
        return self._data

    def _insert(self, x, rows, cols, bs, op):
        from . import _lib

        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=np.float64)
        # dispatch of MatrixCSR::add / set (la/MatrixCSR.h:265-335)
        if bs == self._bs[0] and bs == self._bs[1]:
            kind, d0, d1 = 0, self._bs[0], self._bs[1]
        elif self._bs[0] == 1 and self._bs[1] == 1:
            kind, d0, d1 = 1, bs, bs
        elif bs == 1:
            kind, d0, d1 = 2, 1, 1
        else:
            raise RuntimeError("Unsupported block size in MatrixCSR insertion")
        assert x.size == rows.size * cols.size * d0 * d1
        try:
            _lib.check(_lib.lib.bfx_csr_insert(self._csr, self._values().data_ptr(), kind, d0, d1, x.ctypes.data,
                                               rows.ctypes.data, rows.size, cols.ctypes.data, cols.size, op,
                                               _lib.current_stream()))
        except _lib.BfxError as e:
            if e.status == _lib.ERR_NOT_IN_SPARSITY:
                raise RuntimeError("Entry not in sparsity") from e
            raise
        self._is_zero = False

    def add(self, x, rows, cols, bs=1):
        """MatrixCSR::add (la/MatrixCSR.h:310-335)."""
        self._insert(x, rows, cols, bs, 1)

    def set(self, x, rows, cols, bs=1):
        """MatrixCSR::set (la/MatrixCSR.h:265-292)."""
        self._insert(x, rows, cols, bs, 0)

    def scatter_rev_begin(self):
        from . import _lib

        p = self._device_scatter_plan()
        if p is not None:
            _lib.check(_lib.lib.bfx_csr_scatter_rev_begin(p, self._values().data_ptr(), _lib.current_stream()))

    def scatter_rev_end(self):
        from . import _lib

        p = self._device_scatter_plan()
        if p is not None:
            _lib.check(_lib.lib.bfx_csr_scatter_rev_end(p, self._values().data_ptr(), _lib.current_stream()))

    def scatter_reverse(self):
        """MatrixCSR::scatter_rev (la/MatrixCSR.h:384-468)."""
        self.scatter_rev_begin()
        self.scatter_rev_end()

    def squared_norm(self) -> float:
        """MatrixCSR::squared_norm (la/MatrixCSR.h:473-486)."""
        from . import _lib

        out = C.c_double(0.0)
        _lib.check(_lib.lib.bfx_csr_squared_norm(self._csr, self._values().data_ptr(), C.byref(out), _lib.current_stream()))
        return self._index_maps[0].comm.allreduce_sum(out.value)

    def mult(self, x: Vector, y: Vector, transpose: bool = False):
        """MatrixCSR::mult / multT (la/MatrixCSR.h:877-1016): y += A x (or A^T x).

        x must live on the matrix's column map ``index_map(1)``.
        """
        from . import _lib

        st = _lib.current_stream()
        L = _lib.lib
        multi = self._index_maps[0].comm.size > 1
        if not transpose:
            if not multi:
                _lib.check(L.bfx_spmv(self._csr, self._values().data_ptr(), x.array.data_ptr(), y.array.data_ptr(), _lib.SPMV_FULL, st))
                return
            x.scatter_fwd_begin()
            _lib.check(L.bfx_spmv(self._csr, self._values().data_ptr(), x.array.data_ptr(), y.array.data_ptr(), _lib.SPMV_DIAG, st))
            x.scatter_fwd_end()
            _lib.check(L.bfx_spmv(self._csr, self._values().data_ptr(), x.array.data_ptr(), y.array.data_ptr(), _lib.SPMV_OFFDIAG, st))
        else:
            ncl = self._bs[1] * self._index_maps[1].size_local
            y.array[ncl:].zero_()
            _lib.check(L.bfx_spmvT(self._csr, self._values().data_ptr(), x.array.data_ptr(), y.array.data_ptr(), _lib.SPMV_OFFDIAG, st))
            y.scatter_reverse(InsertMode.add)
            _lib.check(L.bfx_spmvT(self._csr, self._values().data_ptr(), x.array.data_ptr(), y.array.data_ptr(), _lib.SPMV_DIAG, st))

    def matmul(self, B: "MatrixCSR") -> "MatrixCSR":
        """la::matmul (la/matmul.h:538-579; python/dolfinx/la/__init__.py:186-205): C = A B, block size 1.

        impl::fetch_ghost_rows is host-side integer work (``matrix_matmul_plan``: the rows of B behind the ghost columns
        of A, the column map of C); impl::matmul runs on the device, one thread per row, bitwise the reference's rows
        (``bfx_csr_matmul_begin/_end``, csrc/matmul_row.h)."""
        from . import _lib

        torch = _torch()
        if (self.index_map(1).size_local != B.index_map(0).size_local
                or self.index_map(1).size_global != B.index_map(0).size_global):
            raise RuntimeError("Invalid matrix sizes for matmul.")
        if tuple(self._bs) != (1, 1) or tuple(B._bs) != (1, 1):
            raise RuntimeError("Block size not supported in matmul.")
        dev = self._values().device
        comm = self._index_maps[0].comm
        plan = matrix_matmul_plan(self.index_map(1), B.index_map(0), B.index_map(1), B.indptr, B.indices,
                                  lambda ks: B._values()[torch.from_numpy(ks).to(dev)].cpu().numpy())
        cm = plan["col_map"]
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt)
        remap = up(np.append(plan["b_ghost_remap"], 0), torch.int32)  # (one spare entry: never an empty allocation)
        grp = up(plan["ghost_row_ptr"], torch.int64)
        gcols = up(np.append(plan["ghost_cols"], 0), torch.int32)
        gvals = up(np.append(plan["ghost_vals"], 0.0), torch.float64)
        h = C.c_void_p()
        nnz = C.c_int64(0)
        st = _lib.current_stream()
        _lib.check(_lib.lib.bfx_csr_matmul_begin(self._csr, self._values().data_ptr(), B._csr, B.data.data_ptr(),
                                                 B.index_map(1).size_local, remap.data_ptr(), grp.data_ptr(), gcols.data_ptr(),
                                                 gvals.data_ptr(), cm.size_local, C.byref(h), C.byref(nnz), st))
        n_rows = self._index_maps[0].size_local
        rp = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
        od = torch.empty(max(n_rows, 1), dtype=torch.int32, device=dev)
        cols = torch.empty(max(int(nnz.value), 1), dtype=torch.int32, device=dev)
        vals = torch.empty(max(int(nnz.value), 1), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib.bfx_csr_matmul_end(h, rp.data_ptr(), od.data_ptr(), cols.data_ptr(), vals.data_ptr(), st))
        n = int(nnz.value)
        sp = SparsityPattern.from_graph(comm, [IndexMap(comm, n_rows), cm], (1, 1), cols[:n].cpu().numpy(), rp.cpu().numpy(),
                                        od[:n_rows].cpu().numpy())
        Cm = MatrixCSR(sp)
        Cm._data.copy_(vals[:n])
        Cm._is_zero = False
        return Cm

    def transpose(self) -> "MatrixCSR":
        """la::transpose (la/mattrans.h:121-437; python/dolfinx/la/__init__.py:207-209).

        The block owned rows x owned columns is transposed on the device (``bfx_csr_transpose_local``, bit-exact
        with impl::local_transpose).  The entries in ghost columns travel to the column owners as (global row,
        global column, value) triplets over the neighbourhood of the column map and are appended to the owner's
        rows in arrival order, like the reference; the result has no ghost rows."""
        from . import _lib

        torch = _torch()
        m0, m1 = self._index_maps
        bs0, bs1 = self._bs
        nbs = bs0 * bs1
        n_row, n_col = m0.size_local, m1.size_local
        comm = m0.comm
        indptr, indices = self.indptr, self.indices
        cap = int(indptr[n_row])
        dev = self._values().device
        rpT = torch.empty(n_col + 1, dtype=torch.int64, device=dev)
        cT = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        vT = torch.empty(max(cap, 1) * nbs, dtype=torch.float64, device=dev)
        nnzT = C.c_int64(0)
        _lib.check(_lib.lib.bfx_csr_transpose_local(self._csr, self._values().data_ptr(), n_col, rpT.data_ptr(), cT.data_ptr(),
                                                    vT.data_ptr(), cap, C.byref(nnzT), _lib.current_stream()))
        n0 = int(nnzT.value)
        rp0 = rpT.cpu().numpy()
        c0 = cT[:n0].cpu().numpy()
        if comm.size == 1 or (len(m1.src) == 0 and len(m1.dest) == 0):
            # no neighbours in the column map: every column is owned (la/mattrans.h:139-159)
            at_col_map = IndexMap(comm, n_row)
            at_row_map = IndexMap(comm, n_col)
            sp = SparsityPattern.from_graph(comm, [at_row_map, at_col_map], (bs1, bs0), c0, rp0, np.diff(rp0))
            AT = MatrixCSR(sp)
            AT._data.copy_(vT[: n0 * nbs])
            AT._is_zero = False
            return AT
        # ---- entries in ghost columns -> column owners; merged structure (host side of la/mattrans.h:200-434)
        plan = matrix_transpose_plan(m0, m1, self._bs, indptr, indices, self.off_diag_offset, rp0, c0,
                                     lambda ks: self._values().view(-1, nbs)[torch.from_numpy(ks).to(dev)].reshape(-1).cpu().numpy())
        sp = SparsityPattern.from_graph(comm, plan["maps"], (bs1, bs0), plan["cols"], plan["row_ptr"], plan["off_diag"])
        AT = MatrixCSR(sp)
        out = AT._data.view(-1, nbs)
        if n0:
            out[torch.from_numpy(plan["local_dst"]).to(dev)] = vT[: n0 * nbs].view(-1, nbs)
        if len(plan["recv_dst"]):
            out[torch.from_numpy(plan["recv_dst"]).to(dev)] = torch.from_numpy(plan["recv_blocks"]).to(dev)
        AT._is_zero = False
        return AT

    def to_dense(self):
        """MatrixCSR::to_dense (la/MatrixCSR.h:343-370), host numpy."""
        bs0, bs1 = self._bs
        nrows = self.num_all_rows()
        ncols = self._index_maps[1].size_local + self._index_maps[1].num_ghosts
        A = np.zeros((nrows * bs0, ncols * bs1))
        data = self._values().cpu().numpy().reshape(-1, bs0, bs1)
        indptr, indices = self.indptr, self.indices
        for r in range(nrows):
            for j in range(indptr[r], indptr[r + 1]):
                c = indices[j]
                A[r * bs0:(r + 1) * bs0, c * bs1:(c + 1) * bs1] = data[j]
        return A

    def to_scipy(self, ghosted=False):
        """python/dolfinx/la/__init__.py:279-319 — the storage *is* SciPy CSR/BSR."""
        import scipy.sparse as sp

        bs0, bs1 = self._bs
        ncols = self._index_maps[1].size_local + self._index_maps[1].num_ghosts
        nrows = self.num_all_rows() if ghosted else self.num_owned_rows()
        nnzlocal = self.indptr[nrows]
        data = self._values().cpu().numpy()
        indptr, indices = self.indptr[: nrows + 1], self.indices[:nnzlocal]
        if bs0 == 1 and bs1 == 1:
            return sp.csr_matrix((data[:nnzlocal], indices, indptr), shape=(nrows, ncols))
        return sp.bsr_matrix((data[: nnzlocal * bs0 * bs1].reshape(-1, bs0, bs1), indices, indptr),
                             shape=(bs0 * nrows, bs1 * ncols))


def matrix_csr(pattern: SparsityPattern, block_mode: BlockMode = BlockMode.compact) -> MatrixCSR:
    return MatrixCSR(pattern, block_mode)


def vector(index_map: IndexMap, bs: int = 1) -> Vector:
    return Vector(index_map, bs)
