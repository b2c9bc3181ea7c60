"""Host mirror of dolfinx::common — Comm, IndexMap, Scatterer (plan construction).

Reference: cpp/dolfinx/common/IndexMap.{h,cpp}, cpp/dolfinx/common/Scatterer.h
(python surface: python/dolfinx/common.py, wrappers/common.cpp).  Plan
construction is host-side integer work, exactly as in the reference; the MPI-3
neighbourhood collectives it uses there are replaced by
``torch.distributed`` point-to-point exchanges (gloo on CPU, NCCL on GPUs).
The data-path exchanges (C1-C4, C6) run on the device through libbfx.so.
"""

from __future__ import annotations

import ctypes as C
import weakref

import numpy as np


class Comm:
    """Communicator: a ``torch.distributed`` process group (or a serial stand-in).

    Replaces ``MPI_Comm`` / ``dolfinx::MPI::Comm`` (common/MPI.h).  ``nccl`` is the
    libbfx communicator used by the device data path; it is created on first use
    from a unique id broadcast through the process group.
    """

    def __init__(self, group=None, serial=False):
        self._serial = serial
        self._group = group
        self._nccl = None
        if serial:
            self.rank, self.size = 0, 1
        else:
            import torch.distributed as dist

            if not dist.is_initialized():
                self.rank, self.size, self._serial = 0, 1, True
            else:
                self.rank = dist.get_rank(group)
                self.size = dist.get_world_size(group)

    # -- host-side collectives used while building plans --------------------------------------
    def _device(self):
        import torch
        import torch.distributed as dist

        return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self._group) == "nccl" else torch.device("cpu")

    def allgather_int(self, value: int):
        """MPI_Allgather of one int64."""
        if self.size == 1:
            return np.array([value], dtype=np.int64)
        import torch
        import torch.distributed as dist

        dev = self._device()
        t = torch.tensor([value], dtype=torch.int64, device=dev)
        outs = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.size)]
        dist.all_gather(outs, t, group=self._group)
        return np.array([int(o.item()) for o in outs], dtype=np.int64)

    def allgather_array(self, arr: np.ndarray):
        """Variable-length MPI_Allgatherv of int64 arrays -> list per rank."""
        if self.size == 1:
            return [np.asarray(arr, dtype=np.int64)]
        import torch
        import torch.distributed as dist

        arr = np.ascontiguousarray(arr, dtype=np.int64)
        sizes = self.allgather_int(arr.size)
        m = int(sizes.max())
        dev = self._device()
        buf = torch.zeros(m, dtype=torch.int64, device=dev)
        buf[: arr.size] = torch.from_numpy(arr).to(dev)
        outs = [torch.zeros(m, dtype=torch.int64, device=dev) for _ in range(self.size)]
        dist.all_gather(outs, buf, group=self._group)
        return [o[: int(s)].cpu().numpy() for o, s in zip(outs, sizes)]

    def allreduce_sum(self, value: float) -> float:
        """MPI_Allreduce(MPI_SUM) of one double (la/Vector.h:457-458)."""
        if self.size == 1:
            return float(value)
        import torch
        import torch.distributed as dist

        t = torch.tensor([value], dtype=torch.float64, device=self._device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._group)
        return float(t.item())

    def allreduce_max(self, value: float) -> float:
        if self.size == 1:
            return float(value)
        import torch
        import torch.distributed as dist

        t = torch.tensor([value], dtype=torch.float64, device=self._device())
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self._group)
        return float(t.item())

    def neighbor_alltoallv(self, out_ranks, in_ranks, send, dtype=np.int64):
        """MPI_Neighbor_alltoallv on the dist-graph communicator (sources=in_ranks, destinations=out_ranks).

        ``send[i]`` goes to ``out_ranks[i]``; returns ``recv[j]`` from ``in_ranks[j]``.  Counts are
        exchanged first (the reference's MPI_Neighbor_alltoall of sizes).
        """
        out_ranks = [int(r) for r in out_ranks]
        in_ranks = [int(r) for r in in_ranks]
        if self.size == 1:
            assert not out_ranks and not in_ranks
            return []
        import torch
        import torch.distributed as dist

        tdtype = {np.int64: torch.int64, np.int32: torch.int32, np.float64: torch.float64}[np.dtype(dtype).type]
        dev = self._device()
        send = [np.ascontiguousarray(s, dtype=dtype) for s in send]
        # sizes
        sz_out = [torch.tensor([s.size], dtype=torch.int64, device=dev) for s in send]
        sz_in = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in in_ranks]
        ops = [dist.P2POp(dist.irecv, sz_in[j], self._peer(in_ranks[j]), group=self._group) for j in range(len(in_ranks))]
        ops += [dist.P2POp(dist.isend, sz_out[i], self._peer(out_ranks[i]), group=self._group) for i in range(len(out_ranks))]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        bufs_out = [torch.from_numpy(s).to(dev) for s in send]
        bufs_in = [torch.zeros(int(s.item()), dtype=tdtype, device=dev) for s in sz_in]
        ops = [dist.P2POp(dist.irecv, bufs_in[j], self._peer(in_ranks[j]), group=self._group) for j in range(len(in_ranks)) if bufs_in[j].numel()]
        ops += [dist.P2POp(dist.isend, bufs_out[i], self._peer(out_ranks[i]), group=self._group) for i in range(len(out_ranks)) if bufs_out[i].numel()]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        return [b.cpu().numpy() for b in bufs_in]

    def _peer(self, r):
        import torch.distributed as dist

        return dist.get_global_rank(self._group, r) if self._group is not None else r

    def barrier(self):
        if self.size > 1:
            import torch.distributed as dist

            dist.barrier(group=self._group)

    # -- device communicator -------------------------------------------------------------------
    @property
    def nccl(self):
        """libbfx NCCL communicator handle (None for one rank)."""
        if self.size == 1:
            return None
        if self._nccl is None:
            import torch
            import torch.distributed as dist

            from . import _lib

            ident = C.create_string_buffer(128)
            if self.rank == 0:
                _lib.check(_lib.lib.bfx_comm_unique_id(ident))
            dev = self._device()
            t = torch.frombuffer(bytearray(ident.raw), dtype=torch.uint8).clone().to(dev)
            dist.broadcast(t, src=self._peer(0), group=self._group)
            raw = bytes(t.cpu().numpy().tobytes())
            h = C.c_void_p()
            _lib.check(_lib.lib.bfx_comm_create(C.byref(h), raw, self.rank, self.size))
            self._nccl = h
        return self._nccl


COMM_SELF = Comm(serial=True)


def comm_world() -> Comm:
    return Comm()


class IndexMap:
    """dolfinx::common::IndexMap accessor subset (common/IndexMap.h:178-257).

    Local index = owned first, then ghosts in stored order.  Constructors mirror
    common/IndexMap.cpp:865-932: offset by exclusive scan, global size by all-reduce,
    src = sorted unique ghost owners, dest = ranks that ghost my indices.
    """

    def __init__(self, comm: Comm, size_local: int, ghosts=None, owners=None, src_dest=None):
        self.comm = comm
        self._size_local = int(size_local)
        self.ghosts = np.ascontiguousarray(ghosts if ghosts is not None else [], dtype=np.int64)
        self.owners = np.ascontiguousarray(owners if owners is not None else [], dtype=np.int32)
        assert self.ghosts.size == self.owners.size
        sizes = comm.allgather_int(self._size_local)
        off = int(sizes[: comm.rank].sum())
        self.local_range = (off, off + self._size_local)
        self.size_global = int(sizes.sum())
        if src_dest is not None:
            self.src = np.ascontiguousarray(src_dest[0], dtype=np.int32)
            self.dest = np.ascontiguousarray(src_dest[1], dtype=np.int32)
        else:
            # build_src_dest (common/IndexMap.cpp): src = sorted unique owners; dest from who lists me
            self.src = np.unique(self.owners).astype(np.int32)
            if comm.size == 1:
                self.dest = np.zeros(0, dtype=np.int32)
            else:
                all_src = comm.allgather_array(self.src)
                self.dest = np.array([r for r in range(comm.size) if comm.rank in all_src[r]], dtype=np.int32)
        assert np.all(np.diff(self.src) > 0) and np.all(np.diff(self.dest) > 0)

    @property
    def size_local(self) -> int:
        return self._size_local

    @property
    def num_ghosts(self) -> int:
        return int(self.ghosts.size)

    def local_to_global(self, local):
        """common/IndexMap.cpp:957-974"""
        local = np.asarray(local)
        out = np.empty(local.shape, dtype=np.int64)
        m = local < self._size_local
        out[m] = self.local_range[0] + local[m]
        out[~m] = self.ghosts[local[~m] - self._size_local]
        return out


class Scatterer:
    """dolfinx::common::Scatterer plan (common/Scatterer.h:65-198).

    Members have the reference's names and meaning: ``local_inds`` / ``remote_inds`` already
    expanded by the block size, sizes / displacements per neighbour (x bs).  ``device_plan()``
    uploads the plan to the GPU (bfx_scatter_create) for the data-path exchanges.
    """

    def __init__(self, index_map: IndexMap, bs: int):
        m = index_map
        self.map, self.bs = m, int(bs)
        self.src, self.dest = m.src.copy(), m.dest.copy()
        z = np.zeros(0, dtype=np.int32)
        self.local_inds, self.remote_inds = z, z
        self.sizes_remote = np.zeros(len(self.src), dtype=np.int32)
        self.displs_remote = np.zeros(len(self.src) + 1, dtype=np.int32)
        self.sizes_local = np.zeros(len(self.dest), dtype=np.int32)
        self.displs_local = np.zeros(len(self.dest) + 1, dtype=np.int32)
        self._plan = None
        if m.comm.size == 1:  # Scatterer.h:71-72
            return
        # stable sort of ghost positions by owner (Scatterer.h:98-101)
        perm = np.argsort(m.owners, kind="stable").astype(np.int32)
        owners_sorted = m.owners[perm]
        ghosts_sorted = m.ghosts[perm]
        # sizes / displacements of remote data per owning rank (Scatterer.h:122-130)
        hi = np.searchsorted(owners_sorted, self.src, side="right")
        self.displs_remote[1:] = hi
        self.sizes_remote[:] = np.diff(self.displs_remote)
        # send ghost global indices to the owners (comm1: ghost -> owner) (Scatterer.h:142-159)
        send = [ghosts_sorted[self.displs_remote[i]:self.displs_remote[i + 1]] for i in range(len(self.src))]
        recv = m.comm.neighbor_alltoallv(self.src, self.dest, send, dtype=np.int64)
        self.sizes_local[:] = [len(a) for a in recv]
        self.displs_local[1:] = np.cumsum(self.sizes_local)
        recv_buffer = np.concatenate(recv) if recv else np.zeros(0, dtype=np.int64)
        if np.any((recv_buffer < m.local_range[0]) | (recv_buffer >= m.local_range[1])):
            raise RuntimeError("Scatterer: received index outside the owned range")
        # scale by block size and expand (Scatterer.h:171-197)
        k = np.arange(bs, dtype=np.int64)
        self.local_inds = ((recv_buffer[:, None] * bs + k[None, :]) - m.local_range[0] * bs).reshape(-1).astype(np.int32)
        self.remote_inds = (perm.astype(np.int64)[:, None] * bs + k[None, :]).reshape(-1).astype(np.int32)
        for name in ("sizes_local", "displs_local", "sizes_remote", "displs_remote"):
            setattr(self, name, (getattr(self, name) * bs).astype(np.int32))

    def local_indices(self):
        return self.local_inds

    def remote_indices(self):
        return self.remote_inds

    def num_p2p_requests(self):
        return len(self.dest) + len(self.src)

    def new_device_plan(self):
        """A fresh bfx_scatter_t (device index arrays, staging buffers, communication stream, events) from the host
        arrays of this plan.  Every la::Vector owns one, like the buffers and the request of the reference's Vector
        (la/Vector.h:131-138), so that exchanges of two vectors on the same IndexMap may be in flight together."""
        from . import _lib

        h = C.c_void_p()
        li = np.ascontiguousarray(self.local_inds)
        ri = np.ascontiguousarray(self.remote_inds)
        _lib.check(
            _lib.lib.bfx_scatter_create(
                C.byref(h), self.map.comm.nccl, li.ctypes.data, li.size, ri.ctypes.data, ri.size,
                self.sizes_local.ctypes.data, self.displs_local.ctypes.data,
                np.ascontiguousarray(self.dest).ctypes.data, len(self.dest), self.sizes_remote.ctypes.data,
                self.displs_remote.ctypes.data, np.ascontiguousarray(self.src).ctypes.data, len(self.src),
            )
        )
        return h

    def device_plan(self):
        """The plan's own device instance (callers that drive bfx_scatter_* themselves; one exchange at a time)."""
        if self._plan is None:
            self._plan = self.new_device_plan()
            release_with(self, "bfx_scatter_destroy", self._plan)
        return self._plan


def _release(name, handle):
    from . import _lib

    getattr(_lib.lib, name)(handle)


def release_with(owner, destroy_name: str, handle):
    """Destroy a libbfx handle when ``owner`` is collected (not at interpreter exit: the process teardown frees the
    device and the library may already be unloaded)."""
    fin = weakref.finalize(owner, _release, destroy_name, handle)
    fin.atexit = False
    return fin


_scatterer_cache = weakref.WeakValueDictionary()


def cached_scatterer(index_map: IndexMap, bs: int) -> Scatterer:
    """The HOST plan of common::Scatterer is built once per (IndexMap, bs) and shared by the vectors on that map (its
    constructor is a neighbourhood exchange, common/Scatterer.h:142-159); device buffers are per Vector."""
    key = (id(index_map), bs)
    sc = _scatterer_cache.get(key)
    if sc is None or sc.map is not index_map:
        sc = Scatterer(index_map, bs)
        _scatterer_cache[key] = sc
    return sc
