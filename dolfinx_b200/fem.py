"""Host mirror of dolfinx::fem for the assembly hot path, over the libbfx C-ABI.

Reference: cpp/dolfinx/fem/{assembler.h,assemble_matrix_impl.h,assemble_vector_impl.h,pack.h,
DirichletBC.h,Form.h,DofMap.h,FunctionSpace.h,utils.h} and the Python surface
python/dolfinx/fem/{assemble.py,bcs.py,forms.py}.  Function names, argument meaning and error
behaviour follow the reference; the cell loops run as CUDA kernels.  ``Mesh``, ``DofMap``,
``FunctionSpace`` only carry the arrays the path reads (their builders are out of scope).
"""

from __future__ import annotations

import ctypes as C
import os
import enum
import weakref
from dataclasses import dataclass, field

import numpy as np

from . import la
from .common import Comm, IndexMap


class IntegralType(enum.IntEnum):
    cell = 0
    exterior_facet = 1
    interior_facet = 2
    vertex = 3


def _torch():
    import torch

    return torch


def _to_dev(a, dtype=None):
    torch = _torch()
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and a.dtype != dtype:
        a = a.to(dtype)
    return a.to(la._device()).contiguous()


class Mesh:
    """The part of mesh::Mesh the assembler reads: geometry x (N,3), geometry dofmap (C,nx)
    (mesh/Geometry.h:131,155) and the number of owned cells (default cell domain, fem/utils.h:579-588)."""

    def __init__(self, comm: Comm, x, x_dofmap, cell_type: str, num_cells_local=None):
        self.comm = comm
        self.cell_type = cell_type
        self._x_host = np.ascontiguousarray(x, dtype=np.float64) if isinstance(x, np.ndarray) else None
        self._xd_host = np.ascontiguousarray(x_dofmap, dtype=np.int32) if isinstance(x_dofmap, np.ndarray) else None
        self._x_dev = None if isinstance(x, np.ndarray) else x
        self._xd_dev = None if isinstance(x_dofmap, np.ndarray) else x_dofmap
        shape = x_dofmap.shape
        self.num_cells = int(shape[0])
        self.num_cells_local = self.num_cells if num_cells_local is None else int(num_cells_local)
        self.nx = int(shape[1])

    @property
    def x(self):
        """Device coordinates (N,3) float64."""
        if self._x_dev is None:
            self._x_dev = _to_dev(self._x_host)
        return self._x_dev

    @property
    def x_dofmap(self):
        if self._xd_dev is None:
            self._xd_dev = _to_dev(self._xd_host, _torch().int32)
        return self._xd_dev

    @property
    def x_host(self):
        if self._x_host is None:
            self._x_host = self._x_dev.cpu().numpy()
        return self._x_host

    @property
    def x_dofmap_host(self):
        if self._xd_host is None:
            self._xd_host = self._xd_dev.cpu().numpy()
        return self._xd_host


class DofMap:
    """fem::DofMap accessor subset (fem/DofMap.h:127-167): ``map`` (C,nd) int32, ``bs``, ``index_map``."""

    def __init__(self, dofmap, bs: int, index_map: IndexMap, index_map_bs=None):
        self._host = np.ascontiguousarray(dofmap, dtype=np.int32) if isinstance(dofmap, np.ndarray) else None
        self._dev = None if isinstance(dofmap, np.ndarray) else dofmap
        self.shape = tuple(dofmap.shape)
        self.bs = int(bs)
        self.index_map = index_map
        self.index_map_bs = int(bs if index_map_bs is None else index_map_bs)

    def map(self):
        if self._host is None:
            self._host = self._dev.cpu().numpy()
        return self._host

    def cell_dofs(self, c):
        return self.map()[c]

    @property
    def dev(self):
        if self._dev is None:
            self._dev = _to_dev(self._host, _torch().int32)
        return self._dev


class FunctionSpace:
    """fem::FunctionSpace: mesh + element label + dofmap (fem/FunctionSpace.h)."""

    def __init__(self, mesh: Mesh, element: str, dofmap: DofMap):
        self.mesh, self.element, self.dofmap = mesh, element, dofmap
        self._markers = {}  # bc marker arrays by bc combination (_bc_markers)

    def contains(self, V) -> bool:
        """fem/FunctionSpace.h:153 — no sub-spaces on this path: identity."""
        return V is self

    @property
    def space_dimension(self):
        return self.dofmap.shape[1] * self.dofmap.bs


class Function:
    """fem::Function: a FunctionSpace + la::Vector of dof values (block size = dofmap.bs)."""

    def __init__(self, V: FunctionSpace):
        self.function_space = V
        self.x = la.Vector(V.dofmap.index_map, V.dofmap.bs)


class Constant:
    """fem::Constant (fem/Constant.h): flattened value array."""

    def __init__(self, value):
        self.value = np.atleast_1d(np.asarray(value, dtype=np.float64)).reshape(-1)


@dataclass
class IntegralData:
    """fem::integral_data (fem/Form.h:52-87): kernel, entities, active coefficient indices.

    ``kernel`` is a libbfx kernel id (the device replacement of the FFCx function pointer).
    ``entities``: cells (n,) for cell integrals, (n,2) (cell, local_facet) pairs for exterior facets;
    None = all owned cells (fem/utils.h:579-588)."""

    kernel: int
    entities: object = None
    coeffs: list = field(default_factory=list)


class Form:
    """fem::Form (fem/Form.h:116-668), hand-built like cpp/demo/custom_kernel/main.cpp:77-81 and
    python/test/unit/fem/test_custom_jit_kernels.py:91-104."""

    def __init__(self, function_spaces, integrals: dict, coefficients=(), constants=(), mesh: Mesh = None):
        self.function_spaces = list(function_spaces)
        self.rank = len(self.function_spaces)
        self._integrals = {}
        for key, lst in integrals.items():
            itype = key if isinstance(key, IntegralType) else IntegralType(key)
            for idx, entry in enumerate(lst):
                if not isinstance(entry, IntegralData):
                    ident, kernel, entities, active = entry
                    entry = IntegralData(kernel, entities, list(active))
                else:
                    ident = idx
                self._integrals[(itype, ident)] = entry
        self.coefficients = list(coefficients)
        self.constants = list(constants)
        self.mesh = mesh if mesh is not None else self.function_spaces[0].mesh
        # cache of derived device data: entity lists, bfx_asm_t plans (keyed on the integral and on the serial number
        # of the matrix whose sparsity their position tables index), aggregated-strategy flags.  The plans die with
        # the form, or with their matrix, whichever goes first (_release_plans / _drop_matrix_plans).
        self._plans = {}
        weakref.finalize(self, _release_plans, self._plans).atexit = False

    def integral_ids(self, itype):
        return sorted(i for (t, i) in self._integrals if t == itype)

    def integral(self, itype, ident) -> IntegralData:
        return self._integrals[(itype, ident)]

    def coefficient_offsets(self):
        """fem/Form.h:593-604"""
        n = [0]
        for c in self.coefficients:
            n.append(n[-1] + c.function_space.space_dimension)
        return n


def _is_plan_entry(key):
    return isinstance(key, tuple) and key and key[0] in ("plan", "dS")


def _destroy_plan_entry(plans, key):
    """bfx_asm_destroy of one cached plan; the flags keyed on its address go with it (addresses are reused)."""
    from . import _lib

    h = plans.pop(key)[0]
    for k in [k for k in plans if isinstance(k, tuple) and k[0] in ("aggplan", "vchunks", "groups") and k[1] == h.value]:
        del plans[k]
    _lib.lib.bfx_asm_destroy(h)


def _release_plans(plans):
    for key in [k for k in plans if _is_plan_entry(k)]:
        _destroy_plan_entry(plans, key)
    plans.clear()


def _drop_matrix_plans(plans, serial):
    """The matrix with this serial number is gone: its plans (CSR positions of a sparsity that no longer exists) too."""
    for key in [k for k in plans if _is_plan_entry(k) and k[2] == serial]:
        _destroy_plan_entry(plans, key)
    plans.pop(("matrix", serial), None)
    plans.pop(("lean_attempt", serial), None)


def _matrix_key(form: Form, A):
    """Cache key component for the matrix a plan belongs to (None for vectors / functionals); the first plan of a
    (form, matrix) pair registers the clean-up that runs when the matrix is collected."""
    if A is None:
        return None
    reg = ("matrix", A._serial)
    if reg not in form._plans:
        form._plans[reg] = True
        weakref.finalize(A, _drop_matrix_plans, form._plans, A._serial).atexit = False
    return A._serial


def pack_constants(form: Form) -> np.ndarray:
    """fem::pack_constants (fem/pack.h:578-619)."""
    if not form.constants:
        return np.zeros(0)
    return np.concatenate([c.value for c in form.constants])


def _entities_dev(form: Form, integ: IntegralData, itype):
    torch = _torch()
    key = ("ent", id(integ))
    if key not in form._plans:
        ent = integ.entities
        if ent is None:
            form._plans[key] = (None, form.mesh.num_cells_local)
        else:
            t = _to_dev(np.ascontiguousarray(ent, dtype=np.int32) if isinstance(ent, np.ndarray) else ent, torch.int32)
            n = t.shape[0]
            form._plans[key] = (t.reshape(-1), n)
    return form._plans[key]


def pack_coefficients(form: Form):
    """fem::pack_coefficients (fem/pack.h:265-…): {(type, id): (coeffs (n, cstride) device tensor, cstride)}.

    Materialises the reference's packed layout on the device.  The assemblers below do not need it
    (they gather coefficients inside the element kernel) but accept it for parity with callers that
    pre-pack (python/test/unit/fem/test_assembler.py:1111-1205)."""
    from . import _lib

    torch = _torch()
    offsets = form.coefficient_offsets()
    cstride = offsets[-1]
    out = {}
    for (itype, ident), integ in form._integrals.items():
        if itype == IntegralType.interior_facet:
            # (nf, 2, cstride): the two restrictions of every coefficient, cell0 then cell1 (fem/pack.h:196-226)
            f = _interior_facets_dev(form, integ)
            c = torch.zeros((f.shape[0], 2, cstride), dtype=torch.float64, device=la._device())
            for k in integ.coeffs:
                u = form.coefficients[k]
                dm = u.function_space.dofmap
                for side in range(2):
                    d = dm.dev[f[:, side, 0]].long()  # (nf, nd)
                    vals = u.x.array.view(-1, dm.bs)[d].reshape(f.shape[0], -1)
                    c[:, side, offsets[k]:offsets[k] + vals.shape[1]] = vals
            out[(itype, ident)] = (c.reshape(f.shape[0], 2 * cstride), 2 * cstride)
            continue
        ent, n = _entities_dev(form, integ, itype)
        c = torch.zeros((n, cstride), dtype=torch.float64, device=la._device())
        for k in integ.coeffs:
            u = form.coefficients[k]
            dm = u.function_space.dofmap
            is_facet = itype == IntegralType.exterior_facet
            _lib.check(_lib.lib.bfx_pack_coefficient(
                c.data_ptr(), cstride, offsets[k], u.x.array.data_ptr(), dm.dev.data_ptr(), dm.shape[1], dm.bs,
                None if (ent is None or is_facet) else ent.data_ptr(), ent.data_ptr() if is_facet else None, n,
                _lib.current_stream()))
        out[(itype, ident)] = (c, cstride)
    return out


def create_sparsity_pattern(a: Form) -> la.SparsityPattern:
    """fem::create_sparsity_pattern (fem/utils.h:197-218) → sparsitybuild::cells over every cell /
    exterior-facet integral domain (fem/utils.h:225-314)."""
    assert a.rank == 2
    V0, V1 = a.function_spaces
    dm0, dm1 = V0.dofmap, V1.dofmap
    sp = la.SparsityPattern(a.mesh.comm, [dm0.index_map, dm1.index_map], [dm0.index_map_bs, dm1.index_map_bs])
    on_device = _torch().cuda.is_available()
    all_cells = False
    for (itype, ident), integ in a._integrals.items():
        if itype == IntegralType.cell and integ.entities is None:
            all_cells = True
    if all_cells or not a._integrals:
        n = a.mesh.num_cells_local
        if on_device:
            cells = None if n == dm0.shape[0] else _torch().arange(n, dtype=_torch().int32, device=la._device())
            sp.insert_cells(cells, dm0.dev, dm1.dev)
        else:
            sp.insert_cells(np.arange(n), dm0.map(), dm1.map())
    for (itype, ident), integ in a._integrals.items():
        if integ.entities is None:
            continue
        ent = np.asarray(integ.entities if isinstance(integ.entities, np.ndarray) else integ.entities.cpu().numpy())
        if itype == IntegralType.interior_facet:
            # sparsitybuild::interior_facets (fem/sparsitybuild.h:52-85): joint dofs of the two cells of every facet
            f = ent.reshape(-1, 2, 2)
            d0, d1 = dm0.map(), dm1.map()
            sp.insert_cells(None, np.concatenate([d0[f[:, 0, 0]], d0[f[:, 1, 0]]], axis=1),
                            np.concatenate([d1[f[:, 0, 0]], d1[f[:, 1, 0]]], axis=1))
            continue
        cells = ent if itype == IntegralType.cell else ent.reshape(-1, 2)[:, 0]
        if all_cells:
            continue  # already covered
        sp.insert_cells(np.unique(cells), dm0.map(), dm1.map())
    return sp


# ---------------------------------------------------------------------------------------------
# Dirichlet boundary conditions
# ---------------------------------------------------------------------------------------------
class DirichletBC:
    """fem::DirichletBC (fem/DirichletBC.h:262-601): sorted dof list (block indices unrolled by the
    block size, :357-361), value g (Function or Constant), number of owned dofs (:262-271)."""

    def __init__(self, g, dofs, V: FunctionSpace = None):
        torch = _torch()
        if isinstance(g, Function):
            V = g.function_space if V is None else V
        if V is None:
            raise RuntimeError("DirichletBC with a Constant needs the function space")
        self.function_space = V
        self.g = g
        bs = V.dofmap.bs
        dofs = np.asarray(dofs, dtype=np.int32)
        if isinstance(g, Constant) and g.value.size != bs:
            raise RuntimeError(
                "Creating a DirichletBC using a Constant is not supported when the Constant size is not equal to the "
                "block size of the constrained (sub-)space. Use a fem::Function to create the fem::DirichletBC.")
        if bs > 1:
            dofs = (bs * dofs[:, None] + np.arange(bs, dtype=np.int32)[None, :]).reshape(-1)
        self._dofs0 = np.ascontiguousarray(dofs, dtype=np.int32)
        owned_size = V.dofmap.index_map_bs * V.dofmap.index_map.size_local
        self._owned_indices0 = int(np.searchsorted(self._dofs0, owned_size))
        self._dofs0_dev = None
        self._g_dev = None

    def dof_indices(self):
        """(unrolled dofs, number of owned) — fem/DirichletBC.h:465-468"""
        return self._dofs0, self._owned_indices0

    @property
    def dofs_dev(self):
        if self._dofs0_dev is None:
            self._dofs0_dev = _to_dev(self._dofs0, _torch().int32)
        return self._dofs0_dev

    def mark_dofs(self, markers):
        """DirichletBC::mark_dofs (fem/DirichletBC.h:589-601); markers: int8 device tensor."""
        from . import _lib

        if self._dofs0.size and int(self._dofs0.max()) >= markers.numel():
            raise RuntimeError("Marker array is too short for the boundary condition dofs.")
        _lib.check(_lib.lib.bfx_bc_mark(markers.data_ptr(), self.dofs_dev.data_ptr(), self._dofs0.size, _lib.current_stream()))

    def set(self, x, x0=None, alpha: float = 1.0):
        """DirichletBC::set (fem/DirichletBC.h:495-578): x[dof] = alpha (g[dof] - x0[dof])."""
        from . import _lib

        if isinstance(self.g, Function):
            g_dev, kind = self.g.x.array, 0
        else:
            if self._g_dev is None:
                self._g_dev = _to_dev(self.g.value)
            g_dev, kind = self._g_dev, 1
        if x0 is not None:
            assert x.numel() <= x0.numel()
        _lib.check(_lib.lib.bfx_bc_set(x.data_ptr(), x.numel(), self.dofs_dev.data_ptr(), None, self._dofs0.size,
                                       g_dev.data_ptr(), kind, self.function_space.dofmap.bs,
                                       None if x0 is None else x0.data_ptr(), float(alpha), _lib.current_stream()))


def dirichletbc(value, dofs, V=None) -> DirichletBC:
    return DirichletBC(value, dofs, V)


def set_bc(b, bcs, x0=None, alpha: float = 1.0):
    """python/dolfinx/fem/bcs.py set_bc → DirichletBC::set for every bc."""
    arr = b.array if isinstance(b, la.Vector) else b
    for bc in bcs:
        bc.set(arr, x0, alpha)


# ---------------------------------------------------------------------------------------------
# assembly
# ---------------------------------------------------------------------------------------------
def _bc_markers(V: FunctionSpace, bcs):
    """dof markers of length bs*(owned+ghost) (fem/assembler.h:558-577); None when no bc applies.

    The dof set of a DirichletBC is fixed at construction, so the marker array of a (space, bcs) combination is built
    once and kept on the space (a time loop re-assembles with the same bcs every step)."""
    torch = _torch()
    mine = [bc for bc in bcs if V.contains(bc.function_space)]
    if not mine:
        return None
    key = tuple(id(bc) for bc in mine)
    hit = V._markers.get(key)
    if hit is not None and all(r() is bc for r, bc in zip(hit[1], mine)):
        return hit[0]
    im = V.dofmap.index_map
    mk = torch.zeros(V.dofmap.index_map_bs * (im.size_local + im.num_ghosts), dtype=torch.int8, device=la._device())
    for bc in mine:
        bc.mark_dofs(mk)
    if len(V._markers) >= 8:
        V._markers.clear()
    V._markers[key] = (mk, [weakref.ref(bc) for bc in mine])
    return mk


def _asm_plan(form: Form, integ: IntegralData, itype, A: la.MatrixCSR = None, subset=None):
    """bfx_asm_t for one integral of a form (cached on the form).

    ``subset`` = (tag, cells int32 device tensor) restricts a cell integral to part of its domain
    (used to overlap boundary-cell assembly with the ghost-row exchange)."""
    from . import _lib

    key = ("plan", id(integ), _matrix_key(form, A), subset[0] if subset else None)
    if key in form._plans:
        return form._plans[key][0]
    mesh = form.mesh
    # a functional (rank 0) has no test space: its plan carries the dofmap of its (first) coefficient, or - without
    # coefficients (M = 1 dx) - the geometry dofmap in that role
    if form.rank > 0:
        dm0 = form.function_spaces[0].dofmap
    elif form.coefficients:
        dm0 = form.coefficients[0].function_space.dofmap
    else:
        dm0 = None
    dm1 = form.function_spaces[1].dofmap if form.rank == 2 else None
    if subset is not None:
        ent, n = subset[1], int(subset[1].numel())
    elif itype == IntegralType.cell:
        ent, n = _entities_dev(form, integ, itype)
    else:
        ent, n = None, 0  # facet integrals pass their entities at call time
    if dm0 is not None:
        d0_ptr, nd0 = dm0.dev.data_ptr(), dm0.shape[1]
        n_rows = dm0.index_map.size_local + dm0.index_map.num_ghosts
    else:
        d0_ptr, nd0, n_rows = mesh.x_dofmap.data_ptr(), mesh.nx, int(mesh.x.shape[0])
    h = C.c_void_p()
    _lib.check(
        _lib.lib.bfx_asm_create(
            C.byref(h), A._csr if A is not None else None, mesh.x_dofmap.data_ptr(), mesh.nx, d0_ptr,
            nd0, dm1.dev.data_ptr() if dm1 is not None else None, dm1.shape[1] if dm1 is not None else 0,
            mesh.num_cells, None if ent is None else ent.data_ptr(), n, n_rows, 1,
            _lib.current_stream(),
        )
    )
    form._plans[key] = (h, ent)
    return h


# chunk plans of forms on one scalar space use symmetric (i,j)/(j,i) destination pairs; tests switch this off
# to exercise the general plan as well
# grouped vector assembly (one RED per distinct dof of 32 cells) is opt-in: on B200 it measures 2.6 ms against
# 2.1 ms of the cell-parallel kernel for the C2 load vector (both are bound by the L1/LSU pipe, not by the REDs)
GROUPED_VECTORS = os.environ.get("BFX_GROUPED_VECTORS", "0") != "0"
# chunk-aggregated vector assembly (round 2): one RED per chunk-boundary dof instead of one per (cell, local dof)
# measured at C2 / C3 / C4 (profiles/r02_vector_variants.txt): it wins for the P2 kernels (0.53 against 0.62 ms) and loses for
# the P1 / Q1 load vectors (2.44 against 2.08 ms, 3.65 against 2.65 ms), so only the P2 kernels take it by default
CHUNKED_VECTORS = os.environ.get("BFX_CHUNKED_VECTORS", "1") != "0"
CHUNKED_VECTORS_ALL = os.environ.get("BFX_CHUNKED_VECTORS", "1") == "all"
CHUNKS_SYMMETRIC = os.environ.get("BFX_CHUNKS_SYMMETRIC", "1") != "0"
# cells per chunk: 0 = the element's default (256 P1 / 128 P2); 96, 128, 192, 384 (P1) or 64, 96 (symmetric P2)
CHUNKS_CB = int(os.environ.get("BFX_CHUNKS_CB", "0"))
# write-back of the chunk sums in address order (symmetric P1 plans; see BFX_CHUNKS_TWO_STAGE in include/bfx.h)
# kernel variant of the chunk plans: 0 default, 1 = 5 CTAs per SM (BFX_CHUNK_OCC=5), 2 = "diet" list walk (BFX_CHUNK_DIET=1)
CHUNK_KERNEL = int(os.environ["BFX_CHUNK_DBG"]) if "BFX_CHUNK_DBG" in os.environ else (2 if os.environ.get("BFX_CHUNK_DIET", "0") != "0" else 1 if os.environ.get("BFX_CHUNK_OCC", "0") == "5" else 0)
# lean kernel for the symmetric P1-sized kernels (default): padded linear staging with bank-aware list order,
# destinations sorted by list length, 384-cell chunks (4 x 4 x 4 cubes of a Kuhn box under the Morton order):
# C2 launch 2.71 ms against 3.30 ms of the bank-coloured classic kernel (profiles/r02_p1_variants.txt)
CHUNK_LEAN = os.environ.get("BFX_CHUNK_LEAN", "1") != "0"
# distributed overlap of the lean plans: boundary chunks on a high-priority side stream beside the interior launch
# distributed overlap of the chunk-aggregated kernels: one plan of all cells in two launches (0: plans over cell subsets)
OVERLAP_ONE_PLAN = os.environ.get("BFX_OVERLAP_ONE_PLAN", "1") != "0"
OVERLAP_SIDE_STREAM = os.environ.get("BFX_OVERLAP_SIDE_STREAM", "0") != "0"  # measured at N = 2: no gain (3.32 against 3.22 ms per step)
CHUNKS_PAD4 = os.environ.get("BFX_CHUNKS_PAD4", "0") != "0"  # source lists padded to multiples of 4 (round-2 experiment)
CHUNKS_TWO_STAGE = int(os.environ.get("BFX_CHUNKS_TWO_STAGE", "0"))  # 1: one address-ordered list, 2: stores, then REDs


def _matrix_strategy(form: Form, integ: IntegralData, plan, strategy, shared=False, lean_only=False):
    """Assembly strategy of a cell integral: the caller's choice, else the aggregated kernel of the
    element where it has one (chunk-aggregated for the P1 kernels, row-gather for Q1 elasticity; the
    plan's lists are built once, on first use), else fp64 REDs."""
    from . import _lib

    if strategy == _lib.ASM_ATOMIC:
        return strategy
    want = strategy
    if want is None:
        want = (_lib.ASM_CHUNKED if integ.kernel in _lib.CHUNKED_KERNELS
                else _lib.ASM_ROWGATHER if integ.kernel in _lib.ROWGATHER_KERNELS else _lib.ASM_ATOMIC)
    if want == _lib.ASM_ATOMIC:
        return want
    key = ("aggplan", plan.value, want)
    if lean_only and key not in form._plans:
        key = ("aggplan", plan.value, want, "lean_only")  # (a refusal here must not be remembered as "no chunk plan")
    if key not in form._plans:
        if want == _lib.ASM_CHUNKED:
            V0, V1 = form.function_spaces
            flags = _lib.CHUNKS_SYMMETRIC if (CHUNKS_SYMMETRIC and V0 is V1 and V0.dofmap.bs == 1) else 0
            lean = CHUNK_LEAN and integ.kernel in _lib.LEAN_KERNELS and (flags & _lib.CHUNKS_SYMMETRIC) and not CHUNKS_TWO_STAGE
            flags |= _lib.CHUNKS_CB(CHUNKS_CB if CHUNKS_CB else (384 if lean else 0))
            if lean and not CHUNKS_CB:
                flags |= _lib.CHUNKS_CB_SOFT
            if CHUNKS_PAD4:
                flags |= _lib.CHUNKS_PAD4
            if lean:
                flags |= _lib.CHUNKS_LINEAR_STAGING
                if os.environ.get("BFX_CHUNKS_BANK_ORDER", "1") != "0":
                    flags |= _lib.CHUNKS_BANK_ORDER
                if os.environ.get("BFX_CHUNKS_LEN_SORT", "1") != "0":
                    flags |= _lib.CHUNKS_LEN_SORT
            if CHUNKS_TWO_STAGE:
                flags |= _lib.CHUNKS_TWO_STAGE | (_lib.CHUNKS_TWO_STAGE_SPLIT if int(CHUNKS_TWO_STAGE) == 2 else 0)
            if shared:  # a cell subset: other launches add to the same matrix (assemble_matrix_overlapped)
                flags |= _lib.CHUNKS_SHARED_MATRIX
            if lean_only:
                flags |= _lib.CHUNKS_LEAN_ONLY
            st = _lib.lib.bfx_asm_build_chunks(plan, form.mesh.x.data_ptr(), flags, _lib.current_stream())
            if st == _lib.OK and (CHUNK_KERNEL or lean):
                _lib.check(_lib.lib.bfx_asm_chunk_set_kernel(plan, CHUNK_KERNEL if CHUNK_KERNEL else _lib.CHUNK_KERNEL_LEAN))
        else:
            st = _lib.lib.bfx_asm_build_rowgather(plan, _lib.current_stream())
        if st not in (_lib.OK, _lib.ERR_UNSUPPORTED):
            _lib.check(st)
        form._plans[key] = st == _lib.OK
        if lean_only and st == _lib.OK:
            form._plans[("aggplan", plan.value, want)] = True
    if form._plans[key]:
        return want
    if lean_only:
        return None
    if strategy is not None:
        raise NotImplementedError("the requested assembly strategy is not available for this kernel / mesh")
    return _lib.ASM_ATOMIC


def chunk_stats(form: Form, A: la.MatrixCSR, itype=None, ident=0):
    """(chunks, destinations, source-list entries, plan bytes) of the chunk plan of a cell integral."""
    from . import _lib

    integ = form.integral(IntegralType.cell if itype is None else itype, ident)
    plan = _asm_plan(form, integ, IntegralType.cell, A)
    out = [C.c_int64() for _ in range(4)]
    _lib.check(_lib.lib.bfx_asm_chunk_stats(plan, *[C.byref(o) for o in out]))
    return tuple(int(o.value) for o in out)


def chunk_bank_conflicts(form: Form, A: la.MatrixCSR, ident=0):
    """Staged entries the plan's bank colouring could not make conflict free (-1: linear staging layout)."""
    from . import _lib

    plan = _asm_plan(form, form.integral(IntegralType.cell, ident), IntegralType.cell, A)
    n = C.c_int64()
    _lib.check(_lib.lib.bfx_asm_chunk_bank_conflicts(plan, C.byref(n)))
    return int(n.value)


def chunk_two_stage(form: Form, A: la.MatrixCSR, ident=0) -> bool:
    """True if the chunk plan of the cell integral writes back in address order (BFX_CHUNKS_TWO_STAGE)."""
    from . import _lib

    plan = _asm_plan(form, form.integral(IntegralType.cell, ident), IntegralType.cell, A)
    n = C.c_int()
    _lib.check(_lib.lib.bfx_asm_chunk_two_stage(plan, C.byref(n)))
    return bool(n.value)


def _interior_facets_dev(form: Form, integ: IntegralData):
    """(nf, 2, 2) int64 device tensor [[cell0, local0], [cell1, local1]] of an interior-facet integral."""
    torch = _torch()
    key = ("dSf", id(integ))
    if key not in form._plans:
        ent = integ.entities
        form._plans[key] = torch.as_tensor(
            np.ascontiguousarray(np.asarray(ent).reshape(-1, 2, 2), dtype=np.int64) if not torch.is_tensor(ent)
            else ent.reshape(-1, 2, 2).long().cpu(), device=la._device())
    return form._plans[key]


def _joint_dofmap(form: Form, integ: IntegralData, dm: DofMap):
    """[dofs(cell0) | dofs(cell1)] per interior facet (fem/assemble_matrix_impl.h:581-587), cached per dofmap."""
    torch = _torch()
    key = ("dSdm", id(integ), id(dm))
    if key not in form._plans:
        f = _interior_facets_dev(form, integ)
        form._plans[key] = (torch.cat([dm.dev[f[:, 0, 0]], dm.dev[f[:, 1, 0]]], dim=1).contiguous(), dm)
    return form._plans[key][0]


def _interior_facet_plan(form: Form, integ: IntegralData, A: la.MatrixCSR = None):
    """Plan of an interior-facet integral (fem/assemble_matrix_impl.h:442-667, fem/assemble_vector_impl.h:249-339,
    fem/assemble_scalar_impl.h:122-168): every facet becomes a MACRO cell whose geometry nodes and dofs are the joint
    arrays [cell0 | cell1]; the kernel receives the two local facet indices packed as lf0 + 8 lf1.  Bilinear forms
    need the matrix (CSR positions); linear forms and functionals pass None.
    Returns (plan, entities (facet, packed local facets) on device, n)."""
    from . import _lib

    torch = _torch()
    key = ("dS", id(integ), _matrix_key(form, A))
    if key not in form._plans:
        f = _interior_facets_dev(form, integ)
        n = int(f.shape[0])
        mesh = form.mesh
        xdm = torch.cat([mesh.x_dofmap[f[:, 0, 0]], mesh.x_dofmap[f[:, 1, 0]]], dim=1).contiguous()
        if form.rank >= 1:
            dm0 = form.function_spaces[0].dofmap
        else:  # a functional has no test space: the plan carries the dofmap of its first coefficient, or the geometry's
            dm0 = form.coefficients[0].function_space.dofmap if form.coefficients else None
        dm1 = form.function_spaces[1].dofmap if form.rank == 2 else None
        j0 = _joint_dofmap(form, integ, dm0) if dm0 is not None else xdm
        j1 = None if dm1 is None else (j0 if dm1 is dm0 else _joint_dofmap(form, integ, dm1))
        entities = torch.stack([torch.arange(n, device=f.device), f[:, 0, 1] + 8 * f[:, 1, 1]], dim=1).to(torch.int32).contiguous()
        if dm0 is not None:
            n_rows = dm0.index_map.size_local + dm0.index_map.num_ghosts
        else:
            n_rows = int(mesh.x.shape[0])
        h = C.c_void_p()
        _lib.check(_lib.lib.bfx_asm_create(
            C.byref(h), A._csr if A is not None else None, xdm.data_ptr(), xdm.shape[1], j0.data_ptr(), j0.shape[1],
            None if j1 is None else j1.data_ptr(), 0 if j1 is None else j1.shape[1], n, None, 0, n_rows, 1,
            _lib.current_stream()))
        form._plans[key] = (h, entities, n, (xdm, j0, j1))  # the joint arrays are borrowed by the plan: keep them
    return form._plans[key][:3]


def _boundary_interior_cells(form: Form, integ: IntegralData):
    """Split a cell domain into cells that touch a ghost row (their contributions must travel to the
    owner, la/MatrixCSR.h:399-468) and interior cells (SURVEY.md §8e overlap plan)."""
    torch = _torch()
    key = ("split", id(integ))
    if key not in form._plans:
        dm0 = form.function_spaces[0].dofmap
        n_owned = dm0.index_map.size_local
        ent, n = _entities_dev(form, integ, IntegralType.cell)
        dmap = dm0.dev if ent is None else dm0.dev[ent.long()]
        touches = (dmap[:n] >= n_owned).any(dim=1)
        ids = torch.arange(n, dtype=torch.int32, device=dmap.device) if ent is None else ent
        form._plans[key] = (ids[touches].contiguous(), ids[~touches].contiguous())
    return form._plans[key]


def _coeffs_for(form: Form, integ: IntegralData, packed, itype=None):
    """bfx_coeffs_t of an integral: the caller's packed array (reference layout), or the fused gather of the integral's
    active coefficients straight from their dof vectors (all of them, in the order of the kernel's w; on interior
    facets through the joint dofmaps [cell0 | cell1])."""
    from . import _lib

    if packed is not None:
        c, cstride = packed
        off = form.coefficient_offsets()[integ.coeffs[0]] if integ.coeffs else 0
        return _lib.make_coeffs(packed=c, cstride=cstride, offset=off)
    if integ.coeffs:
        if len(integ.coeffs) > 4:
            raise NotImplementedError("fused coefficient gather supports up to four active coefficients; pass pack_coefficients(form)")
        fused = []
        for k in integ.coeffs:
            u = form.coefficients[k]
            dm = u.function_space.dofmap
            if itype == IntegralType.interior_facet:
                fused.append((u.x.array, _joint_dofmap(form, integ, dm), 2 * dm.shape[1], dm.bs))
            else:
                fused.append((u.x.array, dm.dev, dm.shape[1], dm.bs))
        return _lib.make_coeffs(fused=fused)
    return _lib.make_coeffs()


def _translate(e):
    from . import _lib

    if isinstance(e, _lib.BfxError) and e.status == _lib.ERR_NOT_IN_SPARSITY:
        return RuntimeError("Entry not in sparsity")
    return e


def assemble_matrix(A: la.MatrixCSR, a: Form, bcs=(), constants=None, coeffs=None, strategy=None):
    """fem::assemble_matrix(A.mat_add_values(), a, bcs) — fem/assembler.h:513-630.

    Does not zero A and does not finalise it (assembler.h:497-498): call ``A.scatter_reverse()``
    afterwards.  bc rows/columns are zeroed in the element tensors (assemble_matrix_impl.h:161-196);
    use ``set_diagonal`` for the diagonal."""
    from . import _lib

    assert a.rank == 2
    V0, V1 = a.function_spaces
    bc0 = _bc_markers(V0, bcs)
    bc1 = bc0 if V1 is V0 else _bc_markers(V1, bcs)  # same space: one marker array (assembler.h:558-577)
    consts = pack_constants(a) if constants is None else np.asarray(constants, dtype=np.float64)
    carr, nc = _lib.constants_array(consts)
    try:
        for (itype, ident), integ in a._integrals.items():
            plan = None if itype == IntegralType.interior_facet else _asm_plan(a, integ, itype, A)
            cf = _coeffs_for(a, integ, None if coeffs is None else coeffs[(itype, ident)], itype)
            if itype == IntegralType.cell:
                strat = _matrix_strategy(a, integ, plan, strategy)
                mode = _lib.VALUES_OVERWRITE if A._is_zero else _lib.VALUES_ADD
                # a kernel that writes every value once absorbs a pending A.set_value(0)
                vals = A._take_zero_fill() if (A._is_zero and strat == _lib.ASM_ROWGATHER) else A._values()
                _lib.check(_lib.lib.bfx_assemble_matrix_cells(
                    plan, integ.kernel, a.mesh.x.data_ptr(), None if bc0 is None else bc0.data_ptr(),
                    None if bc1 is None else bc1.data_ptr(), C.byref(cf), carr, nc, vals.data_ptr(), strat, mode,
                    _lib.current_stream()))
            elif itype == IntegralType.exterior_facet:
                ent, n = _entities_dev(a, integ, itype)
                _lib.check(_lib.lib.bfx_assemble_matrix_facets(
                    plan, integ.kernel, a.mesh.x.data_ptr(), ent.data_ptr(), n, None if bc0 is None else bc0.data_ptr(),
                    None if bc1 is None else bc1.data_ptr(), C.byref(cf), carr, nc, A._values().data_ptr(),
                    _lib.current_stream()))
            elif itype == IntegralType.interior_facet:
                plan, ent, n = _interior_facet_plan(a, integ, A)
                _lib.check(_lib.lib.bfx_assemble_matrix_facets(
                    plan, integ.kernel, a.mesh.x.data_ptr(), ent.data_ptr(), n, None if bc0 is None else bc0.data_ptr(),
                    None if bc1 is None else bc1.data_ptr(), C.byref(cf), carr, nc, A._values().data_ptr(),
                    _lib.current_stream()))
            else:
                raise NotImplementedError(f"integral type {itype!r} is outside the hot path (SURVEY.md §8f)")
            A._is_zero = False
    except _lib.BfxError as e:
        raise _translate(e) from e
    return A


def assemble_matrix_overlapped(A: la.MatrixCSR, a: Form, bcs=(), constants=None, strategy=None, timeline=None):
    """``assemble_matrix`` + ``A.scatter_reverse()`` with the ghost-row exchange hidden behind the
    interior cells: boundary cells are assembled first, the NCCL exchange of the ghost rows starts on
    the communication stream, interior cells are assembled meanwhile, then the received values are
    added (north_star: "overlapped with interior-cell assembly").  Same result as the two separate
    reference calls (fem/assembler.h:589-602 then la/MatrixCSR.h:384-468) up to summation order."""
    from . import _lib

    def mark(name):
        # (profiling: ``timeline`` collects (phase, CUDA event) pairs recorded on the caller's stream)
        if timeline is not None:
            ev = _torch().cuda.Event(enable_timing=True)
            ev.record()
            timeline.append((name, ev))

    if a.mesh.comm.size == 1:
        assemble_matrix(A, a, bcs, constants=constants, strategy=strategy)
        A.scatter_reverse()
        return A
    V0, V1 = a.function_spaces
    bc0 = _bc_markers(V0, bcs)
    bc1 = bc0 if V1 is V0 else _bc_markers(V1, bcs)
    consts = pack_constants(a) if constants is None else np.asarray(constants, dtype=np.float64)
    carr, nc = _lib.constants_array(consts)
    items = list(a._integrals.items())
    if any(itype != IntegralType.cell for (itype, _), _ in items):
        raise NotImplementedError("overlapped assembly handles cell integrals")

    was_zero = A._is_zero
    if was_zero and len(items) == 1 and items[0][1].kernel in _lib.ROWGATHER_KERNELS and strategy in (None, _lib.ASM_ROWGATHER):
        # Row-gather kernels form every CSR row completely from the cells of this rank, so the overlap splits the ROWS:
        # ghost rows (the tail of the row range) first, their exchange behind the owned rows; every value is written
        # once, which also absorbs the pending zero-fill.
        (itype, ident), integ = items[0]
        plan = _asm_plan(a, integ, itype, A)
        if _matrix_strategy(a, integ, plan, strategy) == _lib.ASM_ROWGATHER:
            tr = C.c_int(0)
            _lib.check(_lib.lib.bfx_asm_rowgather_tile_rows(plan, C.byref(tr)))
            n_owned, n_all = A.num_owned_rows(), A.num_all_rows()
            if tr.value > 0:
                split = (n_owned // tr.value) * tr.value
                vals = A._take_zero_fill()
                cf = _coeffs_for(a, integ, None, itype)

                def rows(r0, r1, reuse):
                    _lib.check(_lib.lib.bfx_assemble_matrix_rows(
                        plan, integ.kernel, a.mesh.x.data_ptr(), None if bc0 is None else bc0.data_ptr(),
                        None if bc1 is None else bc1.data_ptr(), C.byref(cf), carr, nc, vals.data_ptr(), r0, r1, reuse,
                        _lib.current_stream()))

                try:
                    mark("start")
                    rows(split, n_all, 0)
                    mark("ghost rows")
                    A.scatter_rev_begin()
                    mark("pack + send (enqueue)")
                    rows(0, split, 1)
                    mark("owned rows")
                    A.scatter_rev_end()
                    mark("wait + unpack")
                except _lib.BfxError as e:
                    raise _translate(e) from e
                A._is_zero = False
                return A
    def room_for_lean_attempt():
        # The attempt builds the plan of ALL cells up to its warp tables (~250 bytes per cell at the peak: position map,
        # Morton keys, chunk-ordered dofmaps, tables) before it knows whether the mesh takes the lean options; on a shard
        # near the HBM limit (C5: 750 M cells per GPU) that is not affordable beside the subset plans of the other scheme.
        key = ("lean_attempt", A._serial)  # (decided once per (form, matrix))
        if key not in a._plans:
            torch = _torch()
            ncells = _entities_dev(a, items[0][1], IntegralType.cell)[1]
            if ncells >= (1 << 26):
                torch.cuda.empty_cache()  # blocks cached by torch are invisible to the cudaMalloc of the plan builder
            a._plans[key] = torch.cuda.mem_get_info()[0] >= 250 * ncells
        return a._plans[key]

    if (len(items) == 1 and items[0][1].kernel in _lib.CHUNKED_KERNELS and strategy in (None, _lib.ASM_CHUNKED) and CHUNK_LEAN
            and OVERLAP_ONE_PLAN and room_for_lean_attempt()):
        # Lean chunk plans: ONE plan of all cells, launched in two parts - the chunks with a cell on a ghost row first,
        # the others behind the exchange.  The chunks keep the whole-cube geometry and the completeness of the
        # one-launch plan (a plan over the interior cell SUBSET is cut through the cubes: 2.90 against 2.72 ms at C2).
        (itype, ident), integ = items[0]
        plan = _asm_plan(a, integ, itype, A)
        # (P1-sized kernels: only if the mesh takes the lean plan - otherwise the builder stops early and the subset
        # scheme below runs; P2: the plan of the classic kernel, launched through a chunk list / skip flags)
        if _matrix_strategy(a, integ, plan, strategy, lean_only=integ.kernel in _lib.LEAN_KERNELS) == _lib.ASM_CHUNKED:
            n1 = C.c_int64(0)
            st = _lib.lib.bfx_asm_chunk_partition(plan, A.num_owned_rows(), C.byref(n1))
            if st == _lib.OK:
                mark("start")
                vals = A._values()
                mark("zero fill")
                mode = _lib.VALUES_OVERWRITE if was_zero else _lib.VALUES_ADD
                cf = _coeffs_for(a, integ, None, itype)

                def part(k):
                    _lib.check(_lib.lib.bfx_assemble_matrix_cells_part(
                        plan, integ.kernel, a.mesh.x.data_ptr(), None if bc0 is None else bc0.data_ptr(),
                        None if bc1 is None else bc1.data_ptr(), C.byref(cf), carr, nc, vals.data_ptr(), mode, k,
                        _lib.current_stream()))

                try:
                    torch = _torch()
                    if n1.value > 0 and OVERLAP_SIDE_STREAM:
                        # the few boundary chunks run on a high-priority side stream NEXT TO the interior launch (they
                        # write disjoint complete entries and RED the shared ones): their tail does not idle the GPU, and
                        # the pack + send follow them on that stream
                        key = ("side", 0)
                        if key not in a._plans:
                            a._plans[key] = torch.cuda.Stream(priority=-1)
                        side, main = a._plans[key], torch.cuda.current_stream()
                        side.wait_stream(main)  # the zero fill
                        with torch.cuda.stream(side):
                            part(1)
                            A.scatter_rev_begin()
                        mark("boundary chunks + pack + send (side stream, enqueue)")
                        part(2)
                        mark("interior chunks")
                        A.scatter_rev_end()
                        mark("wait + unpack")
                        A._is_zero = False
                        return A
                    if n1.value > 0:
                        part(1)
                    mark("boundary chunks")
                    A.scatter_rev_begin()
                    mark("pack + send (enqueue)")
                    part(2 if n1.value > 0 else 0)  # (no chunk on a ghost row: the plain launch of all chunks)
                    mark("interior chunks")
                    A.scatter_rev_end()
                    mark("wait + unpack")
                except _lib.BfxError as e:
                    raise _translate(e) from e
                A._is_zero = False
                return A
            if st != _lib.ERR_UNSUPPORTED:
                _lib.check(st)
        # no lean plan for this mesh (incomplete warp tables; the builder stopped before the lists): the plan of all
        # cells is not needed by the subset scheme below - on a shard near the memory limit it must not stay beside
        # the two subset plans
        if ("aggplan", plan.value, _lib.ASM_CHUNKED) not in a._plans:
            _destroy_plan_entry(a._plans, ("plan", id(integ), _matrix_key(a, A), None))
    mark("start")
    vals = A._values()  # (a pending set_value(0) is written here)
    mark("zero fill")

    def run(tag_index):
        for (itype, ident), integ in items:
            cells = _boundary_interior_cells(a, integ)[tag_index]
            if cells.numel() == 0:
                continue
            plan = _asm_plan(a, integ, itype, A, subset=(("bnd", "int")[tag_index], cells))
            strat = _matrix_strategy(a, integ, plan, strategy, shared=True)
            # on a zeroed matrix both launches may overwrite: their chunk plans treat an entry as complete only
            # if no cell outside the chunk touches it (CHUNKS_SHARED_MATRIX); everything else is added
            mode = _lib.VALUES_OVERWRITE if (was_zero and strat == _lib.ASM_CHUNKED and len(items) == 1) else _lib.VALUES_ADD
            cf = _coeffs_for(a, integ, None)
            _lib.check(_lib.lib.bfx_assemble_matrix_cells(
                plan, integ.kernel, a.mesh.x.data_ptr(), None if bc0 is None else bc0.data_ptr(),
                None if bc1 is None else bc1.data_ptr(), C.byref(cf), carr, nc, vals.data_ptr(), strat,
                mode, _lib.current_stream()))

    try:
        run(0)
        mark("boundary cells")
        A.scatter_rev_begin()
        mark("pack + send (enqueue)")
        run(1)
        mark("interior cells")
        A.scatter_rev_end()
        mark("wait + unpack")
    except _lib.BfxError as e:
        raise _translate(e) from e
    A._is_zero = False
    return A


def assemble_scalar(M: Form, constants=None, coeffs=None) -> float:
    """fem::assemble_scalar(M) — fem/assembler.h:173-213: the functional summed over the owned cells of this
    rank (the caller reduces over ranks, like the reference: cpp/demo/poisson_matrix_free/main.cpp:241-247)."""
    from . import _lib

    assert M.rank == 0
    consts = pack_constants(M) if constants is None else np.asarray(constants, dtype=np.float64)
    carr, nc = _lib.constants_array(consts)
    total = 0.0
    for (itype, ident), integ in M._integrals.items():
        packed = None if coeffs is None else coeffs[(itype, ident)]
        cf = _coeffs_for(M, integ, packed, itype)
        out = C.c_double(0.0)
        if itype == IntegralType.cell:
            plan = _asm_plan(M, integ, itype, None)
            _lib.check(_lib.lib.bfx_assemble_scalar_cells(plan, integ.kernel, M.mesh.x.data_ptr(), C.byref(cf), carr, nc,
                                                          C.byref(out), _lib.current_stream()))
        elif itype == IntegralType.exterior_facet:
            # fem/assemble_scalar_impl.h:78-113
            plan = _asm_plan(M, integ, itype, None)
            ent, n = _entities_dev(M, integ, itype)
            _lib.check(_lib.lib.bfx_assemble_scalar_facets(plan, integ.kernel, M.mesh.x.data_ptr(), ent.data_ptr(), n,
                                                           C.byref(cf), carr, nc, C.byref(out), _lib.current_stream()))
        elif itype == IntegralType.interior_facet:
            # fem/assemble_scalar_impl.h:122-168
            plan, ent, n = _interior_facet_plan(M, integ, None)
            _lib.check(_lib.lib.bfx_assemble_scalar_facets(plan, integ.kernel, M.mesh.x.data_ptr(), ent.data_ptr(), n,
                                                           C.byref(cf), carr, nc, C.byref(out), _lib.current_stream()))
        else:
            raise NotImplementedError(f"assemble_scalar: integral type {itype!r} is outside the hot path (SURVEY.md §8f)")
        total += out.value
    return total


def assemble_vector(b, L: Form, constants=None, coeffs=None):
    """fem::assemble_vector(b, L) — fem/assembler.h:230-257: accumulates into b (not zeroed, ghosts not
    scattered: call ``b.scatter_reverse(InsertMode.add)``)."""
    from . import _lib

    assert L.rank == 1
    arr = b.array if isinstance(b, la.Vector) else b
    consts = pack_constants(L) if constants is None else np.asarray(constants, dtype=np.float64)
    carr, nc = _lib.constants_array(consts)
    for (itype, ident), integ in L._integrals.items():
        cf = _coeffs_for(L, integ, None if coeffs is None else coeffs[(itype, ident)], itype)
        if itype == IntegralType.interior_facet:
            # fem/assemble_vector_impl.h:249-339: element vector [cell0 | cell1] through the joint dofmap
            plan, ent, n = _interior_facet_plan(L, integ, None)
            _lib.check(_lib.lib.bfx_assemble_vector_facets(plan, integ.kernel, L.mesh.x.data_ptr(), ent.data_ptr(), n,
                                                           C.byref(cf), carr, nc, arr.data_ptr(), _lib.current_stream()))
            continue
        plan = _asm_plan(L, integ, itype, None)
        if itype == IntegralType.cell:
            strat = _lib.ASM_ATOMIC
            if (integ.kernel in (_lib.CHUNKED_VECTOR_KERNELS if CHUNKED_VECTORS_ALL else _lib.CHUNKED_VECTOR_DEFAULT)
                    and CHUNKED_VECTORS and not GROUPED_VECTORS):
                # element vectors summed per chunk of 256 / 384 cells on the SM (plan built once, on first use)
                key = ("vchunks", plan.value)
                if key not in L._plans:
                    # the plan is an optional accelerator: on a shard near the memory limit (its construction peaks at
                    # ~150 bytes per cell: Morton keys, chunk-ordered dofmaps, lists) the RED kernel stays
                    free = _torch().cuda.mem_get_info()[0]
                    ncells = _entities_dev(L, integ, itype)[1]
                    if free < 200 * ncells:
                        st = _lib.ERR_UNSUPPORTED
                    else:
                        st = _lib.lib.bfx_asm_build_chunks_vector(plan, L.mesh.x.data_ptr(), integ.kernel, _lib.current_stream())
                    if st == _lib.ERR_CUDA and b"out of memory" in _lib.lib.bfx_last_error():
                        st = _lib.ERR_UNSUPPORTED
                    if st not in (_lib.OK, _lib.ERR_UNSUPPORTED):
                        _lib.check(st)
                    L._plans[key] = st == _lib.OK
                if L._plans[key]:
                    strat = _lib.ASM_CHUNKED
            elif integ.kernel in _lib.GROUPED_VECTOR_KERNELS and GROUPED_VECTORS:
                # one RED per distinct dof of a group of 32 cells (plan built once, on first use)
                key = ("groups", plan.value)
                if key not in L._plans:
                    st = _lib.lib.bfx_asm_build_groups(plan, L.mesh.x.data_ptr(), _lib.current_stream())
                    if st not in (_lib.OK, _lib.ERR_UNSUPPORTED):
                        _lib.check(st)
                    L._plans[key] = st == _lib.OK
                if L._plans[key]:
                    strat = _lib.ASM_CHUNKED
            _lib.check(_lib.lib.bfx_assemble_vector_cells(plan, integ.kernel, L.mesh.x.data_ptr(), C.byref(cf), carr, nc,
                                                          arr.data_ptr(), strat, _lib.current_stream()))
        elif itype == IntegralType.exterior_facet:
            ent, n = _entities_dev(L, integ, itype)
            _lib.check(_lib.lib.bfx_assemble_vector_facets(plan, integ.kernel, L.mesh.x.data_ptr(), ent.data_ptr(), n,
                                                           C.byref(cf), carr, nc, arr.data_ptr(), _lib.current_stream()))
        else:
            raise NotImplementedError(f"integral type {itype!r} is outside the hot path (SURVEY.md §8f)")
    return b


import itertools as _itertools

_LIFT_SERIAL = _itertools.count()


def apply_lifting(b, a, bcs, x0=None, alpha: float = 1.0, constants=None, coeffs=None):
    """fem::apply_lifting — fem/assembler.h:336-493: b <- b - alpha A_j (g_j - x0_j) for every block j.

    ``a``: list of bilinear forms (or None), ``bcs``: list (per form) of lists of DirichletBC,
    ``x0``: optional list of device tensors."""
    from . import _lib

    torch = _torch()
    arr = b.array if isinstance(b, la.Vector) else b
    if all(ai is None for ai in a):
        return
    if x0 is not None and len(x0) != len(a):
        raise RuntimeError("Mismatch in size between x0 and bilinear form in assembler.")
    if len(a) != len(bcs):
        raise RuntimeError("Mismatch in size between a and bcs in assembler.")
    for j, aj in enumerate(a):
        if aj is None or not bcs[j]:
            continue
        V1 = aj.function_spaces[1]
        im1 = V1.dofmap.index_map
        crange = V1.dofmap.index_map_bs * (im1.size_local + im1.num_ghosts)
        markers = torch.zeros(crange, dtype=torch.int8, device=la._device())
        values = torch.zeros(crange, dtype=torch.float64, device=la._device())
        for bc in bcs[j]:
            bc.mark_dofs(markers)
            bc.set(values, None, 1.0)
        consts = pack_constants(aj) if constants is None else np.asarray(constants[j], dtype=np.float64)
        carr, nc = _lib.constants_array(consts)
        for (itype, ident), integ in aj._integrals.items():
            if itype != IntegralType.cell:
                raise NotImplementedError("lifting of facet integrals")
            # has_bc (fem/assemble_matrix_impl.h:27-34) once per (integral, bcs): the dof sets of a DirichletBC are
            # fixed at construction, so the cells with a marked column dof form a fixed boundary layer; the lifting
            # plan holds only those (the reference tests every cell of the domain on every call)
            lkey = ("liftcells", id(integ), tuple(id(bc) for bc in bcs[j]))
            if lkey not in aj._plans or any(r() is not bc for r, bc in zip(aj._plans[lkey][1], bcs[j])):
                if lkey in aj._plans:  # ids reused by other bc objects: the old cell list and its plan are stale
                    stale = ("plan", id(integ), None, ("lift", aj._plans[lkey][2]))
                    if stale in aj._plans:
                        _destroy_plan_entry(aj._plans, stale)
                ent, n = _entities_dev(aj, integ, itype)
                bsz = V1.dofmap.index_map_bs
                node_marked = markers.view(-1, bsz).any(dim=1) if bsz > 1 else markers.bool()
                picked = []
                for s0 in range(0, n, 1 << 24):  # (slices: the int64 index temporaries of a 750 M-cell shard are 24 GB)
                    s1 = min(n, s0 + (1 << 24))
                    rows = V1.dofmap.dev[s0:s1] if ent is None else V1.dofmap.dev[ent[s0:s1].long()]
                    hit = node_marked[rows.long()].any(dim=1)
                    ids = (torch.arange(s0, s1, dtype=torch.int32, device=rows.device) if ent is None else ent[s0:s1])
                    picked.append(ids[hit])
                    del rows, hit, ids
                lc = torch.cat(picked).contiguous() if picked else torch.zeros(0, dtype=torch.int32, device=markers.device)
                aj._plans[lkey] = (lc, [weakref.ref(bc) for bc in bcs[j]], next(_LIFT_SERIAL))
            lcells, _, serial = aj._plans[lkey]
            if lcells.numel() == 0:
                continue
            # (packed coefficient arrays are indexed by the position in the integral's own entity list: whole list then)
            plan = (_asm_plan(aj, integ, itype, None, subset=(("lift", serial), lcells)) if coeffs is None
                    else _asm_plan(aj, integ, itype, None))
            cf = _coeffs_for(aj, integ, None if coeffs is None else coeffs[j][(itype, ident)])
            _lib.check(_lib.lib.bfx_lift_bc_cells(
                plan, integ.kernel, aj.mesh.x.data_ptr(), C.byref(cf), carr, nc, arr.data_ptr(), values.data_ptr(),
                markers.data_ptr(), None if x0 is None else x0[j].data_ptr(), float(alpha), _lib.current_stream()))


def set_diagonal(A: la.MatrixCSR, V: FunctionSpace, bcs, diagonal: float = 1.0):
    """fem::set_diagonal — fem/assembler.h:644-686: A[dof,dof] = diagonal (SET) on OWNED bc rows."""
    from . import _lib

    for bc in bcs:
        if V.contains(bc.function_space):
            dofs, n_owned = bc.dof_indices()
            if n_owned == 0:
                continue
            try:
                _lib.check(_lib.lib.bfx_csr_set_diagonal(A._csr, A._values().data_ptr(), bc.dofs_dev.data_ptr(), n_owned,
                                                         float(diagonal), _lib.current_stream()))
            except _lib.BfxError as e:
                raise _translate(e) from e
            A._is_zero = False
