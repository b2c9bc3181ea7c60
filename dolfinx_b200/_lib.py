"""ctypes binding of libbfx.so (the C-ABI declared in include/bfx.h).

There is NO CPU fallback: importing this module fails loudly when the shared
library is missing, and every compute call fails with ``BfxError`` when no CUDA
device is present.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbfx.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first (python -m dolfinx_b200.build). "
        "dolfinx_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

OK = 0
ERR_CUDA = 1
ERR_NOT_IN_SPARSITY = 3
ERR_NO_DEVICE = 6

# kernel ids (include/bfx.h)
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
K_ACTION_POISSON_P1_TET_L = 15
K_ACTION_POISSON_P2_TET_L = 16
K_L2NORM2_P1_TET_M = 17
K_AVG_MASS_P1_TRI_DS = 18
K_AVG_LOAD_P1_TRI_DS_L = 19
K_ONE_TRI_DS_M = 20
K_AVG2_COEFF_P1_TRI_DS_M = 21
K_COEFF2_P1_TRI_FACET_M = 22
K_LOAD_PROD_P1_TET_L = 23

USER_KERNEL_BASE = 1000  # ids of kernels registered by the caller (bfx_register_kernel, include/bfx_plugin.cuh)

ASM_ATOMIC, ASM_CHUNKED, ASM_ROWGATHER = 0, 2, 3
ROWGATHER_KERNELS = frozenset({10})
# linear-form kernels with the grouped (one RED per distinct dof of 32 cells) variant
GROUPED_VECTOR_KERNELS = frozenset({1, 3, 7, 15})
CHUNKED_VECTOR_KERNELS = frozenset({1, 3, 7, 9, 11, 15, 16})
# where the chunk plan measures faster than the RED kernel: P2 tetrahedra (chunk-aggregated kernel) and the P1-sized
# kernels (table kernel k_vector_tables on the plan's warp tables: 1.95 against 2.08 ms at C2)
CHUNKED_VECTOR_DEFAULT = frozenset({1, 3, 7, 9, 15, 16})
ERR_UNSUPPORTED = 4
# bilinear kernels whose DEFAULT strategy is the chunk-aggregated variant (csrc/chunked.cu): the P1 kernels,
# and P2 Poisson (with the symmetric plan: 55 staged entries per cell, 128 cells per chunk)
CHUNKED_KERNELS = frozenset({0, 2, 6, 8})
LEAN_KERNELS = frozenset({0, 6})  # symmetric P1-sized kernels without coefficients: k_matrix_lean
CHUNKS_SHARED_MATRIX = 4
CHUNKS_LINEAR_STAGING = 2  # padded linear staging layout (no bank colouring); needed by CHUNK_KERNEL_LEAN
CHUNK_KERNEL_DEFAULT, CHUNK_KERNEL_OCC5, CHUNK_KERNEL_DIET, CHUNK_KERNEL_LEAN, CHUNK_KERNEL_WIDE = 0, 1, 2, 3, 4
CHUNKS_BANK_ORDER = 64  # bank-aware order of the source lists (linear staging layout)
CHUNKS_LEN_SORT = 128  # destinations ordered by list length only (lean kernel)
CHUNKS_PAD4 = 32  # source lists padded to multiples of 4 entries
CHUNKS_LEAN_ONLY = 262144  # fail early (ERR_UNSUPPORTED) instead of building a non-lean plan
CHUNKS_CB_SOFT = 131072  # the chunk size is the lean kernel's preference: default size if the lean options do not apply
CHUNKS_TWO_STAGE_SPLIT = 16
CHUNKS_TWO_STAGE = 8  # write-back in address order through shared memory (symmetric P1 plans)


def CHUNKS_CB(cells):
    """BFX_CHUNKS_CB(cells): cells per chunk other than the element's default (flags of bfx_asm_build_chunks)."""
    return ((cells // 32) & 0xFF) << 8

CHUNKS_SYMMETRIC = 1  # every chunked kernel (ids 0, 2, 6, 8) has a symmetric element matrix
VALUES_ADD, VALUES_OVERWRITE = 0, 1
SPMV_FULL, SPMV_DIAG, SPMV_OFFDIAG = 0, 1, 2


class BfxError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


class KernelInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nx", "nd", "bs", "rank", "w_size", "c_size", "facet")]


class CoeffSrc(C.Structure):
    _fields_ = [("values_dev", C.c_void_p), ("dofmap_dev", C.c_void_p), ("nd", C.c_int), ("bs", C.c_int), ("offset", C.c_int)]


class Coeffs(C.Structure):
    _fields_ = [("packed_dev", C.c_void_p), ("cstride", C.c_int), ("n_fused", C.c_int), ("fused", CoeffSrc * 4)]


vp, i32, i64, f64, ci = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_int
pvp = C.POINTER(C.c_void_p)

_SIGS = {
    "bfx_version": ([], ci),
    "bfx_device_count": ([C.POINTER(ci)], ci),
    "bfx_set_device": ([ci], ci),
    "bfx_malloc": ([pvp, C.c_size_t], ci),
    "bfx_free": ([vp], ci),
    "bfx_memcpy": ([vp, vp, C.c_size_t, vp], ci),
    "bfx_memset": ([vp, ci, C.c_size_t, vp], ci),
    "bfx_stream_sync": ([vp], ci),
    "bfx_host_alloc": ([pvp, C.c_size_t], ci),
    "bfx_host_free": ([vp], ci),
    "bfx_kernel_info": ([ci, C.POINTER(KernelInfo)], ci),
    "bfx_register_kernel": ([ci, C.POINTER(KernelInfo), vp], ci),
    "bfx_csr_create": ([pvp, i32, i32, vp, vp, vp, ci, ci], ci),
    "bfx_csr_destroy": ([vp], ci),
    "bfx_csr_nnz": ([vp], i64),
    "bfx_csr_get_structure": ([vp, vp, vp, vp], ci),
    "bfx_csr_device_ptrs": ([vp, pvp, pvp, pvp], ci),
    "bfx_sparsity_build": ([pvp, i32, i32, i32, vp, ci, vp, ci, vp, i64, vp, vp, i64, ci, ci, vp], ci),
    "bfx_sparsity_ghost_rows": ([i32, i32, vp, ci, vp, ci, vp, i64, vp, vp, vp], ci),
    "bfx_csr_insert": ([vp, vp, ci, ci, ci, vp, vp, ci, vp, ci, ci, vp], ci),
    "bfx_csr_set_diagonal": ([vp, vp, vp, i64, f64, vp], ci),
    "bfx_csr_squared_norm": ([vp, vp, C.POINTER(f64), vp], ci),
    "bfx_spmv": ([vp, vp, vp, vp, ci, vp], ci),
    "bfx_csr_set_spmv_variant": ([vp, ci], ci),
    "bfx_spmvT": ([vp, vp, vp, vp, ci, vp], ci),
    "bfx_asm_create": ([pvp, vp, vp, ci, vp, ci, vp, ci, i64, vp, i64, i32, ci, vp], ci),
    "bfx_asm_destroy": ([vp], ci),
    "bfx_asm_build_chunks": ([vp, vp, ci, vp], ci),
    "bfx_asm_build_rowgather": ([vp, vp], ci),
    "bfx_asm_build_groups": ([vp, vp, vp], ci),
    "bfx_asm_build_chunks_vector": ([vp, vp, ci, vp], ci),
    "bfx_asm_chunk_bank_conflicts": ([vp, C.POINTER(i64)], ci),
    "bfx_asm_chunk_two_stage": ([vp, C.POINTER(ci)], ci),
    "bfx_asm_chunk_set_kernel": ([vp, ci], ci),
    "bfx_csr_transpose_local": ([vp, vp, i32, vp, vp, vp, i64, C.POINTER(i64), vp], ci),
    "bfx_csr_matmul_begin": ([vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, C.POINTER(vp), C.POINTER(i64), vp], ci),
    "bfx_csr_matmul_end": ([vp, vp, vp, vp, vp, vp], ci),
    "bfx_asm_chunk_stats": ([vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)], ci),
    "bfx_assemble_matrix_cells": ([vp, ci, vp, vp, vp, C.POINTER(Coeffs), C.POINTER(f64), ci, vp, ci, ci, vp], ci),
    "bfx_assemble_vector_cells": ([vp, ci, vp, C.POINTER(Coeffs), C.POINTER(f64), ci, vp, ci, vp], ci),
    "bfx_assemble_scalar_cells": ([vp, ci, vp, C.POINTER(Coeffs), C.POINTER(f64), ci, C.POINTER(f64), vp], ci),
    "bfx_asm_chunk_partition": ([vp, i32, C.POINTER(i64)], ci),
    "bfx_assemble_matrix_cells_part": ([vp, ci, vp, vp, vp, C.POINTER(Coeffs), C.POINTER(f64), ci, vp, ci, ci, vp], ci),
    "bfx_assemble_matrix_rows": ([vp, ci, vp, vp, vp, C.POINTER(Coeffs), C.POINTER(f64), ci, vp, i32, i32, ci, vp], ci),
    "bfx_asm_rowgather_tile_rows": ([vp, C.POINTER(ci)], ci),
    "bfx_assemble_scalar_facets": ([vp, ci, vp, vp, i64, C.POINTER(Coeffs), C.POINTER(f64), ci, C.POINTER(f64), vp], ci),
    "bfx_lift_bc_cells": ([vp, ci, vp, C.POINTER(Coeffs), C.POINTER(f64), ci, vp, vp, vp, vp, f64, vp], ci),
    "bfx_assemble_matrix_facets": ([vp, ci, vp, vp, i64, vp, vp, C.POINTER(Coeffs), C.POINTER(f64), ci, vp, vp], ci),
    "bfx_assemble_vector_facets": ([vp, ci, vp, vp, i64, C.POINTER(Coeffs), C.POINTER(f64), ci, vp, vp], ci),
    "bfx_pack_coefficient": ([vp, ci, ci, vp, vp, ci, ci, vp, vp, i64, vp], ci),
    "bfx_assemble_matrix_cells_host": ([vp, ci, vp, i64, vp, vp, i64, vp, i64, ci, C.POINTER(f64), ci, vp, ci, vp], ci),
    "bfx_assemble_matrix_cells_host_begin": ([vp, ci, vp, i64, vp, vp, i64, vp, i64, ci, C.POINTER(f64), ci, vp, ci], ci),
    "bfx_assemble_matrix_cells_host_end": ([vp], ci),
    "bfx_bc_mark": ([vp, vp, i64, vp], ci),
    "bfx_bc_set": ([vp, i32, vp, vp, i64, vp, ci, ci, vp, f64, vp], ci),
    "bfx_dot": ([i64, vp, vp, C.POINTER(f64), vp], ci),
    "bfx_norm": ([i64, vp, ci, C.POINTER(f64), vp], ci),
    "bfx_axpy": ([i64, f64, vp, vp, vp], ci),
    "bfx_fill": ([i64, f64, vp, vp], ci),
    "bfx_comm_unique_id": ([C.c_char_p], ci),
    "bfx_comm_create": ([pvp, C.c_char_p, ci, ci], ci),
    "bfx_comm_destroy": ([vp], ci),
    "bfx_comm_rank": ([vp, C.POINTER(ci), C.POINTER(ci)], ci),
    "bfx_comm_allreduce": ([vp, vp, i64, ci, vp], ci),
    "bfx_scatter_create": ([pvp, vp, vp, i64, vp, i64, vp, vp, vp, ci, vp, vp, vp, ci], ci),
    "bfx_scatter_destroy": ([vp], ci),
    "bfx_scatter_fwd_begin": ([vp, vp, vp], ci),
    "bfx_scatter_fwd_end": ([vp, vp, i64, vp], ci),
    "bfx_scatter_rev_begin": ([vp, vp, i64, vp], ci),
    "bfx_scatter_rev_end": ([vp, vp, ci, vp], ci),
    "bfx_csr_scatter_create": ([pvp, vp, vp, vp, i32, vp, vp, ci, vp, vp, ci, vp], ci),
    "bfx_csr_scatter_destroy": ([vp], ci),
    "bfx_csr_scatter_rev_begin": ([vp, vp, vp], ci),
    "bfx_csr_scatter_rev_end": ([vp, vp, vp], ci),
    "bfx_host_first_touch_i32": ([vp, i64, i32, vp], ci),
    "bfx_host_first_touch_i64": ([vp, i64, i64, vp], ci),
}

EXPORTS = sorted(_SIGS) + ["bfx_last_error", "bfx_status_string"]

for _name, (_args, _res) in _SIGS.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = _res
lib.bfx_last_error.restype = C.c_char_p
lib.bfx_status_string.restype = C.c_char_p
lib.bfx_status_string.argtypes = [ci]


def check(status: int):
    if status != OK:
        msg = lib.bfx_last_error().decode(errors="replace")
        raise BfxError(status, msg or lib.bfx_status_string(status).decode())


def np_ptr(a: np.ndarray):
    """Host pointer of a contiguous numpy array (kept alive by the caller)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def dptr(t):
    """Device (or host) pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return np_ptr(t)
    assert t.is_contiguous()
    return t.data_ptr()


def current_stream():
    import torch

    return torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else None


def host_first_touch(flat: np.ndarray, ndofs: int) -> np.ndarray:
    flat = np.ascontiguousarray(flat)
    if flat.dtype == np.int32:
        out = np.empty(ndofs, dtype=np.int32)
        check(lib.bfx_host_first_touch_i32(flat.ctypes.data, flat.size, ndofs, out.ctypes.data))
    else:
        flat = flat.astype(np.int64, copy=False)
        out = np.empty(ndofs, dtype=np.int64)
        check(lib.bfx_host_first_touch_i64(flat.ctypes.data, flat.size, ndofs, out.ctypes.data))
    return out


_plugins = []  # (CDLL, launcher) pairs kept alive for the lifetime of the process


def register_plugin_kernel(kernel_id: int, library_path: str, symbol: str) -> KernelInfo:
    """Register a caller-built kernel (include/bfx_plugin.cuh: ``BFX_PLUGIN_KERNEL(E, symbol)``) under ``kernel_id``
    (>= USER_KERNEL_BASE): the device-side replacement of plugging a tabulate_tensor pointer into a fem::Form."""
    plug = C.CDLL(library_path)
    info = KernelInfo()
    getattr(plug, symbol + "_info")(C.byref(info))
    fn = getattr(plug, symbol)
    check(lib.bfx_register_kernel(kernel_id, C.byref(info), C.cast(fn, C.c_void_p)))
    _plugins.append((plug, fn))
    return info


def kernel_info(kernel_id: int) -> KernelInfo:
    ki = KernelInfo()
    check(lib.bfx_kernel_info(kernel_id, C.byref(ki)))
    return ki


def make_coeffs(packed=None, cstride=0, fused=None, offset=0):
    """Build a bfx_coeffs_t.  ``fused`` = (values tensor, dofmap tensor, nd, bs), or a list of such tuples for a kernel
    that gathers several coefficients (in the order of the kernel's w)."""
    c = Coeffs()
    if packed is not None:
        c.packed_dev = dptr(packed)
        c.cstride = cstride
        c.n_fused = 1
        c.fused[0].offset = offset
    elif fused is not None:
        items = [fused] if isinstance(fused, tuple) else list(fused)
        assert 1 <= len(items) <= 4
        c.n_fused = len(items)
        for k, (v, dm, nd, bs) in enumerate(items):
            c.fused[k].values_dev = dptr(v)
            c.fused[k].dofmap_dev = dptr(dm)
            c.fused[k].nd = nd
            c.fused[k].bs = bs
            c.fused[k].offset = offset
    return c


def constants_array(constants):
    if constants is None or len(constants) == 0:
        return None, 0
    arr = (f64 * len(constants))(*[float(v) for v in constants])
    return arr, len(constants)
