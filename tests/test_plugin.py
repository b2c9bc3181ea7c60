"""The kernel plug point (fem/kernel.h:18-20, fem/Form.h:76-78): forms libbfx.so does not ship,
written by the caller as element structs, compiled into the caller's own shared library against include/bfx_plugin.cuh
(tests/cpp/user_kernel_plugin.cu) and registered with bfx_register_kernel - no rebuild of libbfx.so.
CPU tier: the plugin compiles for sm_100a, its info entries describe the elements, registration validates its arguments.
GPU tier: the registered ids assemble a matrix (with Dirichlet rows / columns), lift it, assemble a vector and a
functional, against the oracle's quadrature versions of the same forms."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12


@pytest.fixture(scope="module")
def plugin(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("plug") / "libuser_kernels.so")
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-shared",
           "-Xcompiler", "-fPIC", os.path.join(ROOT, "tests", "cpp", "user_kernel_plugin.cu"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return so


def test_plugin_builds_and_registers(plugin):
    from dolfinx_b200 import _lib as K

    base = K.USER_KERNEL_BASE
    ki = K.register_plugin_kernel(base + 0, plugin, "plug_mass_p1_tet")
    assert (ki.nx, ki.nd, ki.bs, ki.rank, ki.w_size, ki.c_size, ki.facet) == (4, 4, 1, 2, 0, 0, 0)
    ki = K.register_plugin_kernel(base + 1, plugin, "plug_source_const_p1_tet")
    assert (ki.rank, ki.c_size) == (1, 1)
    ki = K.register_plugin_kernel(base + 2, plugin, "plug_volume_tet")
    assert ki.rank == 0
    assert K.kernel_info(base + 1).c_size == 1  # registered ids answer bfx_kernel_info like built-in ones
    # ids below the user range and null launchers are refused
    info = K.KernelInfo(4, 4, 1, 2, 0, 0, 0)
    assert K.lib.bfx_register_kernel(5, C.byref(info), C.c_void_p(1)) != K.OK
    assert K.lib.bfx_register_kernel(base + 7, C.byref(info), None) != K.OK
    with pytest.raises(K.BfxError):
        K.kernel_info(base + 99)


@pytest.mark.gpu
def test_registered_kernels_assemble_like_the_oracle(plugin, oracle):
    import torch

    from dolfinx_b200 import _lib as K
    from dolfinx_b200 import common, fem, la

    base = K.USER_KERNEL_BASE
    K.register_plugin_kernel(base + 0, plugin, "plug_mass_p1_tet")
    K.register_plugin_kernel(base + 1, plugin, "plug_source_const_p1_tet")
    K.register_plugin_kernel(base + 2, plugin, "plug_volume_tet")
    p = P.tet_p1(6, numbering="random")
    comm = common.COMM_SELF
    msh = fem.Mesh(comm, p.x, p.x_dofmap, p.cell)
    V = fem.FunctionSpace(msh, "Lagrange", fem.DofMap(p.dofmap, 1, common.IndexMap(comm, p.ndofs)))
    bdofs = np.flatnonzero(p.dof_coords[:, 2] < 1e-12).astype(np.int32)
    bc = fem.DirichletBC(fem.Constant(1.5), bdofs, V)
    mk = np.zeros(p.ndofs, dtype=np.int8)
    mk[bdofs] = 1
    # bilinear form through the registered id: matrix with bc rows / columns zeroed
    a = fem.Form([V, V], {fem.IntegralType.cell: [(0, base + 0, None, [])]})
    sp = fem.create_sparsity_pattern(a)
    sp.finalize()
    A = la.MatrixCSR(sp)
    fem.assemble_matrix(A, a, bcs=[bc])
    pat, ref = P.oracle_assemble_matrix(oracle, p, oracle.ORACLE_MASS_P1_TET_A, bc=mk)
    assert np.array_equal(A.indptr, pat.offsets) and np.array_equal(A.indices, pat.edges)
    assert P.row_scaled_error(A.data.cpu().numpy(), ref, pat.offsets) <= TOL
    # linear form + lifting through registered ids
    L = fem.Form([V], {fem.IntegralType.cell: [(0, base + 1, None, [])]}, constants=[fem.Constant(2.5)])
    b = la.Vector(V.dofmap.index_map, 1)
    fem.assemble_vector(b, L)
    refb = P.oracle_assemble_vector(oracle, p, oracle.ORACLE_SOURCE_CONST_P1_TET_L, constants=np.array([2.5]))
    assert np.max(np.abs(b.array.cpu().numpy() - refb)) <= TOL * np.max(np.abs(refb))
    assert float(b.array.sum()) == pytest.approx(2.5, rel=1e-12)
    fem.apply_lifting(b, [a], [[bc]])
    cells = np.arange(len(p.dofmap), dtype=np.int32)
    g = np.zeros(p.ndofs)
    g[bdofs] = 1.5
    oracle.lift_bc(oracle.ORACLE_MASS_P1_TET_A, p.x_dofmap, p.x, cells, p.dofmap, 1, p.dofmap, 1, refb, g, mk)
    assert np.max(np.abs(b.array.cpu().numpy() - refb)) <= TOL * np.max(np.abs(refb))
    # functional
    M = fem.Form([], {fem.IntegralType.cell: [(0, base + 2, None, [])]}, mesh=msh)
    assert fem.assemble_scalar(M) == pytest.approx(1.0, rel=1e-12)
    assert oracle.assemble_scalar(oracle.ORACLE_VOLUME_TET_M, p.x_dofmap, p.x, cells) == pytest.approx(1.0, rel=1e-12)
    # an unregistered id fails loudly
    bad = fem.Form([V], {fem.IntegralType.cell: [(0, base + 55, None, [])]})
    with pytest.raises(K.BfxError):
        fem.assemble_vector(la.Vector(V.dofmap.index_map, 1), bad)
    assert torch.cuda.is_available()
