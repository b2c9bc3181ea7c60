import ctypes as C, numpy as np, sys
sys.path.insert(0,'.')
import torch
from dolfinx_b200 import _lib as K, common, fem, la
from tests import problems as P
for n in (12,16,24,32):
    p = P.tet_p1(n, numbering="first_touch")
    comm = common.COMM_SELF
    msh = fem.Mesh(comm, p.x, p.x_dofmap, p.cell)
    V = fem.FunctionSpace(msh, "P1", fem.DofMap(p.dofmap, 1, common.IndexMap(comm, p.ndofs)))
    a = fem.Form([V, V], {fem.IntegralType.cell: [(0, K.K_POISSON_P1_TET_A, None, [])]}, constants=[fem.Constant(2.0)])
    sp = fem.create_sparsity_pattern(a); sp.finalize(); A = la.MatrixCSR(sp)
    fem.assemble_matrix(A, a)
    plan = fem._asm_plan(a, a.integral(fem.IntegralType.cell,0), fem.IntegralType.cell, A)
    n1 = C.c_int64(-1)
    st = K.lib.bfx_asm_chunk_partition(plan, p.ndofs//2, C.byref(n1))
    print(n, 'status', st, K.lib.bfx_last_error() if st else '', n1.value, fem.chunk_stats(a, A))
