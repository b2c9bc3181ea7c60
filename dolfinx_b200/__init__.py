"""dolfinx_b200 — B200-native (sm_100a CUDA + NCCL) drop-in for DOLFINx's assembly hot path.

assembly (fem::assemble_matrix / assemble_vector / apply_lifting / DirichletBC) ->
la::MatrixCSR (insert + SpMV) -> la::Vector / common::Scatterer ghost exchange.

Layout:
  csrc/        CUDA kernels + the C-ABI (libbfx.so, declared in include/bfx.h)
  _lib.py      ctypes binding of the C-ABI (fails loudly without the library / a GPU)
  common.py    Comm, IndexMap, Scatterer     (cpp/dolfinx/common)
  la.py        SparsityPattern, MatrixCSR, Vector (cpp/dolfinx/la)
  fem.py       Form, DirichletBC, assemble_*, apply_lifting, set_diagonal (cpp/dolfinx/fem)
  mesh.py      synthetic box-mesh fixtures (inputs of the path)
  cpp/         header-only C++ mirror of the same classes over the C-ABI
"""

__version__ = "0.1.0"
