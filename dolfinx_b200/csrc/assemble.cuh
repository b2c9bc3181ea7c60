// Assembly plan and kernel argument block.
#pragma once
#include "common.cuh"
#include "csr.cuh"

namespace bfx
{
struct CoefArgs
{
  const double* packed; // (n_entities, cstride) reference layout, or NULL
  int cstride;
  struct
  {
    const double* v;   // coefficient dof vector
    const int32_t* dm; // its dofmap (ncells_all x WND)
    int off;           // offset of this coefficient inside w (packed layout)
  } f[1];
};

struct AsmArgs
{
  const int32_t *x_dofmap, *dofmap0, *dofmap1;
  const int32_t* cells;    // cell list or NULL (= identity)
  const int32_t* entities; // (cell, local_facet) pairs or NULL
  int64_t n;               // number of cells / entities
  const double* x;         // geometry, (N,3) row-major
  const int8_t *bc0, *bc1; // Dirichlet markers or NULL
  CoefArgs coef;
  double constants[8];
  // matrix sink
  double* values;
  const int64_t* row_ptr;
  const int32_t* cols;
  const void* pos; // cell -> position-in-row map, or NULL (binary search)
  // vector sink / lifting
  double* b;
  const double* bc_values1;
  const double* x0;
  double alpha;
  int* err;
};
} // namespace bfx

struct bfx_asm
{
  const bfx_csr* csr = nullptr;
  int nx = 0, nd0 = 0, nd1 = 0;
  int64_t ncells_all = 0, ncells = 0;
  int32_t n_rows_all = 0;
  bool owns = true;
  int32_t *x_dofmap = nullptr, *dofmap0 = nullptr, *dofmap1 = nullptr, *cells = nullptr;
  // cell -> CSR position map: per cell nd0*nd1 offsets relative to row_ptr[row], uint8 or uint16
  char* pos = nullptr;
  int pos_bytes = 1, pos_stride = 0;
  // scratch of the host-buffer entry point
  double *h_x = nullptr, *h_coeff = nullptr, *h_values = nullptr;
  int8_t *h_bc0 = nullptr, *h_bc1 = nullptr;
  int64_t h_x_n = 0, h_coeff_n = 0, h_bc_n = 0;
};
