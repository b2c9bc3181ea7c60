// Device-resident la::MatrixCSR structure (la/MatrixCSR.h:67-70: row_ptr int64, cols int32).
#pragma once
#include "common.cuh"

struct bfx_csr
{
  int32_t n_rows_all = 0;   // owned + ghost block rows
  int32_t n_rows_owned = 0; // MatrixCSR::num_owned_rows
  int bs0 = 1, bs1 = 1;
  int64_t nnz = 0;       // block entries, all rows
  int64_t nnz_owned = 0; // row_ptr[n_rows_owned]
  int64_t* row_ptr = nullptr;
  int32_t* cols = nullptr;
  int64_t* off_diag = nullptr;      // MatrixCSR::off_diag_offset (la/MatrixCSR.h:695-703)
  int32_t* offdiag_rows = nullptr;  // owned rows with at least one ghost column
  int32_t n_offdiag_rows = 0;
  int* err_flag = nullptr;          // device flag: "Entry not in sparsity"
  int spmv_variant = -1;            // SpMV kernel: -1 = by average row length on the first bfx_spmv call, -2 = by timing
};

namespace bfx
{
int csr_finish_create(bfx_csr* A);
}
