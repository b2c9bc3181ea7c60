// Host build of dolfinx_b200/csrc/matmul_row.h: the per-row routine of the device SpGEMM, looped over the rows on the
// CPU so that tests/test_matmul_row.py can compare the very code the GPU runs with the oracle (bitwise).
#include "../../dolfinx_b200/csrc/matmul_row.h"
#include <vector>

extern "C" int matmul_rows_host(int32_t n_rows_a, const int64_t* a_row_ptr, const int64_t* a_off_diag, const int32_t* a_cols,
                                const double* a_vals, const int64_t* b_row_ptr, const int32_t* b_cols, const double* b_vals,
                                int32_t n_rows_b, int32_t n_owned_cols_b, const int32_t* b_ghost_remap,
                                const int64_t* g_row_ptr, const int32_t* g_cols, const double* g_vals,
                                int32_t n_owned_cols_c, int64_t capacity, int64_t* c_row_ptr, int32_t* c_off_diag,
                                int32_t* c_cols, double* c_vals)
{
  bfx::MatmulArgs m{a_row_ptr, a_off_diag, a_cols,        a_vals,    b_row_ptr, b_cols, b_vals,
                    n_rows_b,  n_owned_cols_b, b_ghost_remap, g_row_ptr, g_cols,    g_vals, n_owned_cols_c};
  int64_t at = 0;
  c_row_ptr[0] = 0;
  std::vector<int32_t> wc;
  std::vector<double> wv;
  for (int32_t i = 0; i < n_rows_a; ++i)
  {
    const int64_t ub = bfx::matmul_row_bound(m, i);
    wc.resize(ub + 1);
    wv.resize(ub + 1);
    int32_t od = 0;
    const int32_t n = bfx::matmul_row(m, i, wc.data(), wv.data(), &od);
    if (at + n > capacity)
      return 1;
    for (int32_t q = 0; q < n; ++q)
    {
      c_cols[at + q] = wc[q];
      c_vals[at + q] = wv[q];
    }
    at += n;
    c_off_diag[i] = od;
    c_row_ptr[i + 1] = at;
  }
  return 0;
}
