// One row of C = A B, block size 1 - the body of impl::matmul (la/matmul.h:443-524) without its dense accumulator.
//
// The reference keeps a dense accumulator over all columns of C and a list of touched columns; its result depends on
// the ORDER of the additions and on exact zeros: a product that is exactly zero does not create an entry (:457, :477),
// a column whose sum cancels to exactly zero is removed (:486-497), the kept columns are sorted (:501).  The routine
// below produces the same row - same structure, bitwise the same values - in a workspace of the row's candidates:
//   1. the distinct columns with a non-zero product, collected in first-touch order, then sorted;
//   2. the products added to their column in the reference's order (entries of row i of A in storage order, for each
//      the entries of the row of B in storage order; rows of B behind ghost columns of A come from the fetched ghost
//      rows) - a product that is zero changes nothing and is skipped;
//   3. columns whose sum is exactly zero dropped.
// It is compiled for the device (matmul.cu: one thread per row) AND for the host (tests/cpp/matmul_row_host.cpp), so
// that the CPU suite checks the very code the GPU runs against the oracle.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define BFX_HD __host__ __device__ __forceinline__
#else
#define BFX_HD inline
#endif

namespace bfx
{
struct MatmulArgs
{
  // A: owned rows; entries [row_ptr, off_diag) have owned columns (= owned rows of B), [off_diag, row_ptr[i+1]) ghost ones
  const int64_t* a_row_ptr;
  const int64_t* a_off_diag;
  const int32_t* a_cols;
  const double* a_vals;
  // B: all local rows; columns < n_owned_cols_b keep their index in C, ghost columns go through b_ghost_remap
  const int64_t* b_row_ptr;
  const int32_t* b_cols;
  const double* b_vals;
  int32_t n_rows_b; // owned rows of B: ghost column j of A is ghost row j - n_rows_b
  int32_t n_owned_cols_b;
  const int32_t* b_ghost_remap;
  // fetched ghost rows of B, columns already in the numbering of C (la/matmul.h:79-390)
  const int64_t* g_row_ptr;
  const int32_t* g_cols;
  const double* g_vals;
  int32_t n_owned_cols_c;
};

/// Upper bound of the entries of row i of C: one candidate per product
BFX_HD int64_t matmul_row_bound(const MatmulArgs& m, int32_t i)
{
  int64_t n = 0;
  for (int64_t ka = m.a_row_ptr[i]; ka < m.a_off_diag[i]; ++ka)
  {
    const int32_t j = m.a_cols[ka];
    n += m.b_row_ptr[j + 1] - m.b_row_ptr[j];
  }
  for (int64_t ka = m.a_off_diag[i]; ka < m.a_row_ptr[i + 1]; ++ka)
  {
    const int32_t g = m.a_cols[ka] - m.n_rows_b;
    n += m.g_row_ptr[g + 1] - m.g_row_ptr[g];
  }
  return n;
}

/// Visit the products of row i in the reference's order: f(column of C, a * b)
template <class F>
BFX_HD void matmul_row_products(const MatmulArgs& m, int32_t i, F&& f)
{
  for (int64_t ka = m.a_row_ptr[i]; ka < m.a_off_diag[i]; ++ka) // la/matmul.h:449-466
  {
    const int32_t j = m.a_cols[ka];
    const double a = m.a_vals[ka];
    for (int64_t kb = m.b_row_ptr[j]; kb < m.b_row_ptr[j + 1]; ++kb)
    {
      const int32_t c = m.b_cols[kb];
      const int32_t k = c < m.n_owned_cols_b ? c : m.b_ghost_remap[c - m.n_owned_cols_b];
      f(k, a * m.b_vals[kb]);
    }
  }
  for (int64_t ka = m.a_off_diag[i]; ka < m.a_row_ptr[i + 1]; ++ka) // :469-484
  {
    const int32_t g = m.a_cols[ka] - m.n_rows_b;
    const double a = m.a_vals[ka];
    for (int64_t kb = m.g_row_ptr[g]; kb < m.g_row_ptr[g + 1]; ++kb)
      f(m.g_cols[kb], a * m.g_vals[kb]);
  }
}

/// Row i of C into the workspace (room for matmul_row_bound(i) entries): returns the number of entries kept, sorted
/// by column in wcols / wvals[0, n); *off_diag = entries with an owned column (la/matmul.h:505-507).
BFX_HD int32_t matmul_row(const MatmulArgs& m, int32_t i, int32_t* wcols, double* wvals, int32_t* off_diag)
{
  // 1. distinct columns with a non-zero product (first-touch order), then sorted
  int32_t n = 0;
  matmul_row_products(m, i,
                      [&](int32_t k, double v)
                      {
                        if (v == 0.0)
                          return;
                        for (int32_t q = 0; q < n; ++q)
                          if (wcols[q] == k)
                            return;
                        wcols[n++] = k;
                      });
  for (int32_t q = 1; q < n; ++q) // insertion sort: rows are short
  {
    const int32_t k = wcols[q];
    int32_t p = q - 1;
    while (p >= 0 && wcols[p] > k)
    {
      wcols[p + 1] = wcols[p];
      --p;
    }
    wcols[p + 1] = k;
  }
  for (int32_t q = 0; q < n; ++q)
    wvals[q] = 0.0;
  // 2. sums in the reference's order
  matmul_row_products(m, i,
                      [&](int32_t k, double v)
                      {
                        if (v == 0.0)
                          return;
                        int32_t lo = 0, hi = n;
                        while (lo < hi)
                        {
                          const int32_t mid = (lo + hi) >> 1;
                          if (wcols[mid] < k)
                            lo = mid + 1;
                          else
                            hi = mid;
                        }
                        wvals[lo] += v;
                      });
  // 3. exact cancellations leave no entry (:486-497); diagonal-block boundary (:505-507)
  int32_t kept = 0, od = 0;
  for (int32_t q = 0; q < n; ++q)
  {
    if (wvals[q] == 0.0)
      continue;
    wcols[kept] = wcols[q];
    wvals[kept] = wvals[q];
    od += wcols[q] < m.n_owned_cols_c ? 1 : 0;
    ++kept;
  }
  *off_diag = od;
  return kept;
}
} // namespace bfx
