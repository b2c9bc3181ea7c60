// ORACLE SUPPORT — TEST INFRASTRUCTURE ONLY.
// Thin extern "C" shim over the REFERENCE's own la/matrix_csr_impl.h, compiled in place from
// /root/reference (never copied).  It exists only in the build container (oracle/_ref/ is
// git-ignored); tests use it to differential-test oracle.c's restated insert_csr*/spmv*.
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <iterator>
#include <stdexcept>
#include <dolfinx/la/matrix_csr_impl.h>
#include <span>
#include <vector>

using namespace dolfinx::la::impl;

template <int BS0, int BS1>
static int ins(int kind, double* data, std::size_t ndata, const int32_t* cols, std::size_t ncols,
               const int64_t* row_ptr, std::size_t nrp, const double* x, const int32_t* xrows, int nr,
               const int32_t* xcols, int nc, int op)
{
  std::span<double> d(data, ndata);
  std::span<const int32_t> c(cols, ncols);
  std::span<const int64_t> rp(row_ptr, nrp);
  std::span<const int32_t> xr(xrows, nr), xc(xcols, nc);
  try
  {
    if (kind == 0)
    {
      std::span<const double> xs(x, std::size_t(nr) * nc * BS0 * BS1);
      if (op)
        insert_csr<BS0, BS1>(d, c, rp, xs, xr, xc, [](double& a, double b) { a += b; }, int32_t(nrp - 1));
      else
        insert_csr<BS0, BS1>(d, c, rp, xs, xr, xc, [](double& a, double b) { a = b; }, int32_t(nrp - 1));
    }
    else
    {
      std::span<const double> xs(x, std::size_t(nr) * nc * BS0 * BS1);
      if (op)
        insert_blocked_csr<BS0, BS1>(d, c, rp, xs, xr, xc, [](double& a, double b) { a += b; }, int32_t(nrp - 1));
      else
        insert_blocked_csr<BS0, BS1>(d, c, rp, xs, xr, xc, [](double& a, double b) { a = b; }, int32_t(nrp - 1));
    }
  }
  catch (const std::runtime_error&)
  {
    return -1;
  }
  return 0;
}

extern "C"
{
int ref_insert(int kind, int bs0, int bs1, double* data, std::size_t ndata, const int32_t* cols,
               std::size_t ncols, const int64_t* row_ptr, std::size_t nrp, const double* x, const int32_t* xrows,
               int nr, const int32_t* xcols, int nc, int op)
{
#define CASE(a, b)                                                                                                    \
  if (bs0 == a && bs1 == b)                                                                                           \
    return ins<a, b>(kind, data, ndata, cols, ncols, row_ptr, nrp, x, xrows, nr, xcols, nc, op);
  CASE(1, 1) CASE(2, 2) CASE(3, 3) CASE(1, 2) CASE(2, 1) CASE(2, 3) CASE(3, 2)
#undef CASE
  return -3;
}

int ref_insert_nonblocked(int bs0, int bs1, double* data, std::size_t ndata, const int32_t* cols, std::size_t ncols,
                          const int64_t* row_ptr, std::size_t nrp, const double* x, const int32_t* xrows, int nr,
                          const int32_t* xcols, int nc, int op)
{
  std::span<double> d(data, ndata);
  std::span<const int32_t> c(cols, ncols);
  std::span<const int64_t> rp(row_ptr, nrp);
  std::span<const int32_t> xr(xrows, nr), xc(xcols, nc);
  std::span<const double> xs(x, std::size_t(nr) * nc);
  try
  {
    if (op)
      insert_nonblocked_csr(d, c, rp, xs, xr, xc, [](double& a, double b) { a += b; }, int32_t(nrp - 1), bs0, bs1);
    else
      insert_nonblocked_csr(d, c, rp, xs, xr, xc, [](double& a, double b) { a = b; }, int32_t(nrp - 1), bs0, bs1);
  }
  catch (const std::runtime_error&)
  {
    return -1;
  }
  return 0;
}

void ref_spmv(int transpose, const double* values, std::size_t nvals, const int64_t* row_begin,
              const int64_t* row_end, std::size_t nrows, const int32_t* indices, std::size_t nidx, const double* x,
              std::size_t nx, double* y, std::size_t ny, int bs0, int bs1)
{
  std::span<const double> v(values, nvals), xs(x, nx);
  std::span<const int64_t> rb(row_begin, nrows), re(row_end, nrows);
  std::span<const int32_t> idx(indices, nidx);
  std::span<double> ys(y, ny);
  if (transpose)
    spmvT<double>(v, rb, re, idx, xs, ys, bs0, bs1);
  else
    spmv<double>(v, rb, re, idx, xs, ys, bs0, bs1);
}
}
