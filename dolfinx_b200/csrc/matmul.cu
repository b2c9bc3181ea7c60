// la::matmul on the device: C = A B for block size 1, la/matmul.h:395-536.
//
// One thread per row of A runs bfx::matmul_row (matmul_row.h: the reference's order of additions, zero products and
// exact cancellations leave no entry - bitwise the reference's row; the same header is compiled for the host and
// checked against the oracle by tests/test_matmul_row.py).  Three launches: (1) per-row upper bound of the candidates,
// exclusive scan -> workspace offsets; (2) the rows into the workspace, exclusive scan of the kept counts -> row
// pointer of C; (3) compaction into the caller's arrays.  The workspace holds one (column, value) per PRODUCT, which
// is fine for the Galerkin-type products this is meant for and too much for C2-size A A (documented in DESIGN.md).
#include "csr.cuh"
#include "matmul_row.h"
#include <cub/device/device_scan.cuh>

using namespace bfx;

struct bfx_matmul
{
  int32_t n_rows = 0;
  int64_t nnz = 0;
  int64_t *wofs = nullptr, *row_ptr = nullptr; // n_rows + 1 each
  int64_t* cnt = nullptr;                      // n_rows + 1 (kept entries per row, then scanned into row_ptr)
  int32_t *wcols = nullptr, *off_diag = nullptr;
  double* wvals = nullptr;
};

namespace
{
__global__ void k_matmul_bound(MatmulArgs m, int32_t n_rows, int64_t* __restrict__ ub)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_rows; i += (int64_t)gridDim.x * blockDim.x)
    ub[i] = i < n_rows ? matmul_row_bound(m, (int32_t)i) : 0;
}

__global__ void k_matmul_rows(MatmulArgs m, int32_t n_rows, const int64_t* __restrict__ wofs, int32_t* __restrict__ wcols,
                              double* __restrict__ wvals, int64_t* __restrict__ cnt, int32_t* __restrict__ off_diag)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_rows; i += (int64_t)gridDim.x * blockDim.x)
  {
    if (i == n_rows)
    {
      cnt[i] = 0;
      continue;
    }
    int32_t od = 0;
    cnt[i] = matmul_row(m, (int32_t)i, wcols + wofs[i], wvals + wofs[i], &od);
    off_diag[i] = od;
  }
}

// one warp per row: workspace segment -> C
__global__ void k_matmul_compact(int32_t n_rows, const int64_t* __restrict__ wofs, const int64_t* __restrict__ row_ptr,
                                 const int32_t* __restrict__ wcols, const double* __restrict__ wvals,
                                 int32_t* __restrict__ cols, double* __restrict__ vals)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n_rows; i += nwarps)
  {
    const int64_t src = wofs[i], dst = row_ptr[i], n = row_ptr[i + 1] - dst;
    for (int64_t q = lane; q < n; q += 32)
    {
      cols[dst + q] = wcols[src + q];
      vals[dst + q] = wvals[src + q];
    }
  }
}

void free_matmul(bfx_matmul* h)
{
  if (!h)
    return;
  cudaFree(h->wofs);
  cudaFree(h->row_ptr);
  cudaFree(h->cnt);
  cudaFree(h->wcols);
  cudaFree(h->off_diag);
  cudaFree(h->wvals);
  delete h;
}
} // namespace

extern "C"
{
int bfx_csr_matmul_begin(const bfx_csr_t* A, const double* a_values, const bfx_csr_t* B, const double* b_values,
                         int32_t n_owned_cols_b, const int32_t* b_ghost_remap, const int64_t* ghost_row_ptr,
                         const int32_t* ghost_cols, const double* ghost_vals, int32_t n_owned_cols_c, bfx_matmul_t** out,
                         int64_t* nnz_host, bfx_stream_t stream)
{
  BFX_REQUIRE(A && B && a_values && b_values && out && nnz_host && ghost_row_ptr, "bfx_csr_matmul_begin: null argument");
  if (A->bs0 != 1 || A->bs1 != 1 || B->bs0 != 1 || B->bs1 != 1)
    return fail(BFX_ERR_UNSUPPORTED, "Currently matmul only supports block size=1"); // la/matmul.h:549-552
  cudaStream_t st = S(stream);
  *out = nullptr;
  *nnz_host = 0;
  bfx_matmul* h = new bfx_matmul();
  h->n_rows = A->n_rows_owned;
  const size_t n1 = (size_t)h->n_rows + 1;
  MatmulArgs m{A->row_ptr, A->off_diag, A->cols,       a_values,      B->row_ptr, B->cols,    b_values,
               B->n_rows_owned, n_owned_cols_b, b_ghost_remap, ghost_row_ptr, ghost_cols, ghost_vals, n_owned_cols_c};
  int e = BFX_OK;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  int64_t total = 0;
  auto bail = [&](int status)
  {
    cudaFree(tmp);
    free_matmul(h);
    return status;
  };
  if ((e = dev_alloc(&h->wofs, n1)) || (e = dev_alloc(&h->row_ptr, n1)) || (e = dev_alloc(&h->cnt, n1))
      || (e = dev_alloc(&h->off_diag, n1)))
    return bail(e);
  k_matmul_bound<<<grid_for((int64_t)n1, 256, 16), 256, 0, st>>>(m, h->n_rows, h->cnt);
  cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess)
    ce = cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->cnt, h->wofs, (int)n1, st);
  if (ce == cudaSuccess)
    ce = cudaMalloc(&tmp, tmp_bytes);
  if (ce == cudaSuccess)
    ce = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, h->cnt, h->wofs, (int)n1, st);
  if (ce == cudaSuccess)
    ce = cudaMemcpyAsync(&total, h->wofs + h->n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess)
    ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess)
    return bail(fail(BFX_ERR_CUDA, "bfx_csr_matmul_begin: %s", cudaGetErrorString(ce)));
  if ((e = dev_alloc(&h->wcols, (size_t)total + 1)) || (e = dev_alloc(&h->wvals, (size_t)total + 1)))
    return bail(e);
  k_matmul_rows<<<grid_for((int64_t)n1, 128, 32), 128, 0, st>>>(m, h->n_rows, h->wofs, h->wcols, h->wvals, h->cnt, h->off_diag);
  ce = cudaGetLastError();
  if (ce == cudaSuccess)
    ce = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, h->cnt, h->row_ptr, (int)n1, st);
  if (ce == cudaSuccess)
    ce = cudaMemcpyAsync(&h->nnz, h->row_ptr + h->n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess)
    ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess)
    return bail(fail(BFX_ERR_CUDA, "bfx_csr_matmul_begin: %s", cudaGetErrorString(ce)));
  cudaFree(tmp);
  *nnz_host = h->nnz;
  *out = h;
  return BFX_OK;
}

int bfx_csr_matmul_end(bfx_matmul_t* h, int64_t* row_ptr, int32_t* off_diag, int32_t* cols, double* vals, bfx_stream_t stream)
{
  BFX_REQUIRE(h, "bfx_csr_matmul_end: null handle");
  cudaStream_t st = S(stream);
  cudaError_t ce = cudaSuccess;
  if (row_ptr)
    ce = cudaMemcpyAsync(row_ptr, h->row_ptr, sizeof(int64_t) * ((size_t)h->n_rows + 1), cudaMemcpyDeviceToDevice, st);
  if (ce == cudaSuccess && off_diag && h->n_rows)
    ce = cudaMemcpyAsync(off_diag, h->off_diag, sizeof(int32_t) * (size_t)h->n_rows, cudaMemcpyDeviceToDevice, st);
  if (ce == cudaSuccess && cols && vals && h->nnz > 0)
  {
    k_matmul_compact<<<grid_for((int64_t)h->n_rows * 32, 256, 16), 256, 0, st>>>(h->n_rows, h->wofs, h->row_ptr, h->wcols,
                                                                              h->wvals, cols, vals);
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess)
    ce = cudaStreamSynchronize(st);
  free_matmul(h);
  if (ce != cudaSuccess)
    return fail(BFX_ERR_CUDA, "bfx_csr_matmul_end: %s", cudaGetErrorString(ce));
  return BFX_OK;
}
}
