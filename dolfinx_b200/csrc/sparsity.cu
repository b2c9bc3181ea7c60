// Sparsity construction on device: fem::sparsitybuild::cells (fem/sparsitybuild.h:36-50) +
// the local part of la::SparsityPattern::finalize (la/SparsityPattern.cpp:438-478).
//
// The reference appends nd0*nd1 (row, col) pairs per cell to a COO cache (12.9 GB at 256^3 P1
// tets), buckets it by row, de-duplicates with a generation stamp and sorts every row.  Here the
// row -> incident-cells map (the transposed dofmap, cf. fem::transpose_dofmap, fem/DofMap.h:62-64)
// is built with one counting pass, and each row's sorted unique column list is produced by one
// thread from the dofmap rows of its incident cells — no COO cache at all.  The result is the
// same sorted, de-duplicated CSR graph, bit for bit.
#include "csr.cuh"
#include <cub/device/device_scan.cuh>
#include <vector>

using namespace bfx;

namespace
{
__global__ void k_count_incidence(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ dofmap,
                                  int nd, int32_t n_rows, int64_t* __restrict__ counts, int* __restrict__ err)
{
  const int64_t total = n * nd;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t e = t / nd;
    const int i = (int)(t - e * nd);
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    const int32_t r = dofmap[(int64_t)cell * nd + i];
    if (r < 0 || r >= n_rows)
    {
      *err = 2;
      continue;
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(counts + r), 1ULL);
  }
}

__global__ void k_fill_incidence(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ dofmap,
                                 int nd, int32_t n_rows, const int64_t* __restrict__ ptr, int32_t* __restrict__ cursor,
                                 int32_t* __restrict__ rc_cells)
{
  const int64_t total = n * nd;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t e = t / nd;
    const int i = (int)(t - e * nd);
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    const int32_t r = dofmap[(int64_t)cell * nd + i];
    if (r < 0 || r >= n_rows)
      continue;
    const int32_t k = atomicAdd(cursor + r, 1);
    rc_cells[ptr[r] + k] = cell;
  }
}

// ascending cell order inside every row (= the order the CPU loop visits them)
__global__ void k_sort_incidence(int32_t n_rows, const int64_t* __restrict__ ptr, int32_t* __restrict__ rc_cells)
{
  for (int32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x)
  {
    const int64_t b = ptr[r], e = ptr[r + 1];
    for (int64_t i = b + 1; i < e; ++i)
    {
      const int32_t v = rc_cells[i];
      int64_t j = i - 1;
      while (j >= b && rc_cells[j] > v)
      {
        rc_cells[j + 1] = rc_cells[j];
        --j;
      }
      rc_cells[j + 1] = v;
    }
  }
}

template <int CAP>
__device__ __forceinline__ bool insert_sorted(int32_t (&buf)[CAP], int& len, int32_t c)
{
  int lo = 0, hi = len;
  while (lo < hi)
  {
    const int mid = (lo + hi) >> 1;
    if (buf[mid] < c)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo < len && buf[lo] == c)
    return true;
  if (len >= CAP)
    return false;
  for (int k = len; k > lo; --k)
    buf[k] = buf[k - 1];
  buf[lo] = c;
  ++len;
  return true;
}

// One thread per row: union of the trial dofs of the incident cells (+ extra entries), sorted.
// FILL = false: write the count; FILL = true: write columns and the off-diagonal offset.
template <int CAP, bool FILL>
__global__ void __launch_bounds__(128)
    k_row_columns(int32_t n_rows, const int64_t* __restrict__ rc_ptr, const int32_t* __restrict__ rc_cells,
                  const int32_t* __restrict__ dofmap1, int nd1, const int64_t* __restrict__ ex_ptr,
                  const int32_t* __restrict__ ex_cols, int32_t n_cols_owned, int64_t* __restrict__ counts,
                  const int64_t* __restrict__ row_ptr, int32_t* __restrict__ cols, int64_t* __restrict__ off_diag,
                  int* __restrict__ overflow)
{
  for (int32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x)
  {
    int32_t buf[CAP];
    int len = 0;
    bool ok = true;
    for (int64_t k = rc_ptr[r]; k < rc_ptr[r + 1]; ++k)
    {
      const int32_t* d = dofmap1 + (int64_t)rc_cells[k] * nd1;
      for (int j = 0; j < nd1; ++j)
        ok &= insert_sorted<CAP>(buf, len, d[j]);
    }
    if (ex_ptr)
      for (int64_t k = ex_ptr[r]; k < ex_ptr[r + 1]; ++k)
        ok &= insert_sorted<CAP>(buf, len, ex_cols[k]);
    if (!ok)
    {
      *overflow = 1;
      continue;
    }
    if (!FILL)
      counts[r] = len;
    else
    {
      const int64_t b = row_ptr[r];
      int nd = 0;
      for (int k = 0; k < len; ++k)
      {
        cols[b + k] = buf[k];
        nd += buf[k] < n_cols_owned;
      }
      off_diag[r] = b + nd;
    }
  }
}

// Ghost rows: columns in first-occurrence (insertion) order, de-duplicated.
__global__ void k_ghost_row_columns(int32_t n_rows_owned, int32_t n_ghost, const int64_t* __restrict__ rc_ptr,
                                    const int32_t* __restrict__ rc_cells, const int32_t* __restrict__ dofmap1, int nd1,
                                    int64_t* __restrict__ counts, int32_t* __restrict__ out)
{
  for (int32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_ghost; g += gridDim.x * blockDim.x)
  {
    const int32_t r = n_rows_owned + g;
    const int64_t base = (rc_ptr[r] - rc_ptr[n_rows_owned]) * nd1;
    int32_t* o = out + base;
    int len = 0;
    for (int64_t k = rc_ptr[r]; k < rc_ptr[r + 1]; ++k)
    {
      const int32_t* d = dofmap1 + (int64_t)rc_cells[k] * nd1;
      for (int j = 0; j < nd1; ++j)
      {
        const int32_t c = d[j];
        bool seen = false;
        for (int q = 0; q < len && !seen; ++q)
          seen = o[q] == c;
        if (!seen)
          o[len++] = c;
      }
    }
    counts[g] = len;
  }
}

struct Transpose
{
  int64_t* ptr = nullptr;    // [n_rows + 1]
  int32_t* cells = nullptr;  // [total]
  int64_t total = 0;
  void release()
  {
    cudaFree(ptr);
    cudaFree(cells);
    ptr = nullptr;
    cells = nullptr;
  }
};

int exclusive_scan_i64(int64_t* d_in_out, int64_t n, cudaStream_t st)
{
  // in-place exclusive sum over n items (callers pass n_rows + 1 with a trailing zero)
  void* tmp = nullptr;
  size_t bytes = 0;
  BFX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_in_out, d_in_out, n, st));
  BFX_CUDA(cudaMalloc(&tmp, bytes ? bytes : 1));
  BFX_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, d_in_out, d_in_out, n, st));
  BFX_CUDA(cudaStreamSynchronize(st));
  BFX_CUDA(cudaFree(tmp));
  return BFX_OK;
}

int build_transpose(Transpose& T, int32_t n_rows, const int32_t* dofmap0, int nd0, const int32_t* cells, int64_t ncells,
                    cudaStream_t st)
{
  int e;
  if ((e = dev_alloc(&T.ptr, (size_t)n_rows + 1)))
    return e;
  BFX_CUDA(cudaMemsetAsync(T.ptr, 0, sizeof(int64_t) * ((size_t)n_rows + 1), st));
  int* d_err = nullptr;
  BFX_CUDA(cudaMalloc(&d_err, sizeof(int)));
  BFX_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
  if (ncells > 0)
  {
    k_count_incidence<<<grid_for(ncells * nd0, 256, 0), 256, 0, st>>>(ncells, cells, dofmap0, nd0, n_rows, T.ptr, d_err);
    BFX_CHECK_LAUNCH();
  }
  int h_err = 0;
  BFX_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  BFX_CUDA(cudaStreamSynchronize(st));
  BFX_CUDA(cudaFree(d_err));
  if (h_err)
    return fail(BFX_ERR_INVALID, "sparsity: dofmap entry outside [0, %d)", n_rows);
  if ((e = exclusive_scan_i64(T.ptr, (int64_t)n_rows + 1, st)))
    return e;
  BFX_CUDA(cudaMemcpy(&T.total, T.ptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost));
  if ((e = dev_alloc(&T.cells, (size_t)T.total)))
    return e;
  if (T.total > 0)
  {
    int32_t* cursor = nullptr;
    if ((e = dev_alloc(&cursor, (size_t)n_rows)))
      return e;
    BFX_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)n_rows, st));
    k_fill_incidence<<<grid_for(ncells * nd0, 256, 0), 256, 0, st>>>(ncells, cells, dofmap0, nd0, n_rows, T.ptr, cursor,
                                                                     T.cells);
    k_sort_incidence<<<grid_for(n_rows, 128, 0), 128, 0, st>>>(n_rows, T.ptr, T.cells);
    BFX_CHECK_LAUNCH();
    BFX_CUDA(cudaStreamSynchronize(st));
    BFX_CUDA(cudaFree(cursor));
  }
  return BFX_OK;
}

template <int CAP>
int rows_pass(bool fill, int32_t n_rows, const Transpose& T, const int32_t* dofmap1, int nd1, const int64_t* ex_ptr,
              const int32_t* ex_cols, int32_t n_cols_owned, int64_t* counts, const int64_t* row_ptr, int32_t* cols,
              int64_t* off_diag, int* overflow, cudaStream_t st)
{
  const unsigned grid = grid_for(n_rows, 128, 0);
  if (fill)
    k_row_columns<CAP, true><<<grid, 128, 0, st>>>(n_rows, T.ptr, T.cells, dofmap1, nd1, ex_ptr, ex_cols, n_cols_owned,
                                                   counts, row_ptr, cols, off_diag, overflow);
  else
    k_row_columns<CAP, false><<<grid, 128, 0, st>>>(n_rows, T.ptr, T.cells, dofmap1, nd1, ex_ptr, ex_cols,
                                                    n_cols_owned, counts, row_ptr, cols, off_diag, overflow);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}
} // namespace

extern "C"
{
int bfx_sparsity_build(bfx_csr_t** out, int32_t n_rows_all, int32_t n_rows_owned, int32_t n_cols_owned,
                       const int32_t* dofmap0, int nd0, const int32_t* dofmap1, int nd1, const int32_t* cells,
                       int64_t ncells, const int32_t* extra_rows, const int32_t* extra_cols, int64_t n_extra, int bs0,
                       int bs1, bfx_stream_t stream)
{
  BFX_REQUIRE(out && dofmap0 && dofmap1 && nd0 > 0 && nd1 > 0 && n_rows_all >= 0 && n_rows_owned <= n_rows_all,
              "bfx_sparsity_build: bad arguments");
  cudaStream_t st = S(stream);
  Transpose T;
  int e = build_transpose(T, n_rows_all, dofmap0, nd0, cells, ncells, st);
  if (e)
  {
    T.release();
    return e;
  }
  // extra (row, col) entries: bucket by row on the host (few: ghost-row contributions of other ranks)
  int64_t* ex_ptr = nullptr;
  int32_t* ex_cols = nullptr;
  if (n_extra > 0)
  {
    std::vector<int32_t> hr((size_t)n_extra), hc((size_t)n_extra);
    BFX_CUDA(cudaMemcpy(hr.data(), extra_rows, sizeof(int32_t) * n_extra, cudaMemcpyDefault));
    BFX_CUDA(cudaMemcpy(hc.data(), extra_cols, sizeof(int32_t) * n_extra, cudaMemcpyDefault));
    std::vector<int64_t> hp((size_t)n_rows_all + 1, 0);
    for (int64_t k = 0; k < n_extra; ++k)
    {
      if (hr[k] < 0 || hr[k] >= n_rows_all)
      {
        T.release();
        return fail(BFX_ERR_INVALID, "bfx_sparsity_build: extra row %d out of range", hr[k]);
      }
      ++hp[hr[k] + 1];
    }
    for (int32_t r = 0; r < n_rows_all; ++r)
      hp[r + 1] += hp[r];
    std::vector<int32_t> hb((size_t)n_extra);
    std::vector<int64_t> pos(hp.begin(), hp.end() - 1);
    for (int64_t k = 0; k < n_extra; ++k)
      hb[pos[hr[k]]++] = hc[k];
    if ((e = upload(&ex_ptr, hp.data(), hp.size(), st)) || (e = upload(&ex_cols, hb.data(), hb.size(), st)))
      return e;
    BFX_CUDA(cudaStreamSynchronize(st));
  }

  bfx_csr* A = new bfx_csr();
  A->n_rows_all = n_rows_all;
  A->n_rows_owned = n_rows_owned;
  A->bs0 = bs0;
  A->bs1 = bs1;
  int* d_over = nullptr;
  BFX_CUDA(cudaMalloc(&d_over, sizeof(int)));
  if ((e = dev_alloc(&A->row_ptr, (size_t)n_rows_all + 1)) || (e = dev_alloc(&A->off_diag, (size_t)n_rows_all)))
    return e;
  int cap = 96;
  for (;;)
  {
    BFX_CUDA(cudaMemsetAsync(d_over, 0, sizeof(int), st));
    BFX_CUDA(cudaMemsetAsync(A->row_ptr, 0, sizeof(int64_t) * ((size_t)n_rows_all + 1), st));
    if (n_rows_all > 0)
    {
      e = cap == 96 ? rows_pass<96>(false, n_rows_all, T, dofmap1, nd1, ex_ptr, ex_cols, n_cols_owned, A->row_ptr,
                                    nullptr, nullptr, nullptr, d_over, st)
                    : rows_pass<768>(false, n_rows_all, T, dofmap1, nd1, ex_ptr, ex_cols, n_cols_owned, A->row_ptr,
                                     nullptr, nullptr, nullptr, d_over, st);
      if (e)
        return e;
    }
    int h_over = 0;
    BFX_CUDA(cudaMemcpyAsync(&h_over, d_over, sizeof(int), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaStreamSynchronize(st));
    if (!h_over)
      break;
    if (cap == 768)
    {
      T.release();
      return fail(BFX_ERR_UNSUPPORTED, "sparsity: a row has more than 768 distinct columns");
    }
    cap = 768;
  }
  if ((e = exclusive_scan_i64(A->row_ptr, (int64_t)n_rows_all + 1, st)))
    return e;
  BFX_CUDA(cudaMemcpy(&A->nnz, A->row_ptr + n_rows_all, sizeof(int64_t), cudaMemcpyDeviceToHost));
  if ((e = dev_alloc(&A->cols, (size_t)A->nnz)))
    return e;
  if (n_rows_all > 0)
  {
    e = cap == 96 ? rows_pass<96>(true, n_rows_all, T, dofmap1, nd1, ex_ptr, ex_cols, n_cols_owned, nullptr, A->row_ptr,
                                  A->cols, A->off_diag, d_over, st)
                  : rows_pass<768>(true, n_rows_all, T, dofmap1, nd1, ex_ptr, ex_cols, n_cols_owned, nullptr,
                                   A->row_ptr, A->cols, A->off_diag, d_over, st);
    if (e)
      return e;
  }
  BFX_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_over);
  cudaFree(ex_ptr);
  cudaFree(ex_cols);
  T.release();
  if ((e = csr_finish_create(A)))
    return e;
  *out = A;
  return BFX_OK;
}

int bfx_sparsity_ghost_rows(int32_t n_rows_all, int32_t n_rows_owned, const int32_t* dofmap0, int nd0,
                            const int32_t* dofmap1, int nd1, const int32_t* cells, int64_t ncells, int64_t* counts_host,
                            int32_t* cols_host, bfx_stream_t stream)
{
  BFX_REQUIRE(dofmap0 && dofmap1 && counts_host && n_rows_owned <= n_rows_all, "bfx_sparsity_ghost_rows: bad arguments");
  const int32_t n_ghost = n_rows_all - n_rows_owned;
  if (n_ghost == 0)
    return BFX_OK;
  cudaStream_t st = S(stream);
  Transpose T;
  int e = build_transpose(T, n_rows_all, dofmap0, nd0, cells, ncells, st);
  if (e)
  {
    T.release();
    return e;
  }
  std::vector<int64_t> hp((size_t)n_ghost + 1);
  BFX_CUDA(cudaMemcpy(hp.data(), T.ptr + n_rows_owned, sizeof(int64_t) * hp.size(), cudaMemcpyDeviceToHost));
  const int64_t cap = (hp[n_ghost] - hp[0]) * nd1;
  int64_t* d_counts = nullptr;
  int32_t* d_out = nullptr;
  if ((e = dev_alloc(&d_counts, (size_t)n_ghost)) || (e = dev_alloc(&d_out, (size_t)cap)))
    return e;
  k_ghost_row_columns<<<grid_for(n_ghost, 128, 0), 128, 0, st>>>(n_rows_owned, n_ghost, T.ptr, T.cells, dofmap1, nd1,
                                                                 d_counts, d_out);
  BFX_CHECK_LAUNCH();
  BFX_CUDA(cudaMemcpyAsync(counts_host, d_counts, sizeof(int64_t) * n_ghost, cudaMemcpyDeviceToHost, st));
  BFX_CUDA(cudaStreamSynchronize(st));
  if (cols_host)
  {
    std::vector<int32_t> h((size_t)cap);
    BFX_CUDA(cudaMemcpy(h.data(), d_out, sizeof(int32_t) * (size_t)cap, cudaMemcpyDeviceToHost));
    int64_t w = 0;
    for (int32_t g = 0; g < n_ghost; ++g)
    {
      const int64_t base = (hp[g] - hp[0]) * nd1;
      for (int64_t k = 0; k < counts_host[g]; ++k)
        cols_host[w++] = h[base + k];
    }
  }
  cudaFree(d_counts);
  cudaFree(d_out);
  T.release();
  return BFX_OK;
}
}
