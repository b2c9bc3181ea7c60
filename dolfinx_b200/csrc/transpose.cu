// la::transpose on the device: the local part of la/mattrans.h.
//
// impl::local_transpose (la/mattrans.h:47-108) transposes the block "owned rows x owned columns" of a MatrixCSR
// (entries row_ptr[i] .. off_diag_offset[i] of every owned row i): row j of the result lists the rows i of A with an
// entry in column j in ASCENDING i (the reference walks the rows in order with one write cursor per column), every
// bs0 x bs1 block is stored transposed (bs1 x bs0, row-major).
//
// Here: the diagonal-block entries get the key (column, entry index) - one 64-bit key, since the entry index grows
// with the row - and are sorted by one device radix sort (cub); the other entries sort to the end.  The row pointer
// of the result is a lower bound per column over the sorted keys, the row of an entry a binary search in row_ptr.
// Same order, same values as the reference's sequential loop: bit-exact.
#include "csr.cuh"
#include <cub/device/device_radix_sort.cuh>

using namespace bfx;

namespace
{
__global__ void k_transpose_keys(int32_t n_rows, const int64_t* __restrict__ row_ptr, const int64_t* __restrict__ off_diag,
                                 const int32_t* __restrict__ cols, int entry_bits, uint64_t* __restrict__ keys)
{
  // one warp per row: coalesced over the row's entries
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n_rows; i += nwarps)
  {
    const int64_t b = row_ptr[i], d = off_diag[i], e = row_ptr[i + 1];
    for (int64_t k = b + lane; k < e; k += 32)
      keys[k] = k < d ? ((uint64_t)(uint32_t)cols[k] << entry_bits) | (uint64_t)k : ~0ull;
  }
}

// row_ptrT[j] = number of sorted keys whose column is < j
__global__ void k_transpose_row_ptr(int32_t n_cols, int64_t n_keys, int entry_bits, const uint64_t* __restrict__ keys,
                                    int64_t* __restrict__ row_ptrT)
{
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= n_cols; j += (int64_t)gridDim.x * blockDim.x)
  {
    const uint64_t bound = (uint64_t)j << entry_bits; // first possible key of column j
    int64_t lo = 0, hi = n_keys;
    while (lo < hi)
    {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < bound)
        lo = mid + 1;
      else
        hi = mid;
    }
    row_ptrT[j] = lo;
  }
}

__global__ void k_transpose_fill(int64_t nnzT, int32_t n_rows, int entry_bits, const uint64_t* __restrict__ keys,
                                 const int64_t* __restrict__ row_ptr, const double* __restrict__ values, int bs0, int bs1,
                                 int32_t* __restrict__ colsT, double* __restrict__ valsT)
{
  const int nbs = bs0 * bs1;
  const uint64_t mask = (1ull << entry_bits) - 1ull;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnzT; p += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t k = (int64_t)(keys[p] & mask);
    // row of entry k: last i with row_ptr[i] <= k
    int32_t lo = 0, hi = n_rows;
    while (hi - lo > 1)
    {
      const int32_t mid = lo + ((hi - lo) >> 1);
      if (row_ptr[mid] <= k)
        lo = mid;
      else
        hi = mid;
    }
    colsT[p] = lo;
    // A block (bs0 x bs1) row-major -> AT block (bs1 x bs0) row-major (la/mattrans.h:86-99)
    for (int k0 = 0; k0 < bs0; ++k0)
      for (int k1 = 0; k1 < bs1; ++k1)
        valsT[p * nbs + k1 * bs0 + k0] = values[k * nbs + k0 * bs1 + k1];
  }
}
} // namespace

extern "C"
{
int bfx_csr_transpose_local(const bfx_csr_t* A, const double* values, int32_t n_cols_owned, int64_t* row_ptrT,
                            int32_t* colsT, double* valsT, int64_t capacity, int64_t* nnzT_host, bfx_stream_t stream)
{
  BFX_REQUIRE(A && values && row_ptrT && nnzT_host && n_cols_owned >= 0, "bfx_csr_transpose_local: null argument");
  cudaStream_t st = S(stream);
  const int64_t n_keys = A->nnz_owned; // entries of the owned rows
  int entry_bits = 1;
  while (entry_bits < 40 && (1ll << entry_bits) <= n_keys)
    ++entry_bits;
  int col_bits = 1;
  while (col_bits < 32 && (1ll << col_bits) <= (int64_t)n_cols_owned)
    ++col_bits;
  if (entry_bits + col_bits > 63)
    return fail(BFX_ERR_UNSUPPORTED, "bfx_csr_transpose_local: %lld entries x %d columns exceed the 63-bit sort key",
                (long long)n_keys, (int)n_cols_owned);
  *nnzT_host = 0;
  if (n_keys == 0)
  {
    BFX_CUDA(cudaMemsetAsync(row_ptrT, 0, sizeof(int64_t) * ((size_t)n_cols_owned + 1), st));
    BFX_CUDA(cudaStreamSynchronize(st));
    return BFX_OK;
  }
  uint64_t *k0 = nullptr, *k1 = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  int e;
  if ((e = dev_alloc(&k0, (size_t)n_keys)) || (e = dev_alloc(&k1, (size_t)n_keys)))
  {
    cudaFree(k0);
    return e;
  }
  auto cleanup = [&]()
  {
    cudaFree(k0);
    cudaFree(k1);
    cudaFree(tmp);
  };
  k_transpose_keys<<<grid_for((int64_t)A->n_rows_owned * 32, 256, 16), 256, 0, st>>>(A->n_rows_owned, A->row_ptr, A->off_diag,
                                                                                    A->cols, entry_bits, k0);
  // (ghost-column entries carry ~0: every bit above the sorted range is set too, so they end up last)
  const int end_bit = 64;
  cudaError_t ce = cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, k0, k1, n_keys, 0, end_bit, st);
  if (ce == cudaSuccess)
    ce = cudaMalloc(&tmp, tmp_bytes);
  if (ce == cudaSuccess)
    ce = cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, k0, k1, n_keys, 0, end_bit, st);
  if (ce != cudaSuccess)
  {
    cleanup();
    return fail(BFX_ERR_CUDA, "bfx_csr_transpose_local: %s", cudaGetErrorString(ce));
  }
  k_transpose_row_ptr<<<grid_for((int64_t)n_cols_owned + 1, 256, 16), 256, 0, st>>>(n_cols_owned, n_keys, entry_bits, k1,
                                                                                   row_ptrT);
  int64_t nnzT = 0;
  ce = cudaMemcpyAsync(&nnzT, row_ptrT + n_cols_owned, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess)
    ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess)
  {
    cleanup();
    return fail(BFX_ERR_CUDA, "bfx_csr_transpose_local: %s", cudaGetErrorString(ce));
  }
  *nnzT_host = nnzT;
  if (colsT && valsT)
  {
    if (capacity < nnzT)
    {
      cleanup();
      return fail(BFX_ERR_INVALID, "bfx_csr_transpose_local: capacity %lld < %lld entries", (long long)capacity,
                  (long long)nnzT);
    }
    if (nnzT > 0)
      k_transpose_fill<<<grid_for(nnzT, 256, 16), 256, 0, st>>>(nnzT, A->n_rows_owned, entry_bits, k1, A->row_ptr, values,
                                                               A->bs0, A->bs1, colsT, valsT);
    ce = cudaGetLastError();
    if (ce == cudaSuccess)
      ce = cudaStreamSynchronize(st);
  }
  cleanup();
  if (ce != cudaSuccess)
    return fail(BFX_ERR_CUDA, "bfx_csr_transpose_local: %s", cudaGetErrorString(ce));
  return BFX_OK;
}
}
