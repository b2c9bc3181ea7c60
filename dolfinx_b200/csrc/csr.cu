// la::MatrixCSR on device: structure, SpMV / SpMV^T, insertion, set_diagonal, reductions.
#include "csr.cuh"
#include <algorithm>
#include <vector>

using namespace bfx;

// ---------------------------------------------------------------------------------------------
// reductions (la/Vector.h:434-514, la/MatrixCSR.h:473-486)
// ---------------------------------------------------------------------------------------------
namespace
{
constexpr int RED_THREADS = 256;

struct RedWorkspace
{
  double* partial = nullptr; // [max_blocks + 1]
  int max_blocks = 0;
};

RedWorkspace& red_ws()
{
  static thread_local RedWorkspace ws;
  return ws;
}

int ensure_red_ws()
{
  RedWorkspace& ws = red_ws();
  if (!ws.partial)
  {
    ws.max_blocks = sm_count() * 8;
    BFX_CUDA(cudaMalloc(&ws.partial, sizeof(double) * (ws.max_blocks + 1)));
  }
  return BFX_OK;
}

// op: 0 = sum x*y, 1 = sum |x|, 2 = max |x|
template <int OP>
__global__ void __launch_bounds__(RED_THREADS) k_reduce_stage1(int64_t n, const double* __restrict__ x,
                                                               const double* __restrict__ y, double* __restrict__ out)
{
  double acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
  {
    if (OP == 0)
      acc = fma(x[i], y[i], acc);
    else if (OP == 1)
      acc += fabs(x[i]);
    else
      acc = fmax(acc, fabs(x[i]));
  }
  __shared__ double sm[RED_THREADS / 32];
  if (OP == 2)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      acc = fmax(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  }
  else
    acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0)
    sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32)
  {
    double v = threadIdx.x < RED_THREADS / 32 ? sm[threadIdx.x] : 0.0;
    if (OP == 2)
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    else
      v = warp_sum(v);
    if (threadIdx.x == 0)
      out[blockIdx.x] = v;
  }
}

template <int OP>
__global__ void __launch_bounds__(RED_THREADS) k_reduce_stage2(int nb, const double* __restrict__ partial,
                                                               double* __restrict__ out)
{
  // deterministic fixed-order final sum
  double acc = 0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x)
    acc = OP == 2 ? fmax(acc, partial[i]) : acc + partial[i];
  __shared__ double sm[RED_THREADS];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = RED_THREADS / 2; s > 0; s >>= 1)
  {
    if (threadIdx.x < s)
      sm[threadIdx.x] = OP == 2 ? fmax(sm[threadIdx.x], sm[threadIdx.x + s]) : sm[threadIdx.x] + sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *out = sm[0];
}

template <int OP>
int reduce(int64_t n, const double* x, const double* y, double* result_host, cudaStream_t st)
{
  int e = ensure_red_ws();
  if (e)
    return e;
  RedWorkspace& ws = red_ws();
  int nb = (int)std::min<int64_t>(ws.max_blocks, std::max<int64_t>(1, (n + RED_THREADS - 1) / RED_THREADS));
  k_reduce_stage1<OP><<<nb, RED_THREADS, 0, st>>>(n, x, y, ws.partial);
  k_reduce_stage2<OP><<<1, RED_THREADS, 0, st>>>(nb, ws.partial, ws.partial + ws.max_blocks);
  BFX_CHECK_LAUNCH();
  BFX_CUDA(cudaMemcpyAsync(result_host, ws.partial + ws.max_blocks, sizeof(double), cudaMemcpyDeviceToHost, st));
  BFX_CUDA(cudaStreamSynchronize(st));
  return BFX_OK;
}

__global__ void k_fill(int64_t n, double v, double* __restrict__ x)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = v;
}

__global__ void k_axpy(int64_t n, double alpha, const double* __restrict__ x, double* __restrict__ y)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fma(alpha, x[i], y[i]);
}
} // namespace

// ---------------------------------------------------------------------------------------------
// SpMV (la/matrix_csr_impl.h:259-286; split of la/MatrixCSR.h:877-946)
// ---------------------------------------------------------------------------------------------
namespace
{
constexpr int SPMV_THREADS = 256;
constexpr int SPMV_CHUNK = 4096; // products staged per pass (32 KB of shared memory)

// bs = 1.  One CTA owns R = SPMV_THREADS / LPR consecutive rows; their nonzeros are one contiguous
// range of values/cols.  Consecutive lanes take consecutive entries (coalesced 4 B / 8 B streaming
// loads), so one warp instruction gathers x for ~2 adjacent rows, whose column sets share cache
// lines: the x gather — the L1-wavefront limiter of this kernel (ncu, profiles/r01_prof_spmv_p1_128.csv)
// — touches ~2x fewer lines than with a 4-entries-per-lane (128-bit) mapping.  Products are staged
// in shared memory and LPR lanes reduce each row.
// part: FULL sums [row_ptr[i], row_ptr[i+1]); DIAG sums [row_ptr[i], off_diag[i]) (products of the
// ghost columns are staged but never read).
template <int LPR>
__global__ void __launch_bounds__(SPMV_THREADS)
    k_spmv_stream(int32_t n_rows, const int64_t* __restrict__ row_ptr, const int64_t* __restrict__ row_end_sel,
                  const int32_t* __restrict__ cols, const double* __restrict__ values, const double* __restrict__ x,
                  double* __restrict__ y)
{
  constexpr int R = SPMV_THREADS / LPR;
  constexpr int UNROLL = 8;
  __shared__ double prod[SPMV_CHUNK];
  const int32_t r0 = blockIdx.x * R;
  const int32_t r1 = min(r0 + R, n_rows);
  const int64_t start = row_ptr[r0];
  const int64_t end = row_ptr[r1];

  const int my_row = r0 + threadIdx.x / LPR;
  const int sub = threadIdx.x % LPR;
  int64_t rb = 0, re = 0;
  if (my_row < r1)
  {
    rb = row_ptr[my_row];
    re = row_end_sel[my_row];
  }
  double sum = 0.0;

  for (int64_t base = start; base < end; base += SPMV_CHUNK)
  {
    // ---- stage products: SPMV_CHUNK / SPMV_THREADS entries per thread, UNROLL loads in flight
#pragma unroll
    for (int it0 = 0; it0 < SPMV_CHUNK / SPMV_THREADS; it0 += UNROLL)
    {
      int32_t c[UNROLL];
      double v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
      {
        const int64_t k = base + (it0 + u) * SPMV_THREADS + threadIdx.x;
        c[u] = k < end ? __ldg(cols + k) : -1;
        v[u] = k < end ? __ldg(values + k) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (c[u] >= 0)
          prod[(it0 + u) * SPMV_THREADS + threadIdx.x] = v[u] * __ldg(x + c[u]);
    }
    __syncthreads();
    // ---- reduce the part of my row that lies in this chunk
    const int64_t lo = max(rb, base), hi = min(re, base + SPMV_CHUNK);
    for (int64_t k = lo + sub; k < hi; k += LPR)
      sum += prod[k - base];
    __syncthreads();
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1)
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (my_row < r1 && sub == 0)
    y[my_row] += sum;
}

// bs = 1, row-per-thread variant.  The CTA's contiguous (cols, values) range is staged raw in shared
// memory with coalesced 128-bit loads; thread t then walks row r0+t out of shared memory (odd row
// strides are bank-conflict free) and gathers x itself, accumulating in a register.  One warp
// instruction gathers the k-th entries of 32 CONSECUTIVE ROWS: on locality-ordered matrices these
// are neighbouring x entries (2-3 cache lines), which is what the entry-consecutive mapping of
// k_spmv_stream cannot achieve.  Which of the two kernels is faster depends on the numbering, so
// bfx_spmv times both once per matrix (plan-time autotuning) and keeps the winner.
constexpr int SPMV_ROWS_CAP = 3840; // staged entries per pass: 45 KB of shared memory

template <int R>
__global__ void __launch_bounds__(R)
    k_spmv_rows(int32_t n_rows, const int64_t* __restrict__ row_ptr, const int64_t* __restrict__ row_end_sel,
                const int32_t* __restrict__ cols, const double* __restrict__ values, const double* __restrict__ x,
                double* __restrict__ y)
{
  __shared__ __align__(16) double s_val[SPMV_ROWS_CAP];
  __shared__ __align__(16) int32_t s_col[SPMV_ROWS_CAP];
  const int32_t r0 = blockIdx.x * R;
  const int32_t r1 = min(r0 + R, n_rows);
  const int64_t start = row_ptr[r0];
  const int64_t end = row_ptr[r1];
  const int my_row = r0 + threadIdx.x;
  int64_t rb = 0, re = 0;
  if (my_row < r1)
  {
    rb = row_ptr[my_row];
    re = row_end_sel[my_row];
  }
  double sum = 0.0;
  const int64_t astart = start & ~int64_t(3);
  for (int64_t base = astart; base < end; base += SPMV_ROWS_CAP)
  {
    const int64_t lim = min(end, base + SPMV_ROWS_CAP);
    // ---- coalesced 128-bit staging of cols and values (4 entries per thread per step)
    for (int64_t k = base + 4 * (int64_t)threadIdx.x; k < lim; k += 4 * R)
    {
      const int o = (int)(k - base);
      if (k + 3 < lim)
      {
        *reinterpret_cast<int4*>(s_col + o) = ldg_stream(reinterpret_cast<const int4*>(cols + k));
        *reinterpret_cast<double2*>(s_val + o) = ldg_stream(reinterpret_cast<const double2*>(values + k));
        *reinterpret_cast<double2*>(s_val + o + 2) = ldg_stream(reinterpret_cast<const double2*>(values + k + 2));
      }
      else
      {
        for (int m = 0; m < 4 && k + m < lim; ++m)
        {
          s_col[o + m] = cols[k + m];
          s_val[o + m] = values[k + m];
        }
      }
    }
    __syncthreads();
    const int64_t lo = max(rb, base), hi = min(re, lim);
#pragma unroll 4
    for (int64_t k = lo; k < hi; ++k)
      sum = fma(s_val[k - base], __ldg(x + s_col[k - base]), sum);
    __syncthreads();
  }
  if (my_row < r1)
    y[my_row] += sum;
}

// bs = 1, TMA-pipelined row-per-thread variant.  Same arithmetic as k_spmv_rows, but the contiguous
// (cols, values) range of a tile of R rows reaches shared memory through TMA bulk copies
// (cp.async.bulk + mbarrier, SASS UBLKCP) issued by one thread, two stages deep: while the CTA walks
// the rows of one tile the next tile is already in flight, no registers or LSU slots are spent on the
// streaming part, and persistent CTAs (grid = resident CTAs) keep the pipeline full across tiles.
constexpr int SPMV_TMA_CAP = 3840; // entries per stage (multiple of 4: 16-byte granules of cols)

__device__ __forceinline__ uint32_t sp_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sp_mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sp_mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sp_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sp_bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   sp_smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(sp_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sp_mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra DONE_%=;\n"
               "bra WAIT_%=;\n"
               "DONE_%=:\n"
               "}" ::"r"(sp_smem_u32(bar)),
               "r"(parity)
               : "memory");
}

// position in the CTA's sequence of chunks: tile (R rows) and the entry range of the chunk inside it
struct SpmvChunkIt
{
  int64_t tile, base, end;
};

template <int R, int LPR>
__global__ void __launch_bounds__(R * LPR)
    k_spmv_tma(int32_t n_rows, int64_t nnz_total, const int64_t* __restrict__ row_ptr,
               const int64_t* __restrict__ row_end_sel, const int32_t* __restrict__ cols,
               const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y)
{
  extern __shared__ __align__(128) unsigned char sp_raw[];
  double* s_val = reinterpret_cast<double*>(sp_raw);                                          // [2][CAP]
  int32_t* s_col = reinterpret_cast<int32_t*>(sp_raw + 2 * SPMV_TMA_CAP * sizeof(double));     // [2][CAP]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sp_raw + 2 * SPMV_TMA_CAP * (sizeof(double) + sizeof(int32_t)));
  const int64_t ntiles = ((int64_t)n_rows + R - 1) / R;
  if (threadIdx.x == 0)
  {
    sp_mbar_init(&bar[0], 1);
    sp_mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto load_tile = [&](SpmvChunkIt& it)
  {
    if (it.tile < ntiles)
    {
      const int64_t start = row_ptr[it.tile * R];
      it.end = row_ptr[min((it.tile + 1) * R, (int64_t)n_rows)];
      it.base = start & ~int64_t(3);
    }
  };
  auto advance = [&](SpmvChunkIt& it)
  {
    it.base += SPMV_TMA_CAP;
    if (it.base >= it.end)
    {
      it.tile += gridDim.x;
      load_tile(it);
    }
  };
  // bytes the TMA brings for a chunk: whole 16-byte granules that lie inside the arrays
  auto tma_entries = [&](const SpmvChunkIt& it) -> int64_t
  {
    const int64_t lim = min(it.end, it.base + SPMV_TMA_CAP);
    const int64_t want = (lim - it.base + 3) & ~int64_t(3);
    const int64_t avail = (nnz_total - it.base) & ~int64_t(3);
    return max((int64_t)0, min(want, avail));
  };
  auto issue = [&](const SpmvChunkIt& it, int stage)
  {
    const uint32_t n = (uint32_t)tma_entries(it);
    sp_mbar_expect_tx(&bar[stage], n * 12u);
    if (n)
    {
      sp_bulk_g2s(s_col + stage * SPMV_TMA_CAP, cols + it.base, n * 4u, &bar[stage]);
      sp_bulk_g2s(s_val + stage * SPMV_TMA_CAP, values + it.base, n * 8u, &bar[stage]);
    }
  };

  SpmvChunkIt cons, prod;
  cons.tile = blockIdx.x, cons.base = 0, cons.end = 0;
  load_tile(cons);
  prod = cons;
  // prologue: two chunks in flight
  if (threadIdx.x == 0)
  {
    if (prod.tile < ntiles)
      issue(prod, 0);
  }
  if (prod.tile < ntiles)
    advance(prod);
  if (threadIdx.x == 0)
  {
    if (prod.tile < ntiles)
      issue(prod, 1);
  }
  if (prod.tile < ntiles)
    advance(prod);

  // row bounds and y of the CURRENT tile are fetched one chunk ahead (their latency would otherwise
  // sit between the barrier and the first product)
  double sum = 0.0, y_old = 0.0;
  int64_t rb = 0, re = 0;
  {
    const int64_t my_row = cons.tile * R + threadIdx.x / LPR;
    if (cons.tile < ntiles && my_row < n_rows)
    {
      rb = row_ptr[my_row];
      re = row_end_sel[my_row];
      y_old = y[my_row];
    }
  }
  for (uint32_t c = 0; cons.tile < ntiles; ++c)
  {
    const int stage = c & 1;
    const int64_t my_row = cons.tile * R + threadIdx.x / LPR;
    const int64_t lim = min(cons.end, cons.base + SPMV_TMA_CAP);
    const bool last = cons.base + SPMV_TMA_CAP >= cons.end;
    // prefetch for the next tile of this CTA
    SpmvChunkIt nxt = cons;
    advance(nxt);
    int64_t rb_n = 0, re_n = 0;
    double y_n = 0.0;
    if (last && nxt.tile < ntiles)
    {
      const int64_t row_n = nxt.tile * R + threadIdx.x / LPR;
      if (row_n < n_rows)
      {
        rb_n = row_ptr[row_n];
        re_n = row_end_sel[row_n];
        y_n = y[row_n];
      }
    }
    sp_mbar_wait(&bar[stage], (c >> 1) & 1);
    double* sv = s_val + stage * SPMV_TMA_CAP;
    int32_t* sc = s_col + stage * SPMV_TMA_CAP;
    const int64_t got = cons.base + tma_entries(cons);
    if (got < lim)
    {
      // the last < 4 entries of the arrays: not a whole 16-byte granule, loaded by hand
      for (int64_t k = got + threadIdx.x; k < lim; k += R * LPR)
      {
        sv[k - cons.base] = values[k];
        sc[k - cons.base] = cols[k];
      }
      __syncthreads();
    }
    const int lo = (int)(max(rb, cons.base) - cons.base), hi = (int)(min(re, lim) - cons.base);
    // 8 gathers of x in flight per thread, two accumulators; LPR lanes share a row (entries dealt round-robin)
    double s0 = 0.0, s1 = 0.0;
    for (int k = lo + (int)(threadIdx.x % LPR); k < hi; k += 8 * LPR)
    {
      double xv[8], av[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
      {
        const bool in = k + u * LPR < hi;
        av[u] = in ? sv[k + u * LPR] : 0.0;
        xv[u] = in ? __ldg(x + sc[k + u * LPR]) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u += 2)
      {
        s0 = fma(av[u], xv[u], s0);
        s1 = fma(av[u + 1], xv[u + 1], s1);
      }
    }
    sum += s0 + s1;
    if (last)
    {
      double tot = sum;
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1)
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
      if (my_row < n_rows && threadIdx.x % LPR == 0)
        y[my_row] = y_old + tot;
    }
    __syncthreads(); // every thread is done with this stage: it can be refilled
    if (threadIdx.x == 0 && prod.tile < ntiles)
      issue(prod, stage);
    if (prod.tile < ntiles)
      advance(prod);
    if (last)
    {
      rb = rb_n, re = re_n, y_old = y_n;
      sum = 0.0;
    }
    cons = nxt;
  }
}

// Blocked rows, compile-time block sizes: one warp per block row.  The row's nnz*BS0*BS1 contiguous
// scalars are copied to a per-warp shared buffer with coalesced loads (32 blocks per pass), then
// each lane owns one block: BS0*BS1 conflict-free LDS (odd stride), BS1 gathered x values, BS0
// accumulators, shuffle reduction at the end of the row.
template <int BS0, int BS1>
__global__ void __launch_bounds__(256)
    k_spmv_blocked(int32_t n_rows, const int64_t* __restrict__ row_begin, const int64_t* __restrict__ row_end,
                   const int32_t* __restrict__ cols, const double* __restrict__ values, const double* __restrict__ x,
                   double* __restrict__ y, const int32_t* __restrict__ row_list)
{
  constexpr int BS2 = BS0 * BS1;
  constexpr int STRIDE = (BS2 % 2 == 0) ? BS2 + 1 : BS2; // odd stride -> conflict-free per-lane blocks
  __shared__ double stage[8][32 * STRIDE];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* buf = stage[wib];
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n_rows; w += nw)
  {
    const int32_t row = row_list ? row_list[w] : (int32_t)w;
    const int64_t b = row_begin[row], e = row_end[row];
    double acc[BS0];
#pragma unroll
    for (int k = 0; k < BS0; ++k)
      acc[k] = 0.0;
    for (int64_t j0 = b; j0 < e; j0 += 32)
    {
      const int nb = (int)min((int64_t)32, e - j0);
      const double* src = values + j0 * BS2;
      // coalesced copy of nb*BS2 scalars; scalar s of block q goes to buf[q*STRIDE + s]
#pragma unroll
      for (int it = 0; it < BS2; ++it)
      {
        const int t = it * 32 + lane;
        if (t < nb * BS2)
        {
          const int q = t / BS2;
          buf[q * STRIDE + (t - q * BS2)] = __ldg(src + t);
        }
      }
      __syncwarp();
      if (lane < nb)
      {
        const int64_t c = cols[j0 + lane];
        double xv[BS1];
#pragma unroll
        for (int k1 = 0; k1 < BS1; ++k1)
          xv[k1] = __ldg(x + c * BS1 + k1);
#pragma unroll
        for (int k0 = 0; k0 < BS0; ++k0)
#pragma unroll
          for (int k1 = 0; k1 < BS1; ++k1)
            acc[k0] = fma(buf[lane * STRIDE + k0 * BS1 + k1], xv[k1], acc[k0]);
      }
      __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < BS0; ++k)
      acc[k] = warp_sum(acc[k]);
    if (lane == 0)
    {
#pragma unroll
      for (int k = 0; k < BS0; ++k)
        y[(int64_t)row * BS0 + k] += acc[k];
    }
  }
}

// Blocked rows fed by TMA: persistent CTAs stream the contiguous (cols, values) range of a tile of TR block
// rows into shared memory with cp.async.bulk (two stages, like k_spmv_tma); the blocks land in their
// storage layout (BS2 consecutive scalars, odd stride for 3x3: lane-per-block reads are conflict free),
// so the per-lane staging loads/stores of k_spmv_blocked disappear from the LSU pipe.  One warp per block
// row of the tile, one lane per block, shuffle reduction per row.
constexpr int SPMV_BTMA_CAP = 480; // blocks per stage (multiple of 4)
constexpr int SPMV_BTMA_TR = 16;   // block rows per tile

template <int BS0, int BS1, int NW>
__global__ void __launch_bounds__(32 * NW)
    k_spmv_blocked_tma(int32_t n_rows, int64_t nnz_total, const int64_t* __restrict__ row_ptr,
                       const int64_t* __restrict__ row_begin, const int64_t* __restrict__ row_end,
                       const int32_t* __restrict__ cols, const double* __restrict__ values,
                       const double* __restrict__ x, double* __restrict__ y)
{
  constexpr int BS2 = BS0 * BS1, TR = SPMV_BTMA_TR, CAP = SPMV_BTMA_CAP;
  static_assert((BS2 * 8 * 2) % 16 == 0, "two blocks are a whole number of 16-byte granules");
  extern __shared__ __align__(128) unsigned char sp_raw[];
  double* s_val = reinterpret_cast<double*>(sp_raw);                                      // [2][CAP * BS2]
  int32_t* s_col = reinterpret_cast<int32_t*>(sp_raw + 2 * (size_t)CAP * BS2 * sizeof(double)); // [2][CAP]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sp_raw + 2 * (size_t)CAP * (BS2 * sizeof(double) + sizeof(int32_t)));
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ntiles = ((int64_t)n_rows + TR - 1) / TR;
  if (threadIdx.x == 0)
  {
    sp_mbar_init(&bar[0], 1);
    sp_mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto load_tile = [&](SpmvChunkIt& it)
  {
    if (it.tile < ntiles)
    {
      const int64_t start = row_ptr[it.tile * TR];
      it.end = row_ptr[min((it.tile + 1) * TR, (int64_t)n_rows)];
      it.base = start & ~int64_t(3);
    }
  };
  auto advance = [&](SpmvChunkIt& it)
  {
    it.base += CAP;
    if (it.base >= it.end)
    {
      it.tile += gridDim.x;
      load_tile(it);
    }
  };
  auto tma_blocks = [&](const SpmvChunkIt& it) -> int64_t
  {
    const int64_t lim = min(it.end, it.base + CAP);
    const int64_t want = (lim - it.base + 3) & ~int64_t(3);
    const int64_t avail = (nnz_total - it.base) & ~int64_t(3);
    return max((int64_t)0, min(want, avail));
  };
  auto issue = [&](const SpmvChunkIt& it, int stage)
  {
    const uint32_t n = (uint32_t)tma_blocks(it);
    sp_mbar_expect_tx(&bar[stage], n * (uint32_t)(4 + BS2 * 8));
    if (n)
    {
      sp_bulk_g2s(s_col + stage * CAP, cols + it.base, n * 4u, &bar[stage]);
      sp_bulk_g2s(s_val + (size_t)stage * CAP * BS2, values + it.base * BS2, n * (uint32_t)(BS2 * 8), &bar[stage]);
    }
  };

  // three iterators: it0 is consumed, it1 is in flight, it2 is the next to issue.  The row_ptr loads of a tile
  // happen once (in advance(it2)), two iterations before the tile is consumed.
  SpmvChunkIt it0, it1, it2;
  it0.tile = blockIdx.x, it0.base = 0, it0.end = 0;
  load_tile(it0);
  if (threadIdx.x == 0 && it0.tile < ntiles)
    issue(it0, 0);
  it1 = it0;
  if (it1.tile < ntiles)
    advance(it1);
  if (threadIdx.x == 0 && it1.tile < ntiles)
    issue(it1, 1);
  it2 = it1;
  if (it2.tile < ntiles)
    advance(it2);

  // each warp owns rows wib, wib + NW, ... of the tile; partial sums live across the chunks of a tile
  double acc[TR / NW][BS0];
#pragma unroll
  for (int s = 0; s < TR / NW; ++s)
#pragma unroll
    for (int k = 0; k < BS0; ++k)
      acc[s][k] = 0.0;
  for (uint32_t c = 0; it0.tile < ntiles; ++c)
  {
    const int stage = c & 1;
    const int64_t lim = min(it0.end, it0.base + CAP);
    const bool last = it0.base + CAP >= it0.end;
    // row bounds and the old y of my rows: issued before the wait, consumed after it
    int64_t rbv[TR / NW], rev[TR / NW];
    double yold[TR / NW][BS0];
#pragma unroll
    for (int s = 0; s < TR / NW; ++s)
    {
      const int64_t row = it0.tile * TR + wib + NW * s;
      rbv[s] = rev[s] = 0;
      if (row < n_rows)
      {
        rbv[s] = row_begin[row];
        rev[s] = row_end[row];
        if (last && lane < BS0)
          yold[s][0] = y[row * BS0 + lane];
      }
    }
    sp_mbar_wait(&bar[stage], (c >> 1) & 1);
    double* sv = s_val + (size_t)stage * CAP * BS2;
    int32_t* sc = s_col + stage * CAP;
    const int64_t got = it0.base + tma_blocks(it0);
    if (got < lim)
    {
      for (int64_t k = got * BS2 + threadIdx.x; k < lim * BS2; k += blockDim.x)
        sv[k - it0.base * BS2] = values[k];
      for (int64_t k = got + threadIdx.x; k < lim; k += blockDim.x)
        sc[k - it0.base] = cols[k];
      __syncthreads();
    }
#pragma unroll
    for (int s = 0; s < TR / NW; ++s)
    {
      const int64_t lo = max(rbv[s], it0.base), hi = min(rev[s], lim);
      for (int64_t j0 = lo; j0 < hi; j0 += 32)
      {
        const int64_t j = j0 + lane;
        if (j < hi)
        {
          const int o = (int)(j - it0.base);
          const int64_t col = sc[o];
          double xv[BS1];
#pragma unroll
          for (int k1 = 0; k1 < BS1; ++k1)
            xv[k1] = __ldg(x + col * BS1 + k1);
#pragma unroll
          for (int k0 = 0; k0 < BS0; ++k0)
#pragma unroll
            for (int k1 = 0; k1 < BS1; ++k1)
              acc[s][k0] = fma(sv[o * BS2 + k0 * BS1 + k1], xv[k1], acc[s][k0]);
        }
      }
    }
    if (last)
    {
#pragma unroll
      for (int s = 0; s < TR / NW; ++s)
      {
        const int64_t row = it0.tile * TR + wib + NW * s;
        double mine = 0.0;
#pragma unroll
        for (int k = 0; k < BS0; ++k)
        {
          const double v = warp_sum(acc[s][k]);
          mine = lane == k ? v : mine;
          acc[s][k] = 0.0;
        }
        if (row < n_rows && lane < BS0)
          y[row * BS0 + lane] = yold[s][0] + mine; // lanes 0..BS0-1 write the BS0 components
      }
    }
    __syncthreads(); // every warp is done with this stage: it can be refilled
    if (threadIdx.x == 0 && it2.tile < ntiles)
      issue(it2, stage);
    it0 = it1;
    it1 = it2;
    if (it2.tile < ntiles)
      advance(it2);
  }
}

// Runtime block sizes (fallback; k0 outermost like the reference loop)
__global__ void __launch_bounds__(256)
    k_spmv_generic(int32_t n_rows, const int64_t* __restrict__ row_begin, const int64_t* __restrict__ row_end,
                   const int32_t* __restrict__ cols, const double* __restrict__ values, const double* __restrict__ x,
                   double* __restrict__ y, int bs0, int bs1, const int32_t* __restrict__ row_list)
{
  const int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int bs2 = bs0 * bs1;
  for (; w < n_rows; w += nw)
  {
    const int32_t row = row_list ? row_list[w] : (int32_t)w;
    const int64_t b = row_begin[row], e = row_end[row];
    for (int k0 = 0; k0 < bs0; ++k0)
    {
      double acc = 0.0;
      const int64_t n = (e - b) * bs1;
      for (int64_t t = lane; t < n; t += 32)
      {
        const int64_t j = b + t / bs1;
        const int k1 = (int)(t % bs1);
        acc += values[j * bs2 + k0 * bs1 + k1] * x[(int64_t)cols[j] * bs1 + k1];
      }
      acc = warp_sum(acc);
      if (lane == 0)
        y[(int64_t)row * bs0 + k0] += acc;
    }
  }
}

// y[cols[j]*bs1+k1] += values[...] * x[i*bs0+k0]  (la/matrix_csr_impl.h:319-343), scalar-parallel with REDs
__global__ void __launch_bounds__(256)
    k_spmvT(int32_t n_rows, const int64_t* __restrict__ row_begin, const int64_t* __restrict__ row_end,
            const int32_t* __restrict__ cols, const double* __restrict__ values, const double* __restrict__ x,
            double* __restrict__ y, int bs0, int bs1)
{
  const int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int bs2 = bs0 * bs1;
  for (; w < n_rows; w += nw)
  {
    const int64_t b = row_begin[w], e = row_end[w];
    for (int64_t s = b * bs2 + lane; s < e * bs2; s += 32)
    {
      const int64_t j = s / bs2;
      const int rem = (int)(s - j * bs2);
      const int k0 = rem / bs1, k1 = rem - k0 * bs1;
      red_add(y + (int64_t)cols[j] * bs1 + k1, values[s] * x[w * bs0 + k0]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// insertion (la/matrix_csr_impl.h:67-232) and set_diagonal (fem/assembler.h:644-686)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t find_col(const int32_t* __restrict__ cols, int64_t b, int64_t e, int32_t c)
{
  // std::lower_bound over the sorted columns of one row; -1 when absent
  int64_t lo = b, hi = e;
  while (lo < hi)
  {
    const int64_t mid = (lo + hi) >> 1;
    if (cols[mid] < c)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < e && cols[lo] == c) ? lo : -1;
}

// kind 0: insert_csr<dbs0,dbs1> (matrix bs == data bs); 1: insert_blocked_csr (matrix bs = 1);
// 2: insert_nonblocked_csr (data bs = 1, matrix bs = mbs0 x mbs1).  One thread per scalar of x.
__global__ void k_insert(int kind, int dbs0, int dbs1, int mbs0, int mbs1, const int64_t* __restrict__ row_ptr,
                         const int32_t* __restrict__ cols, double* __restrict__ data, const double* __restrict__ x,
                         const int32_t* __restrict__ xrows, int nr, const int32_t* __restrict__ xcols, int nc, int op,
                         int* __restrict__ err)
{
  const int64_t total = (int64_t)nr * dbs0 * nc * dbs1;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t ncs = (int64_t)nc * dbs1;
    const int64_t R = t / ncs, Cc = t - R * ncs; // scalar row / col within x
    const int r = (int)(R / dbs0), i = (int)(R - (int64_t)r * dbs0);
    const int c = (int)(Cc / dbs1), j = (int)(Cc - (int64_t)c * dbs1);
    const double v = x[t];
    int64_t pos;
    if (kind == 0)
    {
      const int32_t row = xrows[r];
      const int64_t d = find_col(cols, row_ptr[row], row_ptr[row + 1], xcols[c]);
      pos = d < 0 ? -1 : d * dbs0 * dbs1 + i * dbs1 + j;
    }
    else if (kind == 1)
    {
      const int32_t row = xrows[r] * dbs0 + i;
      const int64_t d = find_col(cols, row_ptr[row], row_ptr[row + 1], xcols[c] * dbs1);
      // the reference locates the first column of the block and writes dbs1 consecutive entries
      pos = d < 0 ? -1 : d + j;
    }
    else
    {
      const int32_t rq = xrows[r] / mbs0, rr = xrows[r] - rq * mbs0;
      const int32_t cq = xcols[c] / mbs1, cr = xcols[c] - cq * mbs1;
      const int64_t d = find_col(cols, row_ptr[rq], row_ptr[rq + 1], cq);
      pos = d < 0 ? -1 : d * mbs0 * mbs1 + rr * mbs1 + cr;
    }
    if (pos < 0)
    {
      *err = 1;
      continue;
    }
    if (op)
      red_add(data + pos, v);
    else
      data[pos] = v;
  }
}

__global__ void k_set_diagonal(const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols,
                               double* __restrict__ data, const int32_t* __restrict__ rows, int64_t n, int bs0, int bs1,
                               double diag, int* __restrict__ err)
{
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int32_t dof = rows[t];
    const int32_t rq = dof / bs0, rr = dof - rq * bs0;
    const int32_t cq = dof / bs1, cr = dof - cq * bs1;
    const int64_t d = find_col(cols, row_ptr[rq], row_ptr[rq + 1], cq);
    if (d < 0)
      *err = 1;
    else
      data[d * bs0 * bs1 + rr * bs1 + cr] = diag;
  }
}

__global__ void k_list_offdiag_rows(int32_t n, const int64_t* __restrict__ off_diag,
                                    const int64_t* __restrict__ row_ptr, int32_t* __restrict__ list,
                                    int32_t* __restrict__ count)
{
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (off_diag[i] < row_ptr[i + 1])
      list[atomicAdd(count, 1)] = i;
}
} // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
namespace bfx
{
int csr_finish_create(bfx_csr* A)
{
  // nnz, nnz of owned rows, and the list of owned rows that have off-diagonal (ghost-column) entries
  BFX_CUDA(cudaMemcpy(&A->nnz, A->row_ptr + A->n_rows_all, sizeof(int64_t), cudaMemcpyDeviceToHost));
  BFX_CUDA(cudaMemcpy(&A->nnz_owned, A->row_ptr + A->n_rows_owned, sizeof(int64_t), cudaMemcpyDeviceToHost));
  A->n_offdiag_rows = 0;
  A->offdiag_rows = nullptr;
  if (A->n_rows_owned > 0)
  {
    int32_t *list = nullptr, *count = nullptr;
    BFX_CUDA(cudaMalloc(&list, sizeof(int32_t) * A->n_rows_owned));
    BFX_CUDA(cudaMalloc(&count, sizeof(int32_t)));
    BFX_CUDA(cudaMemset(count, 0, sizeof(int32_t)));
    k_list_offdiag_rows<<<grid_for(A->n_rows_owned, 256, 8), 256>>>(A->n_rows_owned, A->off_diag, A->row_ptr, list,
                                                                    count);
    BFX_CHECK_LAUNCH();
    BFX_CUDA(cudaMemcpy(&A->n_offdiag_rows, count, sizeof(int32_t), cudaMemcpyDeviceToHost));
    BFX_CUDA(cudaFree(count));
    if (A->n_offdiag_rows > 0)
    {
      // sort for determinism / locality
      std::vector<int32_t> h(A->n_offdiag_rows);
      BFX_CUDA(cudaMemcpy(h.data(), list, sizeof(int32_t) * h.size(), cudaMemcpyDeviceToHost));
      std::sort(h.begin(), h.end());
      BFX_CUDA(cudaMalloc(&A->offdiag_rows, sizeof(int32_t) * h.size()));
      BFX_CUDA(cudaMemcpy(A->offdiag_rows, h.data(), sizeof(int32_t) * h.size(), cudaMemcpyHostToDevice));
    }
    BFX_CUDA(cudaFree(list));
  }
  BFX_CUDA(cudaMalloc(&A->err_flag, sizeof(int)));
  BFX_CUDA(cudaMemset(A->err_flag, 0, sizeof(int)));
  return BFX_OK;
}
} // namespace bfx

extern "C"
{
int bfx_csr_create(bfx_csr_t** out, int32_t n_rows_all, int32_t n_rows_owned, const int64_t* row_ptr,
                   const int32_t* cols, const int64_t* off_diag, int bs0, int bs1)
{
  BFX_REQUIRE(out && row_ptr && n_rows_all >= 0 && n_rows_owned >= 0 && n_rows_owned <= n_rows_all && bs0 > 0
                  && bs1 > 0,
              "bfx_csr_create: bad arguments");
  bfx_csr* A = new bfx_csr();
  A->n_rows_all = n_rows_all;
  A->n_rows_owned = n_rows_owned;
  A->bs0 = bs0;
  A->bs1 = bs1;
  int e = upload(&A->row_ptr, row_ptr, (size_t)n_rows_all + 1);
  if (e)
    return e;
  int64_t nnz = 0;
  BFX_CUDA(cudaMemcpy(&nnz, A->row_ptr + n_rows_all, sizeof(int64_t), cudaMemcpyDeviceToHost));
  if ((e = upload(&A->cols, cols, (size_t)nnz)))
    return e;
  if (off_diag)
  {
    if ((e = upload(&A->off_diag, off_diag, (size_t)n_rows_all)))
      return e;
  }
  else
  {
    // no ghost columns: off_diag = row end
    if ((e = dev_alloc(&A->off_diag, (size_t)n_rows_all)))
      return e;
    BFX_CUDA(cudaMemcpy(A->off_diag, A->row_ptr + 1, sizeof(int64_t) * n_rows_all, cudaMemcpyDeviceToDevice));
  }
  if ((e = csr_finish_create(A)))
    return e;
  *out = A;
  return BFX_OK;
}

int bfx_csr_destroy(bfx_csr_t* A)
{
  if (!A)
    return BFX_OK;
  cudaFree(A->row_ptr);
  cudaFree(A->cols);
  cudaFree(A->off_diag);
  cudaFree(A->offdiag_rows);
  cudaFree(A->err_flag);
  delete A;
  return BFX_OK;
}

int64_t bfx_csr_nnz(const bfx_csr_t* A) { return A ? A->nnz : -1; }

int bfx_csr_get_structure(const bfx_csr_t* A, int64_t* row_ptr, int32_t* cols, int64_t* off_diag)
{
  BFX_REQUIRE(A, "null csr");
  if (row_ptr)
    BFX_CUDA(cudaMemcpy(row_ptr, A->row_ptr, sizeof(int64_t) * ((size_t)A->n_rows_all + 1), cudaMemcpyDefault));
  if (cols && A->nnz)
    BFX_CUDA(cudaMemcpy(cols, A->cols, sizeof(int32_t) * (size_t)A->nnz, cudaMemcpyDefault));
  if (off_diag && A->n_rows_all)
    BFX_CUDA(cudaMemcpy(off_diag, A->off_diag, sizeof(int64_t) * (size_t)A->n_rows_all, cudaMemcpyDefault));
  return BFX_OK;
}

int bfx_csr_device_ptrs(const bfx_csr_t* A, const int64_t** row_ptr, const int32_t** cols, const int64_t** off_diag)
{
  BFX_REQUIRE(A, "null csr");
  if (row_ptr)
    *row_ptr = A->row_ptr;
  if (cols)
    *cols = A->cols;
  if (off_diag)
    *off_diag = A->off_diag;
  return BFX_OK;
}

int bfx_csr_set_spmv_variant(bfx_csr_t* A, int variant)
{
  BFX_REQUIRE(A && variant >= -2 && variant <= 2, "bfx_csr_set_spmv_variant: variant must be -2 (select by timing), -1 (by row length) .. 2");
  A->spmv_variant = variant;
  return BFX_OK;
}

int bfx_spmv(const bfx_csr_t* A, const double* values, const double* x, double* y, int part, bfx_stream_t stream)
{
  BFX_REQUIRE(A && ((values && x && y) || A->nnz == 0 || A->n_rows_owned == 0), "bfx_spmv: null argument");
  if (A->nnz == 0)
    return BFX_OK; // y += 0 (an empty rank: la/MatrixCSR.h:877-946 loops over no rows)
  const int32_t n = A->n_rows_owned;
  if (n == 0)
    return BFX_OK;
  cudaStream_t st = S(stream);
  const int64_t* rb = A->row_ptr;
  const int64_t* re = A->row_ptr + 1;
  const int32_t* row_list = nullptr;
  int32_t n_launch = n;
  if (part == BFX_SPMV_DIAG)
    re = A->off_diag;
  else if (part == BFX_SPMV_OFFDIAG)
  {
    rb = A->off_diag;
    row_list = A->offdiag_rows; // only rows that have ghost columns
    n_launch = A->n_offdiag_rows;
    if (n_launch == 0)
      return BFX_OK;
  }
  else if (part != BFX_SPMV_FULL)
    return fail(BFX_ERR_INVALID, "bfx_spmv: bad part %d", part);

  if (A->bs0 == 1 && A->bs1 == 1 && part != BFX_SPMV_OFFDIAG)
  {
    const double avg = (double)A->nnz_owned / (double)n;
    auto launch = [&](int variant, double* yy)
    {
      if (variant == 2)
      {
        // TMA-pipelined persistent CTAs: rows per tile so that one stage usually covers a tile; longer rows get
        // more lanes per row so that a CTA keeps 8-16 warps busy
        const size_t smem = 2 * SPMV_TMA_CAP * 12 + 64;
        const int per_sm = 2;
#define BFX_TMA_LAUNCH(RR, LL)                                                                                        \
  {                                                                                                                   \
    cudaFuncSetAttribute(k_spmv_tma<RR, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
    const unsigned grid = (unsigned)std::min<int64_t>(((int64_t)n + RR - 1) / RR, (int64_t)sm_count() * per_sm);     \
    k_spmv_tma<RR, LL><<<grid, RR * LL, smem, st>>>(n, A->nnz, A->row_ptr, re, A->cols, values, x, yy);               \
  }
        if (avg <= 15.0)
          BFX_TMA_LAUNCH(256, 1)
        else if (avg <= 30.0)
          BFX_TMA_LAUNCH(128, 4)
        else
          BFX_TMA_LAUNCH(64, 8)
#undef BFX_TMA_LAUNCH
        return;
      }
      if (variant == 1)
      {
        // row-per-thread out of staged (cols, values); rows per CTA so that one pass covers them
        if (avg <= 15.0)
          k_spmv_rows<256><<<(n + 255) / 256, 256, 0, st>>>(n, A->row_ptr, re, A->cols, values, x, yy);
        else if (avg <= 30.0)
          k_spmv_rows<128><<<(n + 127) / 128, 128, 0, st>>>(n, A->row_ptr, re, A->cols, values, x, yy);
        else
          k_spmv_rows<64><<<(n + 63) / 64, 64, 0, st>>>(n, A->row_ptr, re, A->cols, values, x, yy);
        return;
      }
      // entry-consecutive stream; rows per CTA chosen so that one pass of SPMV_CHUNK products covers them
      if (avg <= 14.0 * 1.15)
        k_spmv_stream<1><<<(n + 255) / 256, SPMV_THREADS, 0, st>>>(n, A->row_ptr, re, A->cols, values, x, yy);
      else if (avg <= 30.0 * 1.05)
        k_spmv_stream<2><<<(n + 127) / 128, SPMV_THREADS, 0, st>>>(n, A->row_ptr, re, A->cols, values, x, yy);
      else if (avg <= 62.0)
        k_spmv_stream<4><<<(n + 63) / 64, SPMV_THREADS, 0, st>>>(n, A->row_ptr, re, A->cols, values, x, yy);
      else
        k_spmv_stream<8><<<(n + 31) / 32, SPMV_THREADS, 0, st>>>(n, A->row_ptr, re, A->cols, values, x, yy);
    };
    if (A->spmv_variant == -1)
    {
      // default: chosen by the average row length alone, so that every rank and every run take the same kernel (the
      // kernels sum a row in different orders: a timed choice would make y differ in the last bits from run to run).
      // Short rows (P1: 15 entries) -> row per thread behind the TMA pipeline, long rows (P2: 29) -> entry stream;
      // measured at C2 / C3: 0.70 against 0.83 / 1.12 ms, 1.83 against 2.47 / 2.98 ms (profiles/r02_spmv_variants.txt)
      const_cast<bfx_csr*>(A)->spmv_variant = avg <= 20.0 ? 2 : 0;
    }
    if (A->spmv_variant < 0)
    {
      // opt-in autotuning (bfx_csr_set_spmv_variant(-2)): time the kernels into a scratch y, keep the fastest
      double* scratch = nullptr;
      cudaEvent_t e0, e1;
      if (cudaMalloc(&scratch, sizeof(double) * (size_t)n) == cudaSuccess && cudaEventCreate(&e0) == cudaSuccess
          && cudaEventCreate(&e1) == cudaSuccess)
      {
        float best = 1e30f;
        int best_v = 0;
        for (int v = 0; v < 3; ++v)
        {
          float ms = 1e30f;
          for (int rep = 0; rep < 3; ++rep)
          {
            cudaEventRecord(e0, st);
            launch(v, scratch);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float t = 0;
            cudaEventElapsedTime(&t, e0, e1);
            ms = t < ms ? t : ms;
          }
          if (ms < best)
          {
            best = ms;
            best_v = v;
          }
        }
        const_cast<bfx_csr*>(A)->spmv_variant = best_v;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
      }
      else
        const_cast<bfx_csr*>(A)->spmv_variant = 0;
      cudaFree(scratch);
      BFX_CHECK_LAUNCH();
    }
    launch(A->spmv_variant, y);
  }
  else
  {
    // bs = 3 (elasticity): two kernels, the faster one is kept per matrix (timed once, like the bs = 1 variants).
    // The TMA variant reads whole rows (full / diagonal part), not the row-list form of the off-diagonal pass.
    if (A->bs0 == 3 && A->bs1 == 3 && !row_list)
    {
      auto launch3 = [&](int variant, double* yy)
      {
        if (variant == 1)
        {
          const size_t smem = 2 * (size_t)SPMV_BTMA_CAP * (9 * 8 + 4) + 64;
          cudaFuncSetAttribute(k_spmv_blocked_tma<3, 3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          const int64_t ntiles = ((int64_t)n_launch + SPMV_BTMA_TR - 1) / SPMV_BTMA_TR;
          const unsigned g = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * 3);
          k_spmv_blocked_tma<3, 3, 16><<<g, 512, smem, st>>>(n_launch, A->nnz, A->row_ptr, rb, re, A->cols, values, x, yy);
        }
        else
          k_spmv_blocked<3, 3><<<grid_for((int64_t)n_launch * 32, 256, 32), 256, 0, st>>>(n_launch, rb, re, A->cols,
                                                                                           values, x, yy, nullptr);
      };
      if (A->spmv_variant == -1)
        const_cast<bfx_csr*>(A)->spmv_variant = 1; // deterministic default: the TMA-fed kernel (3.50 against 4.10 ms at C4)
      if (A->spmv_variant < 0)
      {
        double* scratch = nullptr;
        cudaEvent_t e0, e1;
        int best_v = 0;
        if (cudaMalloc(&scratch, sizeof(double) * 3 * (size_t)n_launch) == cudaSuccess
            && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess)
        {
          float best = 1e30f;
          for (int v = 0; v < 2; ++v)
          {
            float ms = 1e30f;
            for (int rep = 0; rep < 3; ++rep)
            {
              cudaEventRecord(e0, st);
              launch3(v, scratch);
              cudaEventRecord(e1, st);
              cudaEventSynchronize(e1);
              float t = 0;
              cudaEventElapsedTime(&t, e0, e1);
              ms = t < ms ? t : ms;
            }
            if (ms < best)
              best = ms, best_v = v;
          }
          cudaEventDestroy(e0);
          cudaEventDestroy(e1);
        }
        cudaFree(scratch);
        const_cast<bfx_csr*>(A)->spmv_variant = best_v;
        BFX_CHECK_LAUNCH();
      }
      launch3(A->spmv_variant > 1 ? 1 : A->spmv_variant, y);
      BFX_CHECK_LAUNCH();
      return BFX_OK;
    }
    const unsigned grid = grid_for((int64_t)n_launch * 32, 256, 32);
#define BLK(a, b)                                                                                                    \
  if (A->bs0 == a && A->bs1 == b)                                                                                    \
    k_spmv_blocked<a, b><<<grid, 256, 0, st>>>(n_launch, rb, re, A->cols, values, x, y, row_list);                   \
  else
    BLK(1, 1) BLK(2, 2) BLK(3, 3) BLK(1, 2) BLK(2, 1) BLK(2, 3) BLK(3, 2)
#undef BLK
    k_spmv_generic<<<grid, 256, 0, st>>>(n_launch, rb, re, A->cols, values, x, y, A->bs0, A->bs1, row_list);
  }
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

int bfx_spmvT(const bfx_csr_t* A, const double* values, const double* x, double* y, int part, bfx_stream_t stream)
{
  BFX_REQUIRE(A && ((values && x && y) || A->nnz == 0 || A->n_rows_owned == 0), "bfx_spmvT: null argument");
  if (A->nnz == 0)
    return BFX_OK;
  const int32_t n = A->n_rows_owned;
  if (n == 0)
    return BFX_OK;
  const int64_t* rb = A->row_ptr;
  const int64_t* re = A->row_ptr + 1;
  if (part == BFX_SPMV_DIAG)
    re = A->off_diag;
  else if (part == BFX_SPMV_OFFDIAG)
    rb = A->off_diag;
  k_spmvT<<<grid_for((int64_t)n * 32, 256, 32), 256, 0, S(stream)>>>(n, rb, re, A->cols, values, x, y, A->bs0,
                                                                      A->bs1);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

static int check_err_flag(const bfx_csr_t* A, cudaStream_t st, const char* what)
{
  int h = 0;
  BFX_CUDA(cudaMemcpyAsync(&h, A->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  BFX_CUDA(cudaStreamSynchronize(st));
  if (h)
  {
    BFX_CUDA(cudaMemsetAsync(A->err_flag, 0, sizeof(int), st));
    return fail(BFX_ERR_NOT_IN_SPARSITY, "%s: Entry not in sparsity", what);
  }
  return BFX_OK;
}

int bfx_csr_insert(const bfx_csr_t* A, double* values, int kind, int dbs0, int dbs1, const double* x,
                   const int32_t* xrows, int nr, const int32_t* xcols, int nc, int op, bfx_stream_t stream)
{
  BFX_REQUIRE(A && values && x && xrows && xcols && nr >= 0 && nc >= 0, "bfx_csr_insert: bad arguments");
  if (kind == 0)
    BFX_REQUIRE(dbs0 == A->bs0 && dbs1 == A->bs1, "insert_csr: data block size must equal matrix block size");
  else if (kind == 1)
    BFX_REQUIRE(A->bs0 == 1 && A->bs1 == 1, "insert_blocked_csr: matrix block size must be 1");
  else if (kind == 2)
    BFX_REQUIRE(dbs0 == 1 && dbs1 == 1, "insert_nonblocked_csr: data block size must be 1");
  else
    return fail(BFX_ERR_INVALID, "bfx_csr_insert: bad kind %d", kind);
  if (nr == 0 || nc == 0)
    return BFX_OK;
  cudaStream_t st = S(stream);
  const size_t nx = (size_t)nr * dbs0 * nc * dbs1;
  // not a hot path (explicit MatrixCSR::add/set calls): stage the operands on the device
  double* dx = nullptr;
  int32_t *dr = nullptr, *dc = nullptr;
  int e;
  if ((e = upload(&dx, x, nx, st)) || (e = upload(&dr, xrows, (size_t)nr, st)) || (e = upload(&dc, xcols, (size_t)nc, st)))
    return e;
  k_insert<<<grid_for((int64_t)nx, 256, 8), 256, 0, st>>>(kind, dbs0, dbs1, A->bs0, A->bs1, A->row_ptr, A->cols,
                                                          values, dx, dr, nr, dc, nc, op, A->err_flag);
  BFX_CHECK_LAUNCH();
  e = check_err_flag(A, st, "MatrixCSR insert");
  cudaFree(dx);
  cudaFree(dr);
  cudaFree(dc);
  return e;
}

int bfx_csr_set_diagonal(const bfx_csr_t* A, double* values, const int32_t* rows, int64_t n, double diag,
                         bfx_stream_t stream)
{
  BFX_REQUIRE(A && (values || A->nnz == 0), "bfx_csr_set_diagonal: null argument");
  if (n == 0)
    return BFX_OK;
  cudaStream_t st = S(stream);
  k_set_diagonal<<<grid_for(n, 256, 8), 256, 0, st>>>(A->row_ptr, A->cols, values, rows, n, A->bs0, A->bs1, diag,
                                                      A->err_flag);
  BFX_CHECK_LAUNCH();
  return check_err_flag(A, st, "set_diagonal");
}

int bfx_csr_squared_norm(const bfx_csr_t* A, const double* values, double* result_host, bfx_stream_t stream)
{
  BFX_REQUIRE(A && (values || A->nnz == 0) && result_host, "bfx_csr_squared_norm: null argument");
  if (A->nnz == 0)
  {
    *result_host = 0.0;
    return BFX_OK;
  }
  return reduce<0>(A->nnz_owned * A->bs0 * A->bs1, values, values, result_host, S(stream));
}

int bfx_dot(int64_t n, const double* x, const double* y, double* result_host, bfx_stream_t stream)
{
  return reduce<0>(n, x, y, result_host, S(stream));
}

int bfx_norm(int64_t n, const double* x, int type, double* result_host, bfx_stream_t stream)
{
  if (type == 0)
    return reduce<1>(n, x, x, result_host, S(stream));
  if (type == 1)
    return reduce<0>(n, x, x, result_host, S(stream));
  if (type == 2)
    return reduce<2>(n, x, x, result_host, S(stream));
  return fail(BFX_ERR_INVALID, "bfx_norm: bad type %d", type);
}

int bfx_axpy(int64_t n, double alpha, const double* x, double* y, bfx_stream_t stream)
{
  if (n <= 0)
    return BFX_OK;
  k_axpy<<<grid_for(n, 256, 16), 256, 0, S(stream)>>>(n, alpha, x, y);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

int bfx_fill(int64_t n, double value, double* x, bfx_stream_t stream)
{
  if (n <= 0)
    return BFX_OK;
  BFX_REQUIRE(x, "bfx_fill: null array");
  if (value == 0.0)
  {
    BFX_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * (size_t)n, S(stream)));
    return BFX_OK;
  }
  k_fill<<<grid_for(n, 256, 16), 256, 0, S(stream)>>>(n, value, x);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}
}
