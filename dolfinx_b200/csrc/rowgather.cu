// Row-gather assembly of the Q1-hexahedron linear-elasticity matrix (strategy BFX_ASM_ROWGATHER):
// fem::impl::assemble_cells_matrix (fem/assemble_matrix_impl.h:92-200) + MatrixCSR::add with
// BS0 = BS1 = 3 (la/MatrixCSR.h:310-335, la/matrix_csr_impl.h:67-109) for kernel
// BFX_K_ELASTICITY_Q1_HEX_A (python/demo/demo_elasticity.py:131-150).
//
// The cell-parallel kernel issues 576 fp64 REDs per cell, at 1.2-1.9 SM-cycles per RED lane
// (profiles/r01_microbench_red_bulk_lds.txt).  Here every CSR value is written ONCE by plain
// coalesced stores and no atomics are used, so the result is bitwise reproducible:
//   pre-pass   one thread per cell: K = J^{-1} and |det J| at the cell centre (80 bytes per cell);
//              cells that are not parallelepipeds are flagged and listed;
//   main       a CTA owns 32 consecutive block rows, 8 lanes per row.  For every cell incident to
//              the row (transposed dofmap, built once) lane j re-computes the 3x3 tensor
//              D_ij = int grad(phi_i) (x) grad(phi_j) from the cell record (pre-integrated reference
//              tensor, 81 FMAs) and adds it to the row's block accumulators in shared memory - the 8
//              lanes of a row hit 8 different blocks, rows are private to their lanes, so no atomics;
//   flush      block -> mu (tr D I + D^T) + lambda D (linear in D, applied once per block), bc rows /
//              columns zeroed, then the tile's values leave as one contiguous coalesced stream;
//   fallback   the flagged (non-affine) cells are added by the RED kernel (2x2x2 Gauss).
// The element work is done once per (row, cell) instead of once per cell (the symmetric half is not
// shared between rows): 2x the flops of the cell-parallel kernel, bought back many times by the
// missing atomics.
#include "asm_device.cuh"
#include "elements.cuh"
#include <cub/device/device_scan.cuh>

using namespace bfx;

namespace
{
constexpr int RG_ROWS = 32;        // block rows per CTA
constexpr int RG_THREADS = RG_ROWS * 8;
constexpr int RG_CAP = RG_ROWS * 27; // block accumulators per pass (27 = interior row of a hex mesh)
constexpr int RG_STRIDE = 10;      // doubles per accumulator (9 + 1 pad: 16-byte aligned 128-bit accesses)

struct RGArgs
{
  int32_t n_rows;
  const int64_t* row_ptr;
  const int32_t* cols;
  const int64_t* tptr;   // transposed dofmap: row -> [tptr[r], tptr[r+1])
  const uint32_t* tent;  // (entity << 3) | local node, ascending entity inside a row
  const char* pos;
  int pos_bytes, pos_stride;
  const double* rec; // per entity: K (9, row-major), |det J| (< 0: not a parallelepiped)
  const int8_t *bc0, *bc1;
  double mu, lmbda;
  double* values;
  int overwrite;
};

__global__ void k_rg_max_row_len(int32_t n, const int64_t* __restrict__ row_ptr, int* __restrict__ out)
{
  int m = 0;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    m = max(m, (int)(row_ptr[i + 1] - row_ptr[i]));
  for (int o = 16; o > 0; o >>= 1)
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0)
    atomicMax(out, m);
}

// ---- plan: transposed dofmap with local node ids ---------------------------------------------------
__global__ void k_rg_count(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ dofmap,
                           int64_t* __restrict__ counts)
{
  const int64_t total = n * 8;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t e = t >> 3;
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    atomicAdd(reinterpret_cast<unsigned long long*>(counts + dofmap[(int64_t)cell * 8 + (t & 7)]), 1ULL);
  }
}

__global__ void k_rg_fill(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ dofmap,
                          const int64_t* __restrict__ tptr, int32_t* __restrict__ cursor, uint32_t* __restrict__ tent)
{
  const int64_t total = n * 8;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t e = t >> 3;
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    const int32_t r = dofmap[(int64_t)cell * 8 + (t & 7)];
    tent[tptr[r] + atomicAdd(cursor + r, 1)] = ((uint32_t)e << 3) | (uint32_t)(t & 7);
  }
}

// fixed summation order: ascending entity (= the order the CPU loop visits the cells)
__global__ void k_rg_sort(int32_t n_rows, const int64_t* __restrict__ tptr, uint32_t* __restrict__ tent)
{
  for (int32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x)
  {
    const int64_t b = tptr[r], e = tptr[r + 1];
    for (int64_t i = b + 1; i < e; ++i)
    {
      const uint32_t v = tent[i];
      int64_t j = i - 1;
      while (j >= b && tent[j] > v)
      {
        tent[j + 1] = tent[j];
        --j;
      }
      tent[j + 1] = v;
    }
  }
}

// ---- pre-pass: cell records ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    k_rg_records(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ x_dofmap,
                 const double* __restrict__ x, double* __restrict__ rec, int32_t* __restrict__ na_cells,
                 unsigned long long* __restrict__ na_count)
{
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
  {
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    int32_t xd[8];
    load_ints<8>(x_dofmap + (int64_t)cell * 8, xd);
    double xc[8][3];
    gather_coords<8>(x, xd, xc);
    // parallelepiped test (same criterion as the cell-parallel kernel)
    double dev2 = 0.0, h2 = 0.0;
#pragma unroll
    for (int m = 0; m < 3; ++m)
    {
      const double e1 = xc[1][m] - xc[0][m], e2 = xc[2][m] - xc[0][m], e3 = xc[4][m] - xc[0][m];
      const double d3 = xc[3][m] - xc[0][m] - e1 - e2, d5 = xc[5][m] - xc[0][m] - e1 - e3;
      const double d6 = xc[6][m] - xc[0][m] - e2 - e3, d7 = xc[7][m] - xc[0][m] - e1 - e2 - e3;
      dev2 += d3 * d3 + d5 * d5 + d6 * d6 + d7 * d7;
      h2 += e1 * e1 + e2 * e2 + e3 * e3;
    }
    const bool affine = dev2 <= 1e-26 * h2;
    const double X[3] = {0.5, 0.5, 0.5};
    double K[3][3];
    const double det = el::HexQ1::jacobian_inverse(xc, X, K);
    double2* out = reinterpret_cast<double2*>(rec + e * RG_STRIDE);
    out[0] = make_double2(K[0][0], K[0][1]);
    out[1] = make_double2(K[0][2], K[1][0]);
    out[2] = make_double2(K[1][1], K[1][2]);
    out[3] = make_double2(K[2][0], K[2][1]);
    out[4] = make_double2(K[2][2], affine ? fabs(det) : -1.0);
    if (!affine)
      na_cells[atomicAdd(na_count, 1ULL)] = cell;
  }
}

// D[p][b] = |det| sum_cd K[c][p] That_ij[c][d] K[d][b], That from the 1-D integrals of N0 = 1 - s, N1 = s:
// mass 1/3 | 1/6, stiffness +1 | -1, mixed +-1/2
__device__ __forceinline__ void q1_affine_D(int i, int j, const double (&K)[3][3], double adet, double (&D)[3][3])
{
  double Mm[3], Ss[3], Cij[3], Cji[3];
#pragma unroll
  for (int m = 0; m < 3; ++m)
  {
    const int bi = (i >> m) & 1, bj = (j >> m) & 1;
    Mm[m] = bi == bj ? (1.0 / 3.0) : (1.0 / 6.0);
    Ss[m] = bi == bj ? 1.0 : -1.0;
    Cij[m] = bi ? 0.5 : -0.5;
    Cji[m] = bj ? 0.5 : -0.5;
  }
  double M1[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
  {
    double T[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      T[d] = c == d ? Ss[c] * Mm[(c + 1) % 3] * Mm[(c + 2) % 3] : Cij[c] * Cji[d] * Mm[3 - c - d];
#pragma unroll
    for (int b = 0; b < 3; ++b)
      M1[c][b] = T[0] * K[0][b] + T[1] * K[1][b] + T[2] * K[2][b];
  }
#pragma unroll
  for (int p = 0; p < 3; ++p)
#pragma unroll
    for (int b = 0; b < 3; ++b)
      D[p][b] = adet * (K[0][p] * M1[0][b] + K[1][p] * M1[1][b] + K[2][p] * M1[2][b]);
}

template <typename PosT>
__global__ void __launch_bounds__(RG_THREADS, 2) k_q1_rowgather(const RGArgs g)
{
  extern __shared__ __align__(16) double acc[]; // RG_CAP accumulators of RG_STRIDE doubles
  __shared__ __align__(16) double s_rec[RG_ROWS][8][RG_STRIDE];
  __shared__ uint16_t s_pos[RG_ROWS][8][8 + 2]; // +2: the 4 rows of a warp read different banks
  __shared__ int s_i[RG_ROWS][8];
  __shared__ int s_re;
  const int grp = threadIdx.x >> 3, j = threadIdx.x & 7;
  const unsigned gmask = 0xffu << (8 * (grp & 3)); // the 8 lanes of this row
  const int32_t r0 = blockIdx.x * RG_ROWS;
  const int32_t r1 = min(r0 + RG_ROWS, g.n_rows);
  if (!g.overwrite && g.tptr[r1] == g.tptr[r0])
    return; // no cell of this plan touches the tile (cell subsets)

  for (int32_t rs = r0; rs < r1;)
  {
    // rows [rs, re) of this pass: as many as fit the accumulators
    if (threadIdx.x == 0)
    {
      int32_t re = rs + 1;
      const int64_t b0 = g.row_ptr[rs];
      while (re < r1 && g.row_ptr[re + 1] - b0 <= RG_CAP)
        ++re;
      s_re = re;
    }
    for (int t = threadIdx.x; t < RG_CAP * RG_STRIDE / 2; t += RG_THREADS)
      reinterpret_cast<double2*>(acc)[t] = make_double2(0.0, 0.0);
    __syncthreads();
    const int32_t re = s_re;
    const int64_t b0 = g.row_ptr[rs];
    const int nb = (int)(g.row_ptr[re] - b0);

    // ---- gather: 8 lanes per row.  The records and positions of up to 8 incident cells are fetched
    //      by the 8 lanes at once (one global latency per batch), then consumed from shared memory.
    const int32_t r = rs + grp;
    if (r < re)
    {
      double* rowacc = acc + (g.row_ptr[r] - b0) * RG_STRIDE;
      const int64_t te = g.tptr[r + 1];
      for (int64_t t0 = g.tptr[r]; t0 < te; t0 += 8)
      {
        const int nbatch = (int)min((int64_t)8, te - t0);
        if (j < nbatch)
        {
          const uint32_t ent = g.tent[t0 + j];
          const int64_t e = ent >> 3;
          const int i = (int)(ent & 7u);
          const double2* rp = reinterpret_cast<const double2*>(g.rec + e * RG_STRIDE);
          double2* sp = reinterpret_cast<double2*>(&s_rec[grp][j][0]);
          const double2 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3), q4 = __ldg(rp + 4);
          const PosT* pp = reinterpret_cast<const PosT*>(g.pos + e * g.pos_stride) + i * 8;
          PosT pr[8];
          if constexpr (sizeof(PosT) == 1)
            *reinterpret_cast<uint2*>(pr) = __ldg(reinterpret_cast<const uint2*>(pp));
          else
            *reinterpret_cast<uint4*>(pr) = __ldg(reinterpret_cast<const uint4*>(pp));
          sp[0] = q0, sp[1] = q1, sp[2] = q2, sp[3] = q3, sp[4] = q4;
#pragma unroll
          for (int m = 0; m < 8; ++m)
            s_pos[grp][j][m] = (uint16_t)pr[m];
          s_i[grp][j] = i;
        }
        __syncwarp(gmask);
        for (int k = 0; k < nbatch; ++k)
        {
          const double2* sp = reinterpret_cast<const double2*>(&s_rec[grp][k][0]);
          const double2 q0 = sp[0], q1 = sp[1], q2 = sp[2], q3 = sp[3], q4 = sp[4];
          if (q4.y >= 0.0)
          {
            const double K[3][3] = {{q0.x, q0.y, q1.x}, {q1.y, q2.x, q2.y}, {q3.x, q3.y, q4.x}};
            double D[3][3];
            q1_affine_D(s_i[grp][k], j, K, q4.y, D);
            const uint32_t p = s_pos[grp][k][j];
            double2* a2 = reinterpret_cast<double2*>(rowacc + (size_t)p * RG_STRIDE);
            double2 v0 = a2[0], v1 = a2[1], v2 = a2[2], v3 = a2[3];
            double v4 = rowacc[(size_t)p * RG_STRIDE + 8];
            v0.x += D[0][0], v0.y += D[0][1], v1.x += D[0][2];
            v1.y += D[1][0], v2.x += D[1][1], v2.y += D[1][2];
            v3.x += D[2][0], v3.y += D[2][1], v4 += D[2][2];
            a2[0] = v0, a2[1] = v1, a2[2] = v2, a2[3] = v3;
            rowacc[(size_t)p * RG_STRIDE + 8] = v4;
          }
          __syncwarp(gmask); // the next cell may touch a block another lane of this row just updated
        }
      }
    }
    __syncthreads();

    // ---- blocks: D -> mu (tr D I + D^T) + lambda D, bc rows / columns zeroed (in place)
    if (r < re)
    {
      const int64_t rb = g.row_ptr[r];
      const int len = (int)(g.row_ptr[r + 1] - rb);
      double* rowacc = acc + (rb - b0) * RG_STRIDE;
      unsigned zr = 0;
      if (g.bc0)
        zr = (g.bc0[3 * (int64_t)r] ? 1u : 0u) | (g.bc0[3 * (int64_t)r + 1] ? 2u : 0u) | (g.bc0[3 * (int64_t)r + 2] ? 4u : 0u);
      for (int p = j; p < len; p += 8)
      {
        double* a = rowacc + (size_t)p * RG_STRIDE;
        double D[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int l = 0; l < 3; ++l)
            D[k][l] = a[3 * k + l];
        unsigned zc = 0;
        if (g.bc1)
        {
          const int64_t c = g.cols[rb + p];
          zc = (g.bc1[3 * c] ? 1u : 0u) | (g.bc1[3 * c + 1] ? 2u : 0u) | (g.bc1[3 * c + 2] ? 4u : 0u);
        }
        const double tr = D[0][0] + D[1][1] + D[2][2];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int l = 0; l < 3; ++l)
          {
            const double v = g.mu * ((k == l ? tr : 0.0) + D[l][k]) + g.lmbda * D[k][l];
            a[3 * k + l] = (((zr >> k) | (zc >> l)) & 1u) ? 0.0 : v;
          }
      }
    }
    __syncthreads();

    // ---- the tile's values leave as one contiguous stream (add mode: 8 loads in flight per thread)
    double* out = g.values + b0 * 9;
    const int total = nb * 9;
    if (g.overwrite)
    {
      for (int t = threadIdx.x; t < total; t += RG_THREADS)
      {
        const int b = t / 9;
        out[t] = acc[b * RG_STRIDE + (t - 9 * b)];
      }
    }
    else
    {
      for (int t0 = threadIdx.x; t0 < total; t0 += 8 * RG_THREADS)
      {
        double old[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
        {
          const int t = t0 + u * RG_THREADS;
          old[u] = t < total ? out[t] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
        {
          const int t = t0 + u * RG_THREADS;
          if (t < total)
          {
            const int b = t / 9;
            out[t] = old[u] + acc[b * RG_STRIDE + (t - 9 * b)];
          }
        }
      }
    }
    __syncthreads();
    rs = re;
  }
}
} // namespace

namespace bfx
{
void free_rowgather(bfx_rowgather* g)
{
  if (!g)
    return;
  cudaFree(g->tptr);
  cudaFree(g->tent);
  cudaFree(g->rec);
  cudaFree(g->na_cells);
  cudaFree(g->na_count);
  delete g;
}

int launch_rowgather_q1(const bfx_asm* P, const AsmArgs& a, int values_mode, cudaStream_t st)
{
  const bfx_rowgather* G = P->rowgather;
  if (!G)
    return fail(BFX_ERR_INVALID, "BFX_ASM_ROWGATHER needs bfx_asm_build_rowgather() on the plan first");
  if (a.dofmap1 != a.dofmap0)
    return fail(BFX_ERR_UNSUPPORTED, "row-gather assembly needs identical test and trial dofmaps");
  const bfx_csr* csr = P->csr;
  // pre-pass: cell records + list of cells that are not parallelepipeds
  BFX_CUDA(cudaMemsetAsync(G->na_count, 0, sizeof(unsigned long long), st));
  if (P->ncells > 0)
    k_rg_records<<<grid_for(P->ncells, 128, 0), 128, 0, st>>>(P->ncells, P->cells, P->x_dofmap, a.x, G->rec, G->na_cells,
                                                               G->na_count);
  RGArgs g;
  g.n_rows = csr->n_rows_all;
  g.row_ptr = csr->row_ptr;
  g.cols = csr->cols;
  g.tptr = G->tptr;
  g.tent = G->tent;
  g.pos = P->pos;
  g.pos_bytes = P->pos_bytes;
  g.pos_stride = P->pos_stride;
  g.rec = G->rec;
  g.bc0 = a.bc0;
  g.bc1 = a.bc1;
  g.mu = a.constants[0];
  g.lmbda = a.constants[1];
  g.values = a.values;
  g.overwrite = values_mode == BFX_VALUES_OVERWRITE;
  const size_t smem = sizeof(double) * RG_CAP * RG_STRIDE;
  const unsigned grid = (unsigned)((csr->n_rows_all + RG_ROWS - 1) / RG_ROWS);
  if (grid > 0)
  {
    if (P->pos_bytes == 1)
    {
      BFX_CUDA(cudaFuncSetAttribute(k_q1_rowgather<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_q1_rowgather<uint8_t><<<grid, RG_THREADS, smem, st>>>(g);
    }
    else
    {
      BFX_CUDA(cudaFuncSetAttribute(k_q1_rowgather<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_q1_rowgather<uint16_t><<<grid, RG_THREADS, smem, st>>>(g);
    }
  }
  BFX_CHECK_LAUNCH();
  // general trilinear cells: 2x2x2 Gauss through the RED kernel, entries located by binary search
  AsmArgs b = a;
  b.cells = G->na_cells;
  b.n = P->ncells;
  b.n_dev = G->na_count;
  b.pos = nullptr;
  return launch_q1_red(P, b, st);
}
} // namespace bfx

extern "C"
{
int bfx_asm_build_rowgather(bfx_asm_t* P, bfx_stream_t stream)
{
  BFX_REQUIRE(P && P->csr && P->pos, "bfx_asm_build_rowgather: plan has no matrix / position map");
  const bfx_csr* csr = P->csr;
  if (!(P->nd0 == 8 && P->nd1 == 8 && P->nx == 8 && csr->bs0 == 3 && csr->bs1 == 3))
    return fail(BFX_ERR_UNSUPPORTED, "row-gather assembly is implemented for Q1 hexahedra with block size 3");
  if (P->dofmap1 && P->dofmap1 != P->dofmap0)
    return fail(BFX_ERR_UNSUPPORTED, "row-gather assembly needs identical test and trial dofmaps");
  if (P->ncells >= (1LL << 29))
    return fail(BFX_ERR_UNSUPPORTED, "row-gather assembly: more than 2^29 cells");
  cudaStream_t st = S(stream);
  // rows longer than the accumulator file cannot be processed
  {
    int* d_max = nullptr;
    int h_max = 0;
    BFX_CUDA(cudaMalloc(&d_max, sizeof(int)));
    BFX_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
    k_rg_max_row_len<<<grid_for(csr->n_rows_all, 256, 8), 256, 0, st>>>(csr->n_rows_all, csr->row_ptr, d_max);
    BFX_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_max);
    if (h_max > RG_CAP)
      return fail(BFX_ERR_UNSUPPORTED, "row-gather assembly: a row has %d blocks (limit %d)", h_max, RG_CAP);
  }
  free_rowgather(P->rowgather);
  P->rowgather = nullptr;
  bfx_rowgather* G = new bfx_rowgather();
  int e = BFX_OK;
  auto bail = [&](int status)
  {
    free_rowgather(G);
    return status;
  };
  const int32_t n_rows = csr->n_rows_all;
  const int64_t n = P->ncells;
  if ((e = dev_alloc(&G->tptr, (size_t)n_rows + 1)) || (e = dev_alloc(&G->tent, (size_t)n * 8 + 1))
      || (e = dev_alloc(&G->rec, (size_t)n * RG_STRIDE + 2)) || (e = dev_alloc(&G->na_cells, (size_t)n + 1))
      || (e = dev_alloc(&G->na_count, 1)))
    return bail(e);
  BFX_CUDA(cudaMemsetAsync(G->tptr, 0, sizeof(int64_t) * ((size_t)n_rows + 1), st));
  if (n > 0)
    k_rg_count<<<grid_for(n * 8, 256, 16), 256, 0, st>>>(n, P->cells, P->dofmap0, G->tptr);
  {
    void* tmp = nullptr;
    size_t bytes = 0;
    BFX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, G->tptr, G->tptr, n_rows + 1, st));
    BFX_CUDA(cudaMalloc(&tmp, bytes ? bytes : 1));
    BFX_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, G->tptr, G->tptr, n_rows + 1, st));
    BFX_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
  }
  if (n > 0)
  {
    int32_t* cursor = nullptr;
    if ((e = dev_alloc(&cursor, (size_t)n_rows)))
      return bail(e);
    BFX_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)n_rows, st));
    k_rg_fill<<<grid_for(n * 8, 256, 16), 256, 0, st>>>(n, P->cells, P->dofmap0, G->tptr, cursor, G->tent);
    k_rg_sort<<<grid_for(n_rows, 128, 0), 128, 0, st>>>(n_rows, G->tptr, G->tent);
    BFX_CHECK_LAUNCH();
    BFX_CUDA(cudaStreamSynchronize(st));
    cudaFree(cursor);
  }
  P->rowgather = G;
  return BFX_OK;
}
}
