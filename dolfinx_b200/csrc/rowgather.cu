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
#include <algorithm>
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
// One thread per cell; the 80-byte records of a warp's 32 consecutive cells leave as five fully coalesced 512-byte
// stores through a shared-memory transpose (the per-thread version wrote 16-byte pieces at an 80-byte stride).
// cmask != NULL: the column Dirichlet masks of the cell's 8 nodes too (3 bits per local node, from the per-node mask
// bytes) - a block's contributions all name its column node as (cell, j), so the block-gather kernel reads the block's
// column mask from the record of its first contribution (no pass over cols, no scattered byte loads).
__global__ void __launch_bounds__(128)
    k_rg_records(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ x_dofmap,
                 const double* __restrict__ x, double* __restrict__ rec, int32_t* __restrict__ na_cells,
                 unsigned long long* __restrict__ na_count, const int32_t* __restrict__ dofmap,
                 const uint8_t* __restrict__ node_mask, uint32_t* __restrict__ cmask)
{
  __shared__ double2 s_out[4][32 * 5];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t n_pad = (n + 31) & ~(int64_t)31; // (whole warps stay in the loop: the stores are cooperative)
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_pad; e += (int64_t)gridDim.x * blockDim.x)
  {
    if (e < n)
    {
      const int32_t cell = cells ? cells[e] : (int32_t)e;
      int32_t xd[8];
      load_ints<8>(x_dofmap + (int64_t)cell * 8, xd);
      uint32_t cm = 0;
      if (cmask)
      {
        int32_t d[8];
        load_ints<8>(dofmap + (int64_t)cell * 8, d);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          cm |= (uint32_t)__ldg(node_mask + d[j]) << (3 * j);
      }
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
      double2* out = s_out[wib] + lane * 5;
      out[0] = make_double2(K[0][0], K[0][1]);
      out[1] = make_double2(K[0][2], K[1][0]);
      out[2] = make_double2(K[1][1], K[1][2]);
      out[3] = make_double2(K[2][0], K[2][1]);
      out[4] = make_double2(K[2][2], affine ? fabs(det) : -1.0);
      if (cmask)
        cmask[e] = cm;
      if (!affine)
        na_cells[atomicAdd(na_count, 1ULL)] = cell;
    }
    __syncwarp();
    const int64_t e0 = e - lane;
    const int cnt = (int)min((int64_t)32, n - e0) * 5;
    double2* gout = reinterpret_cast<double2*>(rec + e0 * RG_STRIDE);
#pragma unroll
    for (int w = 0; w < 5; ++w)
      if (w * 32 + lane < cnt)
        gout[w * 32 + lane] = s_out[wib][w * 32 + lane];
    __syncwarp();
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

// The same tensor ADDED to A, written for instruction count (block-gather kernel): with s_i[c] = +-1 the sign of
// d(phi_i)/d(xi_c) and M_m = 1/3 | 1/6 (nodes i, j agree | differ in direction m),
//   That_ij[c][d] = s_i[c] s_j[d] W[c][d],  W[c][c] = M_{c+1} M_{c+2},  W[c][d] = M_{3-c-d} / 4  (c != d),
// so D = |det| sum_c K[c][:]^T (sum_d W'[c][d] K[d][:]) with the signs folded into W' by sign-bit flips; 1/3 and 1/6
// differ in one exponent bit, so M_m costs one integer instruction.  9 + 27 + 27 fp64 operations.
__device__ __forceinline__ void q1_affine_D_add(int i, int j, const double (&K)[3][3], double adet, double (&A)[3][3])
{
  const int e = i ^ j;
  double M[3];
  M[0] = __hiloint2double(0x3FD55555 - ((e << 20) & 0x100000), 0x55555555); // 1/3 or 1/6
  M[1] = __hiloint2double(0x3FD55555 - ((e << 19) & 0x100000), 0x55555555);
  M[2] = __hiloint2double(0x3FD55555 - ((e << 18) & 0x100000), 0x55555555);
  const double a0 = adet * M[0], a1 = adet * M[1], aq = 0.25 * adet;
  double W[3][3];
  W[0][0] = a1 * M[2];
  W[1][1] = a0 * M[2];
  W[2][2] = a0 * M[1];
  W[0][1] = W[1][0] = aq * M[2];
  W[0][2] = W[2][0] = 0.25 * a1;
  W[1][2] = W[2][1] = 0.25 * a0;
  // signs: s_i[c] s_j[d] = -1 iff bit c of i differs from bit d of j (s = +1 for a set bit)
  const int si[3] = {(i << 31) & (int)0x80000000, (i << 30) & (int)0x80000000, (i << 29) & (int)0x80000000};
  const int sj[3] = {(j << 31) & (int)0x80000000, (j << 30) & (int)0x80000000, (j << 29) & (int)0x80000000};
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int d = 0; d < 3; ++d)
      W[c][d] = __hiloint2double(__double2hiint(W[c][d]) ^ si[c] ^ sj[d], __double2loint(W[c][d]));
#pragma unroll
  for (int c = 0; c < 3; ++c)
  {
    double V[3];
#pragma unroll
    for (int b = 0; b < 3; ++b)
      V[b] = W[c][0] * K[0][b] + W[c][1] * K[1][b] + W[c][2] * K[2][b];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        A[p][b] = fma(K[c][p], V[b], A[p][b]);
  }
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

// =====================================================================================================
// Block-gather kernel (round 2).  ncu of k_q1_rowgather (profiles/r01_prof_q1_192_rowgather.csv): 17 345 instructions
// per cell, 160 of 300 shared-memory wavefronts per cell in the 72-byte read-modify-writes of the accumulators, every
// D_ij recomputed by a lane that shares its row with 7 others behind __syncwarp.  Here ONE THREAD OWNS ONE CSR BLOCK:
// its 9 accumulators live in registers, it walks the precomputed list of the block's (cell, i, j) contributions (1, 2,
// 4 or 8 on a structured mesh), reads the cell record (K, |det J|) from the tile's shared-memory copy and applies
// mu (tr D I + D^T) + lambda D once.  Whole blocks are dealt to the threads at plan time so that every thread of a tile
// runs the same number of contribution steps (k_bg_plan, longest-processing-time rule).  The finished blocks are written into a shared-memory image of the tile's contiguous value
// range, which leaves the SM as ONE TMA bulk store (cp.async.bulk.global.shared::cta) - no store instructions, no
// atomics, every value written once, bitwise reproducible.
// =====================================================================================================
constexpr int BG_ROWS = 16;     // block rows per tile
constexpr int BG_THREADS = 128;
constexpr int BG_NBCAP = 512;   // blocks per tile the plan accepts (rank and block index travel in 10 bits)
constexpr int BG_CCAP = 128;    // distinct cells per tile
constexpr int BG_KMAX = 16;     // contributions per block
constexpr int BG_INCCAP = 256;  // (row, cell) incidences per tile
constexpr int BG_SCAP = 16;     // contribution steps per thread (step lists: [tile][BG_SCAP][BG_THREADS])
constexpr int BG_QCAP = 8;      // blocks per thread (block lists: [tile][BG_QCAP][BG_THREADS])
// step-list entry (16 bits): j | i << 3 | cell slot << 6 | BG_LAST (the block's last contribution: finish and write
// it) | BG_EMPTY (a block no cell of the plan touches: write zeros); BG_IDLE = no further step for this thread
constexpr uint32_t BG_LAST = 0x2000u, BG_EMPTY = 0x4000u, BG_IDLE = 0xFFFFu;

// per-tile header of the block-gather plan (64 bytes)
struct BGTile
{
  int64_t b0;             // first block of the tile (row_ptr[r0])
  int32_t nb, ninc, nuc;  // blocks, (row, cell) incidences, distinct cells
  int32_t nsteps;         // longest step list of a thread
  uint16_t nlev[BG_KMAX]; // blocks with more than k contributions
  int32_t nq;             // longest block list of a thread
  int32_t pad2;
};
static_assert(sizeof(BGTile) == 64, "tile header layout");

struct BGPlanArgs
{
  int32_t n_rows;
  const int64_t* row_ptr;
  const int64_t* tptr;
  const uint32_t* tent;
  const char* pos;
  int pos_bytes, pos_stride;
  int32_t* cells;  // [tile][BG_CCAP]
  uint16_t* perm;  // [tile][BG_QCAP][BG_THREADS]: q-th block of thread t (block index in tile | row in tile << 10)
  uint16_t* ent;   // [tile][BG_SCAP][BG_THREADS]: s-th contribution step of thread t
  BGTile* tiles;
  int* err;     // 1: tile over a capacity
  int* maxima;  // [0] cells per tile, [1] blocks per tile, [2] steps per thread, [3] blocks per thread
};

__global__ void __launch_bounds__(BG_THREADS) k_bg_plan(const BGPlanArgs g)
{
  __shared__ uint32_t s_cand[BG_INCCAP];
  __shared__ int32_t s_ucell[BG_CCAP];
  __shared__ uint8_t s_first[BG_INCCAP];
  __shared__ uint8_t s_cnt[BG_NBCAP], s_row[BG_NBCAP];
  __shared__ uint16_t s_ent[BG_NBCAP][BG_KMAX];
  __shared__ uint16_t s_rank[BG_NBCAP];
  __shared__ int s_hist[BG_KMAX + 2], s_start[BG_KMAX + 2], s_wsum[BG_THREADS / 32], s_nuc, s_bad;
  __shared__ uint16_t s_byrank[BG_NBCAP];
  __shared__ uint8_t s_owner[BG_NBCAP], s_q[BG_NBCAP], s_s0[BG_NBCAP];
  __shared__ int s_sum[BG_THREADS], s_nblk[BG_THREADS], s_S, s_Q;
  __shared__ uint16_t s_steps[BG_SCAP * BG_THREADS], s_blocks[BG_QCAP * BG_THREADS];
  const int tid = threadIdx.x;
  const int64_t tile = blockIdx.x;
  const int32_t r0 = (int32_t)(tile * BG_ROWS), r1 = min(r0 + BG_ROWS, g.n_rows);
  const int64_t b0 = g.row_ptr[r0];
  const int nb = (int)(g.row_ptr[r1] - b0);
  const int64_t t0 = g.tptr[r0];
  const int ninc = (int)(g.tptr[r1] - t0);
  if (tid == 0)
  {
    s_nuc = 0;
    s_bad = 0;
  }
  if (nb > BG_NBCAP || ninc > BG_INCCAP)
  {
    if (tid == 0)
      *g.err = 1;
    return;
  }
  for (int k = tid; k < BG_NBCAP; k += BG_THREADS)
    s_cnt[k] = 0;
  for (int k = tid; k < BG_KMAX + 2; k += BG_THREADS)
    s_hist[k] = 0;
  for (int k = tid; k < ninc; k += BG_THREADS)
    s_cand[k] = g.tent[t0 + k] >> 3;
  __syncthreads();
  // ---- distinct cells of the tile in ascending entity order
  for (int k = tid; k < ninc; k += BG_THREADS)
  {
    bool first = true;
    for (int m = 0; m < k && first; ++m)
      first = s_cand[m] != s_cand[k];
    s_first[k] = first;
    if (first)
      atomicAdd(&s_nuc, 1);
  }
  __syncthreads();
  const int nuc = s_nuc;
  if (nuc > BG_CCAP)
  {
    if (tid == 0)
      *g.err = 1;
    return;
  }
  for (int k = tid; k < ninc; k += BG_THREADS)
    if (s_first[k])
    {
      int at = 0;
      for (int m = 0; m < ninc; ++m)
        at += s_first[m] && s_cand[m] < s_cand[k];
      s_ucell[at] = (int32_t)s_cand[k];
    }
  __syncthreads();
  // ---- contributions of every block: one thread per row, cells in ascending entity order (fixed summation order)
  if (tid < r1 - r0)
  {
    const int32_t r = r0 + tid;
    const int rowb = (int)(g.row_ptr[r] - b0);
    for (int64_t t = g.tptr[r]; t < g.tptr[r + 1]; ++t)
    {
      const uint32_t en = g.tent[t];
      const int64_t e = en >> 3;
      const int i = (int)(en & 7u);
      int lo = 0, hi = nuc - 1;
      while (lo < hi)
      {
        const int mid = (lo + hi) >> 1;
        if (s_ucell[mid] < (int32_t)e)
          lo = mid + 1;
        else
          hi = mid;
      }
      const char* prow = g.pos + e * g.pos_stride;
      for (int j = 0; j < 8; ++j)
      {
        const int p = g.pos_bytes == 1 ? (int)reinterpret_cast<const uint8_t*>(prow)[i * 8 + j]
                                       : (int)reinterpret_cast<const uint16_t*>(prow)[i * 8 + j];
        const int bidx = rowb + p;
        const int k = s_cnt[bidx];
        if (k >= BG_KMAX)
        {
          s_bad = 1;
          continue;
        }
        s_ent[bidx][k] = (uint16_t)((lo << 6) | (i << 3) | j);
        s_cnt[bidx] = (uint8_t)(k + 1);
      }
    }
    for (int b = rowb; b < (int)(g.row_ptr[r + 1] - b0); ++b)
      s_row[b] = (uint8_t)tid;
  }
  __syncthreads();
  if (s_bad)
  {
    if (tid == 0)
      *g.err = 1;
    return;
  }
  // ---- rank the blocks by list length (descending), stable in the block index
  for (int b = tid; b < nb; b += BG_THREADS)
    atomicAdd(&s_hist[s_cnt[b]], 1);
  __syncthreads();
  if (tid == 0)
  {
    int run = 0;
    for (int c = BG_KMAX; c >= 0; --c)
    {
      s_start[c] = run;
      run += s_hist[c];
    }
  }
  __syncthreads();
  {
    // blocks tid*4 .. tid*4+3; for every length c a block-wide exclusive scan of the flags (cnt == c)
    const int lane = tid & 31, wib = tid >> 5;
    for (int c = BG_KMAX; c >= 0; --c)
    {
      if (s_hist[c] == 0)
        continue; // (block-uniform)
      int mine = 0;
      for (int u = 0; u < BG_NBCAP / BG_THREADS; ++u)
      {
        const int b = tid * (BG_NBCAP / BG_THREADS) + u;
        mine += b < nb && s_cnt[b] == c;
      }
      int incl = mine;
      for (int o = 1; o < 32; o <<= 1)
      {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
          incl += v;
      }
      if (lane == 31)
        s_wsum[wib] = incl;
      __syncthreads();
      int base = s_start[c] + incl - mine;
      for (int w = 0; w < wib; ++w)
        base += s_wsum[w];
      for (int u = 0; u < BG_NBCAP / BG_THREADS; ++u)
      {
        const int b = tid * (BG_NBCAP / BG_THREADS) + u;
        if (b < nb && s_cnt[b] == c)
          s_rank[b] = (uint16_t)base++;
      }
      __syncthreads();
    }
  }
  // ---- blocks -> threads: every thread of the kernel owns whole blocks (accumulators in registers) and should run the
  // same number of contribution steps - on a structured mesh a tile has 16 x 8 x 8 = 1024 steps in blocks of 8, 4, 2
  // and 1, i.e. exactly 8 per thread.  Longest-processing-time rule over the ranked blocks: the next block goes to the
  // thread with the fewest steps so far (ties: lowest thread, which keeps similar lists in one warp); one warp decides.
  for (int b = tid; b < nb; b += BG_THREADS)
    s_byrank[s_rank[b]] = (uint16_t)b;
  s_sum[tid] = 0;
  s_nblk[tid] = 0;
  if (tid == 0)
  {
    s_S = 0;
    s_Q = 0;
  }
  for (int k = tid; k < BG_SCAP * BG_THREADS; k += BG_THREADS)
    s_steps[k] = (uint16_t)BG_IDLE;
  for (int k = tid; k < BG_QCAP * BG_THREADS; k += BG_THREADS)
    s_blocks[k] = 0;
  __syncthreads();
  if (tid < 32)
  {
    for (int r = 0; r < nb; ++r)
    {
      const int b = s_byrank[r];
      const int c = max((int)s_cnt[b], 1); // (a block without contributions still takes one step: it is written)
      int best = 0x7fffffff;
#pragma unroll
      for (int u = 0; u < BG_THREADS / 32; ++u)
      {
        const int t = tid + 32 * u;
        best = min(best, s_sum[t] * BG_THREADS + t);
      }
#pragma unroll
      for (int o = 16; o; o >>= 1)
        best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (tid == 0)
      {
        const int t = best % BG_THREADS;
        s_owner[b] = (uint8_t)t;
        s_s0[b] = (uint8_t)min(s_sum[t], 255);
        s_q[b] = (uint8_t)min(s_nblk[t], 255);
        s_sum[t] += c;
        s_nblk[t] += 1;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  atomicMax(&s_S, s_sum[tid]);
  atomicMax(&s_Q, s_nblk[tid]);
  __syncthreads();
  const int S = s_S, Q = s_Q;
  if (S > BG_SCAP || Q > BG_QCAP)
  {
    if (tid == 0)
      *g.err = 1;
    return;
  }
  for (int b = tid; b < nb; b += BG_THREADS)
  {
    const int t = s_owner[b], cnt = s_cnt[b], s0 = s_s0[b];
    s_blocks[(int)s_q[b] * BG_THREADS + t] = (uint16_t)(b | ((int)s_row[b] << 10));
    for (int k = 0; k < cnt; ++k) // cells in ascending entity order: the summation order of a block is fixed
      s_steps[(s0 + k) * BG_THREADS + t] = (uint16_t)(s_ent[b][k] | (k == cnt - 1 ? BG_LAST : 0u));
    if (cnt == 0)
      s_steps[s0 * BG_THREADS + t] = (uint16_t)(BG_LAST | BG_EMPTY);
  }
  __syncthreads();
  // ---- output (fixed strides per tile: every address of the kernel's prologue follows from the tile index)
  if (tid == 0)
  {
    BGTile h;
    h.b0 = b0;
    h.nb = nb, h.ninc = ninc, h.nuc = nuc, h.nsteps = S, h.nq = Q, h.pad2 = 0;
    for (int k = 0; k < BG_KMAX; ++k)
    {
      int nk = 0; // blocks with more than k contributions
      for (int c = k + 1; c <= BG_KMAX; ++c)
        nk += s_hist[c];
      h.nlev[k] = (uint16_t)nk;
    }
    g.tiles[tile] = h;
    atomicMax(g.maxima, nuc);
    atomicMax(g.maxima + 1, nb);
    atomicMax(g.maxima + 2, S);
    atomicMax(g.maxima + 3, Q);
  }
  for (int k = tid; k < nuc; k += BG_THREADS)
    g.cells[tile * BG_CCAP + k] = s_ucell[k];
  for (int k = tid; k < S * BG_THREADS; k += BG_THREADS)
    g.ent[tile * (BG_SCAP * BG_THREADS) + k] = s_steps[k];
  for (int k = tid; k < Q * BG_THREADS; k += BG_THREADS)
    g.perm[tile * (BG_QCAP * BG_THREADS) + k] = s_blocks[k];
}

// int8 markers of a block-size-3 space -> one byte per node (bit k = component k)
__global__ void k_bg_node_masks(int64_t n_nodes, const int8_t* __restrict__ bc, uint8_t* __restrict__ mask)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = (uint8_t)((bc[3 * i] ? 1 : 0) | (bc[3 * i + 1] ? 2 : 0) | (bc[3 * i + 2] ? 4 : 0));
}

struct BGArgs
{
  int64_t ntiles;     // all tiles of the plan
  int64_t tile_begin; // this launch: tiles [tile_begin, tile_end)
  int64_t tile_end;
  const BGTile* tiles;
  const int32_t* cells;
  const uint16_t *perm, *ent;
  const double* rec;
  const uint8_t* mask0; // per node: bit k set = component k of the ROW carries a Dirichlet condition (or NULL)
  const uint8_t* zcb;   // per (tile, block in tile): the same for the block's COLUMN, precomputed per call (or NULL)
  // round 2: the column mask of a block looked up by the thread that owns the block - its column index and the node's
  // mask byte are requested before the contribution loop and used after it - instead of a per-call pass over all
  // blocks (k_bg_block_masks: 0.67 ms of a 6.44 ms launch at C4)
  const int32_t* cols;  // CSR column (node) of every block
  const uint8_t* mask1; // per node: bit k set = component k of the COLUMN carries a Dirichlet condition (or NULL)
  int32_t n_mask_nodes;
  // default since: the same masks per plan cell (written by k_rg_records), staged with the cell records; a block reads its
  // column mask through the (cell slot, j) of its first contribution (or NULL)
  const uint32_t* cmask;
  double mu, lmbda;
  double* values;
  int dbg; // profiling only: 1 = no contribution loop, 2 = no store
  int overwrite;
  int img_cap, cell_cap, step_cap, q_cap, nb_cap; // shared-memory layout: image doubles; per stage: cells, steps and
                                                  // blocks per thread, blocks
};

__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}

// per (tile, block): Dirichlet mask of the block's column node (one pass over cols per call, coalesced)
__global__ void k_bg_block_masks(int64_t ntiles, const BGTile* __restrict__ tiles, const int32_t* __restrict__ cols,
                                 const uint8_t* __restrict__ mask1, int32_t n_nodes, uint8_t* __restrict__ zcb)
{
  const int64_t total = ntiles * BG_NBCAP;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t tile = t / BG_NBCAP;
    const int b = (int)(t - tile * BG_NBCAP);
    const BGTile& h = tiles[tile];
    if (b < h.nb)
    {
      const int32_t c = cols[h.b0 + b];
      zcb[t] = c < n_nodes ? mask1[c] : (uint8_t)0; // (columns beyond the rows' map: ghost columns no cell touches)
    }
  }
}

// Persistent CTAs, two shared-memory stages: while tile T is computed, the lists of the CTA's next tile N arrive by
// cp.async (their addresses follow from the tile index: fixed strides), the header of the tile after that is on its
// way, and the cell records of N (second level: they need N's cell ids) are requested after T's contribution loop.
// One TMA bulk store per tile; two block-wide barriers per tile.
// CM: where the column Dirichlet masks come from - 0: none / per-block bytes of a per-call pass (g.zcb), 1: column index
// -> node mask in the kernel (g.cols, g.mask1), 2: per-cell masks staged with the records (g.cmask; default)
// Every thread walks its own list of contribution steps (k_bg_plan balances them: 8 per thread on a structured mesh);
// a step flagged BG_LAST finishes its block: mu (tr D I + D^T) + lambda D, Dirichlet masks, 9 values into the image.
// (The first version ranked the blocks by list length and gave thread t the blocks t, t + 128, ...: warp 0 ran 13
// steps per tile and warps 2, 3 seven - 2.2 warps per issue slot parked at the tile barrier, r02_prof_q1_..._v3.csv.)
template <int CM>
__global__ void __launch_bounds__(BG_THREADS, 4) k_q1_blockgather(const BGArgs g)
{
  constexpr int NT = BG_THREADS;
  const bool use_cmask = CM == 2, use_mask1 = CM == 1;
  extern __shared__ __align__(16) unsigned char bg_raw[];
  double* img_base = reinterpret_cast<double*>(bg_raw);
  BGTile* s_hdr = reinterpret_cast<BGTile*>(bg_raw + sizeof(double) * g.img_cap); // ring of 3
  unsigned char* stage0 = reinterpret_cast<unsigned char*>(s_hdr + 3);
  const size_t rec_bytes = (size_t)g.cell_cap * RG_STRIDE * 8, ent_bytes = (size_t)g.step_cap * NT * 2,
               perm_bytes = (size_t)g.q_cap * NT * 2, zc_bytes = ((size_t)g.nb_cap + 15) & ~(size_t)15;
  // column node of every block of the tile (mask1 scheme) or column masks of every cell of the tile (cmask scheme)
  const size_t col_bytes = ((size_t)(use_cmask ? g.cell_cap : (use_mask1 ? g.nb_cap : 0)) * 4 + 15) & ~(size_t)15;
  const size_t stage_bytes = rec_bytes + ent_bytes + perm_bytes + zc_bytes + 16 + col_bytes;
  const int tid = threadIdx.x;
  const int64_t G = gridDim.x;

  auto fetch_hdr = [&](int64_t tile, int slot)
  {
    if (tid < 4)
      cp_async16(reinterpret_cast<unsigned char*>(s_hdr + slot) + 16 * tid, reinterpret_cast<const unsigned char*>(g.tiles + tile) + 16 * tid);
  };
  // lists of a tile whose header is in shared memory -> stage st; returns this thread's cell id (second-level key)
  auto fetch_lists = [&](int64_t tile, const BGTile& h, int st) -> int32_t
  {
    unsigned char* base = stage0 + (size_t)st * stage_bytes;
    const uint4* e4 = reinterpret_cast<const uint4*>(g.ent + tile * (BG_SCAP * NT));
    for (int k = tid; k < h.nsteps * (NT / 8); k += NT)
      cp_async16(base + rec_bytes + 16 * (size_t)k, e4 + k);
    const uint4* p4 = reinterpret_cast<const uint4*>(g.perm + tile * (BG_QCAP * NT));
    for (int k = tid; k < h.nq * (NT / 8); k += NT)
      cp_async16(base + rec_bytes + ent_bytes + 16 * (size_t)k, p4 + k);
    if (CM == 0 && g.zcb)
    {
      const uint4* z4 = reinterpret_cast<const uint4*>(g.zcb + tile * BG_NBCAP);
      for (int k = tid; k < (h.nb + 15) / 16; k += NT)
        cp_async16(base + rec_bytes + ent_bytes + perm_bytes + 16 * (size_t)k, z4 + k);
    }
    if (g.mask0 && tid == 0)
      cp_async16(base + rec_bytes + ent_bytes + perm_bytes + zc_bytes, g.mask0 + tile * BG_ROWS);
    // (tried: a per-tile "no marked column node in [cmin, cmax]" test through a per-call prefix count, to skip the
    // lookups away from the boundary layer - its two loads sit on the critical path between tiles: 6.52 against 6.21 ms)
    if (use_mask1) // the tile's blocks are consecutive CSR blocks: their columns are one contiguous run (4-byte copies:
    {            // the run starts at an arbitrary block)
      const int32_t* cg = g.cols + h.b0;
      unsigned char* cs = base + rec_bytes + ent_bytes + perm_bytes + zc_bytes + 16;
      for (int k = tid; k < h.nb; k += NT)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(cs + 4 * (size_t)k)),
                     "l"(cg + k)
                     : "memory");
    }
    return tid < h.nuc ? __ldg(g.cells + tile * BG_CCAP + tid) : -1;
  };
  auto fetch_records = [&](int32_t cid, int st)
  {
    if (cid >= 0)
    {
      const double2* rp = reinterpret_cast<const double2*>(g.rec) + (int64_t)cid * 5;
      double2* sp = reinterpret_cast<double2*>(stage0 + (size_t)st * stage_bytes) + tid * 5;
#pragma unroll
      for (int w = 0; w < 5; ++w)
        cp_async16(sp + w, rp + w);
      if (use_cmask)
      {
        unsigned char* cs = stage0 + (size_t)st * stage_bytes + rec_bytes + ent_bytes + perm_bytes + zc_bytes + 16;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(cs + 4 * (size_t)tid)),
                     "l"(g.cmask + cid)
                     : "memory");
      }
    }
  };

  int64_t T = g.tile_begin + blockIdx.x;
  if (T >= g.tile_end)
    return;
  // ---- prologue: headers of the first two tiles, lists and records of the first
  fetch_hdr(T, 0);
  if (T + G < g.tile_end)
    fetch_hdr(T + G, 1);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  {
    const int32_t cid = fetch_lists(T, s_hdr[0], 0);
    fetch_records(cid, 0);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  for (int it = 0; T < g.tile_end; ++it, T += G)
  {
    const int st = it & 1;
    const BGTile& h = s_hdr[it % 3];
    const int64_t N = T + G, NN = T + 2 * G;
    // ---- next tile: lists (first level); the tile after it: header
    int32_t cid_n = -1;
    if (N < g.tile_end)
      cid_n = fetch_lists(N, s_hdr[(it + 1) % 3], st ^ 1);
    if (NN < g.tile_end)
      fetch_hdr(NN, (it + 2) % 3);

    const unsigned char* base = stage0 + (size_t)st * stage_bytes;
    const double* s_rec = reinterpret_cast<const double*>(base);
    const uint16_t* s_ent = reinterpret_cast<const uint16_t*>(base + rec_bytes);
    const uint16_t* s_perm = reinterpret_cast<const uint16_t*>(base + rec_bytes + ent_bytes);
    const uint8_t* s_zc = base + rec_bytes + ent_bytes + perm_bytes;
    const uint8_t* s_zr = s_zc + zc_bytes;
    const int32_t* s_col = reinterpret_cast<const int32_t*>(s_zr + 16);
    const uint32_t* s_cm = reinterpret_cast<const uint32_t*>(s_zr + 16);
    const int nb = h.nb;
    const int64_t b0 = h.b0;
    // smem element i of the image sits at the same offset modulo 16 bytes as global element b0 * 9 + i
    const int head = (int)((b0 * 9) & 1);
    double* img = img_base + head;
    const bool skip = nb == 0 || (!g.overwrite && h.ninc == 0); // (add mode: no cell of the plan touches the tile)
    if (!skip)
    {
      const int nsteps = h.nsteps;
      double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      int q = 0;
#pragma unroll 1
      for (int sidx = 0; sidx < nsteps; ++sidx)
      {
        const uint32_t en = s_ent[sidx * NT + tid];
        if (en == BG_IDLE)
          break;
        const int slot = (int)((en >> 6) & 127u), jn = (int)(en & 7u);
        if (!(en & BG_EMPTY) && g.dbg != 1)
        {
          const double2* sp = reinterpret_cast<const double2*>(s_rec + slot * RG_STRIDE);
          const double2 q0 = sp[0], q1 = sp[1], q2 = sp[2], q3 = sp[3], q4 = sp[4];
          if (q4.y >= 0.0)
          {
            const double K[3][3] = {{q0.x, q0.y, q1.x}, {q1.y, q2.x, q2.y}, {q3.x, q3.y, q4.x}};
            q1_affine_D_add((int)((en >> 3) & 7u), jn, K, q4.y, A);
          }
        }
        if (en & BG_LAST)
        {
          const uint32_t pmt = s_perm[q * NT + tid];
          ++q;
          const int bidx = (int)(pmt & 1023u), rowid = (int)(pmt >> 10);
          unsigned zc = 0u;
          if (use_cmask) // every contribution of a block names its column node as (cell, j): this one will do
            zc = (en & BG_EMPTY) ? 0u : ((s_cm[slot] >> (3 * jn)) & 7u);
          else if (use_mask1)
          {
            const int32_t cnode = s_col[bidx];
            zc = cnode < g.n_mask_nodes ? (unsigned)__ldg(g.mask1 + cnode) : 0u; // (ghost columns no cell touches)
          }
          else if (g.zcb)
            zc = s_zc[bidx];
          const unsigned zr = g.mask0 ? s_zr[rowid] : 0u;
          const double tr = A[0][0] + A[1][1] + A[2][2];
          double* o = img + bidx * 9;
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = 0; l < 3; ++l)
            {
              const double v = g.mu * ((k == l ? tr : 0.0) + A[l][k]) + g.lmbda * A[k][l];
              o[3 * k + l] = (((zr >> k) | (zc >> l)) & 1u) ? 0.0 : v;
            }
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = 0; l < 3; ++l)
              A[k][l] = 0.0;
        }
      }
    }
    // ---- next tile: cell records (second level; the ids were requested before the contribution loop)
    fetch_records(cid_n, st ^ 1);
    double* out = g.values + b0 * 9;
    const int total = nb * 9;
    if (g.overwrite)
    {
      // generic-proxy writes of the image -> visible to the async proxy, then one bulk store of the aligned part
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0 && !skip)
      {
        if (head)
          out[0] = img[0];
        const int nbulk = g.dbg == 2 ? 0 : (total - head) & ~1;
        if (nbulk > 0)
        {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + head),
                       "r"((uint32_t)__cvta_generic_to_shared(img + head)), "r"((uint32_t)nbulk * 8u)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (total - head - nbulk)
          out[total - 1] = img[total - 1];
      }
    }
    else
    {
      __syncthreads();
      if (!skip)
        for (int t0i = tid; t0i < total; t0i += 8 * NT)
        {
          double old[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
          {
            const int t = t0i + u * NT;
            old[u] = t < total ? out[t] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
          {
            const int t = t0i + u * NT;
            if (t < total)
              out[t] = old[u] + img[t];
          }
        }
    }
    // ---- the next tile's data has landed, the image may be overwritten
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (tid == 0)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
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
  cudaFree(g->bg_cells);
  cudaFree(g->bg_tiles);
  cudaFree(g->bg_perm);
  cudaFree(g->bg_ent);
  cudaFree(g->bg_zcb);
  cudaFree(g->bg_mask0);
  cudaFree(g->bg_mask1);
  cudaFree(g->bg_cmask);
  delete g;
}

int rowgather_tile_rows(const bfx_asm* P) { return (P && P->rowgather && P->rowgather->bg_ok) ? BG_ROWS : 0; }

// row_begin / row_end (block rows, multiples of BG_ROWS; row_end < 0: all rows): the rows this call writes.  Every row
// is formed completely from the plan's cells, so calls on disjoint row ranges add up to the one-launch result: ghost
// rows first, their exchange behind the owned rows (fem.assemble_matrix_overlapped).  reuse_records: the cell records
// and the list of non-affine cells of the previous call on the same geometry are still valid.
int launch_rowgather_q1(const bfx_asm* P, const AsmArgs& a, int values_mode, cudaStream_t st, int32_t row_begin,
                        int32_t row_end, bool reuse_records)
{
  const bfx_rowgather* G = P->rowgather;
  if (!G)
    return fail(BFX_ERR_INVALID, "BFX_ASM_ROWGATHER needs bfx_asm_build_rowgather() on the plan first");
  if (a.dofmap1 != a.dofmap0)
    return fail(BFX_ERR_UNSUPPORTED, "row-gather assembly needs identical test and trial dofmaps");
  const bfx_csr* csr = P->csr;
  const bool ranged = row_end >= 0 && !(row_begin == 0 && row_end >= csr->n_rows_all);
  const bool use_bg = G->bg_ok && !getenv("BFX_ROWGATHER_V1");
  if (ranged && !use_bg)
    return fail(BFX_ERR_UNSUPPORTED, "row-range assembly needs the block-gather plan");
  // per-call Dirichlet masks per node (rows and columns share the node numbering: one dofmap)
  const uint8_t *m0 = nullptr, *m1 = nullptr;
  if (use_bg && a.bc0)
  {
    k_bg_node_masks<<<grid_for(G->bg_n_nodes, 256, 8), 256, 0, st>>>(G->bg_n_nodes, a.bc0, G->bg_mask0);
    m0 = G->bg_mask0;
  }
  if (use_bg && a.bc1)
  {
    m1 = G->bg_mask0;
    if (a.bc1 != a.bc0)
    {
      k_bg_node_masks<<<grid_for(G->bg_n_nodes, 256, 8), 256, 0, st>>>(G->bg_n_nodes, a.bc1, G->bg_mask1);
      m1 = G->bg_mask1;
    }
  }
  // where the kernel takes a block's column mask from: 2 = per-cell masks written by the records pre-pass (default),
  // 1 = column index -> node mask in the kernel, 0 = one pass over all blocks per call (both kept for A/B runs)
  const int col_scheme = !m1 ? -1 : ((getenv("BFX_BG_BLOCK_MASKS") && G->bg_zcb) ? 0 : (getenv("BFX_BG_COL_LOOKUP") ? 1 : 2));
  // pre-pass: cell records (+ column masks per cell) + list of cells that are not parallelepipeds.  reuse_records: the
  // records AND the cell masks of the previous call are still valid (same geometry, same column markers)
  if (!reuse_records)
  {
    BFX_CUDA(cudaMemsetAsync(G->na_count, 0, sizeof(unsigned long long), st));
    if (P->ncells > 0)
      k_rg_records<<<grid_for(P->ncells, 128, 0), 128, 0, st>>>(P->ncells, P->cells, P->x_dofmap, a.x, G->rec, G->na_cells,
                                                                 G->na_count, a.dofmap1, m1,
                                                                 col_scheme == 2 ? G->bg_cmask : nullptr);
  }
  if (use_bg) // block-gather kernel (default); the round-1 kernel stays for A/B runs
  {
    BGArgs b;
    b.ntiles = G->bg_ntiles;
    b.tile_begin = ranged ? row_begin / BG_ROWS : 0;
    b.tile_end = ranged ? std::min<int64_t>(((int64_t)row_end + BG_ROWS - 1) / BG_ROWS, G->bg_ntiles) : G->bg_ntiles;
    b.tiles = static_cast<const BGTile*>(G->bg_tiles);
    b.cells = G->bg_cells, b.perm = G->bg_perm, b.ent = G->bg_ent;
    b.rec = G->rec;
    b.mask0 = m0;
    b.zcb = nullptr;
    b.cols = csr->cols;
    b.mask1 = nullptr;
    b.cmask = nullptr;
    b.n_mask_nodes = (int32_t)G->bg_n_nodes;
    if (col_scheme == 0)
    {
      k_bg_block_masks<<<grid_for(G->bg_ntiles * BG_NBCAP, 256, 16), 256, 0, st>>>(G->bg_ntiles, static_cast<const BGTile*>(G->bg_tiles), csr->cols, m1,
                                                                                (int32_t)G->bg_n_nodes, G->bg_zcb);
      b.zcb = G->bg_zcb;
    }
    else if (col_scheme == 1)
      b.mask1 = m1;
    else if (col_scheme == 2)
      b.cmask = G->bg_cmask;
    b.mu = a.constants[0], b.lmbda = a.constants[1];
    b.values = a.values;
    b.overwrite = values_mode == BFX_VALUES_OVERWRITE;
    b.dbg = getenv("BFX_BG_DBG") ? atoi(getenv("BFX_BG_DBG")) : 0;
    b.img_cap = (G->bg_max_blocks * 9 + 3) & ~1;
    b.cell_cap = G->bg_max_cells;
    b.step_cap = G->bg_max_steps;
    b.q_cap = G->bg_max_q;
    b.nb_cap = G->bg_max_blocks;
    const size_t stage = (size_t)b.cell_cap * RG_STRIDE * 8 + (size_t)b.step_cap * BG_THREADS * 2 + (size_t)b.q_cap * BG_THREADS * 2
                         + (((size_t)b.nb_cap + 15) & ~(size_t)15) + 16
                         + (((size_t)(b.cmask ? b.cell_cap : (b.mask1 ? b.nb_cap : 0)) * 4 + 15) & ~(size_t)15);
    const size_t smem = sizeof(double) * (size_t)b.img_cap + 3 * sizeof(BGTile) + 2 * stage;
    if (b.tile_end > b.tile_begin)
    {
      void (*kern)(const BGArgs) = b.cmask ? k_q1_blockgather<2> : (b.mask1 ? k_q1_blockgather<1> : k_q1_blockgather<0>);
      const int threads = BG_THREADS;
      BFX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      BFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
      const int64_t grid = std::min<int64_t>(b.tile_end - b.tile_begin, (int64_t)sm_count() * std::max(per_sm, 1));
      kern<<<(unsigned)grid, threads, smem, st>>>(b);
    }
    BFX_CHECK_LAUNCH();
    AsmArgs nb = a;
    nb.cells = G->na_cells;
    nb.n = P->ncells;
    nb.n_dev = G->na_count;
    nb.pos = nullptr;
    if (ranged) // the non-affine cells add to this call's rows only
    {
      nb.row_lo = (int32_t)(b.tile_begin * BG_ROWS);
      nb.row_hi = (int32_t)std::min<int64_t>(b.tile_end * BG_ROWS, csr->n_rows_all);
    }
    return launch_q1_red(P, nb, st);
  }
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
  BFX_REQUIRE(P && P->csr, "bfx_asm_build_rowgather: plan has no matrix");
  if (P->ncells == 0)
    return fail(BFX_ERR_UNSUPPORTED, "row-gather plan of an empty cell list");
  BFX_REQUIRE(P->pos, "bfx_asm_build_rowgather: plan has no position map");
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
  // ---- block-gather plan (tiles of BG_ROWS rows); a tile over a capacity keeps the round-1 kernel
  if (n > 0 && n_rows > 0)
  {
    const int64_t ntiles = ((int64_t)n_rows + BG_ROWS - 1) / BG_ROWS;
    int *d_err = nullptr, *d_max = nullptr;
    if ((e = dev_alloc(&G->bg_cells, (size_t)ntiles * BG_CCAP)) || (e = dev_alloc(reinterpret_cast<BGTile**>(&G->bg_tiles), (size_t)ntiles))
        || (e = dev_alloc(&G->bg_perm, (size_t)ntiles * BG_QCAP * BG_THREADS))
        || (e = dev_alloc(&G->bg_ent, (size_t)ntiles * BG_SCAP * BG_THREADS)) || (e = dev_alloc(&d_err, 1)) || (e = dev_alloc(&d_max, 4)))
      return bail(e);
    if (getenv("BFX_BG_BLOCK_MASKS") && (e = dev_alloc(&G->bg_zcb, (size_t)ntiles * BG_NBCAP))) // (A/B scheme only)
      return bail(e);
    BFX_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    BFX_CUDA(cudaMemsetAsync(d_max, 0, 4 * sizeof(int), st));
    BGPlanArgs b;
    b.n_rows = n_rows;
    b.row_ptr = csr->row_ptr;
    b.tptr = G->tptr;
    b.tent = G->tent;
    b.pos = P->pos;
    b.pos_bytes = P->pos_bytes;
    b.pos_stride = P->pos_stride;
    b.cells = G->bg_cells, b.perm = G->bg_perm, b.ent = G->bg_ent, b.tiles = static_cast<BGTile*>(G->bg_tiles);
    b.err = d_err;
    b.maxima = d_max;
    k_bg_plan<<<(unsigned)ntiles, BG_THREADS, 0, st>>>(b);
    BFX_CHECK_LAUNCH();
    int h_err = 0, h_max[4] = {0, 0, 0, 0};
    BFX_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaMemcpyAsync(h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_err);
    cudaFree(d_max);
    // per-call scratch of the Dirichlet node masks (rows and columns share the node numbering: one dofmap)
    G->bg_n_nodes = n_rows;
    if ((e = dev_alloc(&G->bg_mask0, (size_t)n_rows + 2 * BG_ROWS)) || (e = dev_alloc(&G->bg_mask1, (size_t)n_rows + 2 * BG_ROWS))
        || (e = dev_alloc(&G->bg_cmask, (size_t)n + 1)))
      return bail(e);
    G->bg_ntiles = ntiles;
    G->bg_max_cells = h_max[0];
    G->bg_max_blocks = h_max[1];
    G->bg_max_steps = h_max[2];
    G->bg_max_q = h_max[3];
    G->bg_ok = h_err == 0;
  }
  P->rowgather = G;
  return BFX_OK;
}
}
