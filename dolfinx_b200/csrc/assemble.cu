// Cell / exterior-facet assembly on device: fem::impl::assemble_cells_matrix<LiftingMode>
// (fem/assemble_matrix_impl.h:92-200), assemble_entities (:264-379), assemble_cells (vector,
// fem/assemble_vector_impl.h:72-116), assemble_entities (:157-215), lift_bc (:361-414),
// pack_coefficient_entity (fem/pack.h:121-176) and DirichletBC::mark_dofs/set
// (fem/DirichletBC.h:495-601).
#include "assemble.cuh"
#include "elements.cuh"
#include "asm_device.cuh"
#include "cell_kernels.cuh"
#include <map>
#include <vector>

using namespace bfx;

namespace
{
// ---------------------------------------------------------------------------------------------
// Q1 hexahedron linear elasticity (bs = 3): 64 threads per cell, one per (i, j) node pair.
// Phase 1: 8 threads per cell evaluate K = J^{-1}, |det J| and the 8 physical gradients at the
// 2x2x2 Gauss points into shared memory.  Phase 2: thread (i, j) accumulates
// D[a][b] = sum_q w_q g_i[a] g_j[b] and forms the 3x3 block
//   A[(i,k),(j,l)] = mu (delta_kl tr D + D[l][k]) + lambda D[k][l]
// (python/demo/demo_elasticity.py:131-150), then 9 REDs into the bs=3 CSR block.
// ---------------------------------------------------------------------------------------------
constexpr int Q1_CELLS = 4;

template <typename PosT, int MODE>
__global__ void __launch_bounds__(64 * Q1_CELLS, 3) k_elasticity_q1(const AsmArgs a)
{
  __shared__ double s_xc[Q1_CELLS][8][3];
  __shared__ double s_g[Q1_CELLS][8][8][3];
  __shared__ double s_w[Q1_CELLS][8];
  __shared__ int32_t s_d0[Q1_CELLS][8], s_d1[Q1_CELLS][8];
  __shared__ int s_skip[Q1_CELLS];
  __shared__ int s_affine[Q1_CELLS];
  __shared__ double s_K[Q1_CELLS][10]; // affine cells: K = J^{-1} (9) and |det J|
  // staging of the 64 3x3 blocks of a cell: the 576 scalars are then issued as REDs in address
  // order (9 consecutive doubles per block), so that the lanes of one RED instruction share 32 B
  // sectors (each fp64 RED costs a full sector on the L1->L2 crossbar)
  __shared__ double s_blk[MODE == 0 ? Q1_CELLS : 1][64 * 9];
  __shared__ long long s_p[MODE == 0 ? Q1_CELLS : 1][64];

  const int cl = threadIdx.x >> 6, t = threadIdx.x & 63;
  const double mu = a.constants[0], lmbda = a.constants[1];
  const int64_t n_ent = a.n_dev ? min(a.n, (int64_t)*a.n_dev) : a.n;
  for (int64_t e0 = (int64_t)blockIdx.x * Q1_CELLS; e0 < n_ent; e0 += (int64_t)gridDim.x * Q1_CELLS)
  {
    const int64_t e = e0 + cl;
    const bool active = e < n_ent;
    int32_t cell = 0;
    if (active)
      cell = a.cells ? a.cells[e] : (int32_t)e;
    if (active && t < 24)
    {
      const int node = t / 3, m = t % 3;
      s_xc[cl][node][m] = a.x[3 * (int64_t)a.x_dofmap[(int64_t)cell * 8 + node] + m];
    }
    if (active && t >= 32 && t < 40)
      s_d0[cl][t - 32] = a.dofmap0[(int64_t)cell * 8 + (t - 32)];
    if (active && t >= 40 && t < 48)
      s_d1[cl][t - 40] = a.dofmap1[(int64_t)cell * 8 + (t - 40)];
    __syncthreads();
    if (t == 0)
    {
      int skip = !active;
      if (MODE == 1 && active)
      {
        int any = 0;
        for (int j = 0; j < 8; ++j)
          for (int k = 0; k < 3; ++k)
            any |= a.bc1[3 * (int64_t)s_d1[cl][j] + k];
        skip = !any;
      }
      s_skip[cl] = skip;
    }
    if (active && t < 8)
    {
      double xc[8][3];
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int m = 0; m < 3; ++m)
          xc[n][m] = s_xc[cl][n][m];
      // Parallelepiped test: x3 = x1+x2-x0, x5 = x1+x4-x0, x6 = x2+x4-x0, x7 = x1+x2+x4-2x0.
      // On such (affine) cells the integrals are pre-integrated exactly from the 1-D reference
      // integrals; general trilinear cells use 2x2x2 Gauss points.
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
      if (affine)
      {
        if (t == 0)
        {
          const double X[3] = {0.5, 0.5, 0.5};
          double K[3][3];
          const double det = el::HexQ1::jacobian_inverse(xc, X, K);
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int m = 0; m < 3; ++m)
              s_K[cl][3 * c + m] = K[c][m];
          s_K[cl][9] = fabs(det);
          s_affine[cl] = 1;
        }
      }
      else
      {
        if (t == 0)
          s_affine[cl] = 0;
        const double g0 = 0.5 - 0.28867513459481287, g1 = 0.5 + 0.28867513459481287;
        const double X[3] = {(t & 1) ? g1 : g0, (t & 2) ? g1 : g0, (t & 4) ? g1 : g0};
        double K[3][3];
        const double det = el::HexQ1::jacobian_inverse(xc, X, K);
        s_w[cl][t] = 0.125 * fabs(det);
#pragma unroll
        for (int n = 0; n < 8; ++n)
        {
          double d[3], phi;
          el::HexQ1::dphi(n, X, d, phi);
#pragma unroll
          for (int m = 0; m < 3; ++m)
            s_g[cl][t][n][m] = d[0] * K[0][m] + d[1] * K[1][m] + d[2] * K[2][m];
        }
      }
    }
    __syncthreads();
    const bool work = !s_skip[cl];
    if (work)
    {
      const int i = t >> 3, j = t & 7;
      double D[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      if (s_affine[cl])
      {
        // reference tensor That[c][d] = int d_c phi_i d_d phi_j over the unit cube, from the 1-D
        // integrals of N0 = 1 - s, N1 = s: mass 1/3 | 1/6, stiffness +1 | -1, mixed +-1/2
        double Mm[3], Ss[3], Cij[3], Cji[3];
#pragma unroll
        for (int m = 0; m < 3; ++m)
        {
          const int bi = (i >> m) & 1, bj = (j >> m) & 1;
          Mm[m] = bi == bj ? (1.0 / 3.0) : (1.0 / 6.0);
          Ss[m] = bi == bj ? 1.0 : -1.0;
          Cij[m] = bi ? 0.5 : -0.5; // int N'_i N_j
          Cji[m] = bj ? 0.5 : -0.5; // int N_i N'_j
        }
        double That[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int d = 0; d < 3; ++d)
          {
            const int o = 3 - c - d; // the third direction when c != d
            That[c][d] = c == d ? Ss[c] * Mm[(c + 1) % 3] * Mm[(c + 2) % 3] : Cij[c] * Cji[d] * Mm[o];
          }
        double K[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int m = 0; m < 3; ++m)
            K[c][m] = s_K[cl][3 * c + m];
        const double adet = s_K[cl][9];
        double M1[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int b = 0; b < 3; ++b)
            M1[c][b] = That[c][0] * K[0][b] + That[c][1] * K[1][b] + That[c][2] * K[2][b];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int b = 0; b < 3; ++b)
            D[p][b] = adet * (K[0][p] * M1[0][b] + K[1][p] * M1[1][b] + K[2][p] * M1[2][b]);
      }
      else
      {
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
          const double wq = s_w[cl][q];
          const double gi[3] = {s_g[cl][q][i][0], s_g[cl][q][i][1], s_g[cl][q][i][2]};
          const double gj[3] = {s_g[cl][q][j][0], s_g[cl][q][j][1], s_g[cl][q][j][2]};
#pragma unroll
          for (int p = 0; p < 3; ++p)
          {
            const double wg = wq * gi[p];
#pragma unroll
            for (int r = 0; r < 3; ++r)
              D[p][r] = fma(wg, gj[r], D[p][r]);
          }
        }
      }
      const double tr = D[0][0] + D[1][1] + D[2][2];
      const int32_t r = s_d0[cl][i], c = s_d1[cl][j];
      if constexpr (MODE == 0)
      {
        int64_t p;
        const int64_t rb = a.row_ptr[r];
        if (a.pos)
          p = rb + static_cast<const PosT*>(a.pos)[e * 64 + t];
        else
          p = find_col(a.cols, rb, a.row_ptr[r + 1], c);
        if (p < 0)
          *a.err = 1;
        const bool outside = a.row_hi > 0 && (r < a.row_lo || r >= a.row_hi); // (row-range call: another call's rows)
        s_p[cl][t] = (p < 0 || outside) ? -1 : p * 9;
        bool zr[3] = {false, false, false}, zc[3] = {false, false, false};
        if (a.bc0)
        {
#pragma unroll
          for (int k = 0; k < 3; ++k)
            zr[k] = a.bc0[3 * (int64_t)r + k];
        }
        if (a.bc1)
        {
#pragma unroll
          for (int l = 0; l < 3; ++l)
            zc[l] = a.bc1[3 * (int64_t)c + l];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int l = 0; l < 3; ++l)
          {
            const double v = mu * ((k == l ? tr : 0.0) + D[l][k]) + lmbda * D[k][l];
            s_blk[cl][t * 9 + k * 3 + l] = (zr[k] || zc[l]) ? 0.0 : v;
          }
      }
      else
      {
        double dv[3];
#pragma unroll
        for (int l = 0; l < 3; ++l)
        {
          const int64_t jj = 3 * (int64_t)c + l;
          dv[l] = a.bc1[jj] ? a.alpha * (a.bc_values1[jj] - (a.x0 ? a.x0[jj] : 0.0)) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
          double acc = 0.0;
#pragma unroll
          for (int l = 0; l < 3; ++l)
            acc = fma(mu * ((k == l ? tr : 0.0) + D[l][k]) + lmbda * D[k][l], dv[l], acc);
          if (acc != 0.0)
            red_add(a.b + 3 * (int64_t)r + k, -acc);
        }
      }
    }
    if constexpr (MODE == 0)
    {
      __syncthreads();
      if (work)
      {
#pragma unroll
        for (int it = 0; it < 9; ++it)
        {
          const int s = it * 64 + t;
          const int b = s / 9;
          const long long p = s_p[cl][b];
          const double v = s_blk[cl][s];
          if (p >= 0 && v != 0.0)
            red_add(a.values + p + (s - 9 * b), v);
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Q1 elasticity, second generation (assembly mode only).  8 cells per CTA, 36 threads per cell: one
// thread per UNIQUE node pair i <= j (the form is symmetric, block(j,i) = block(i,j)^T), so the
// fp64 work per cell drops from 64 to 36 block evaluations.  Parallelepiped cells use the
// pre-integrated reference tensor (no quadrature loop, no per-point gradients); general trilinear
// cells fall back to 2x2x2 Gauss points staged in shared memory.  All 64 blocks of a cell are
// staged in shared memory with the bc rows/columns zeroed, then the CTA issues the 8 x 576 scalars
// as REDs in address order (sector-sharing lanes, see k_elasticity_q1).
// ---------------------------------------------------------------------------------------------
constexpr int Q2_CELLS = 8;
constexpr int Q2_THREADS = 36 * Q2_CELLS;

struct Q2Smem
{
  double blk[Q2_CELLS][64 * 9];
  double g[Q2_CELLS][8][8][3];
  double w[Q2_CELLS][8];
  double xc[Q2_CELLS][8][3];
  double K[Q2_CELLS][10];
  long long p[Q2_CELLS][64];
  int32_t d0[Q2_CELLS][8], d1[Q2_CELLS][8];
  int affine[Q2_CELLS];
  unsigned char z0[Q2_CELLS][24], z1[Q2_CELLS][24];
  unsigned char pi[36], pj[36];
};

template <typename PosT>
__global__ void __launch_bounds__(Q2_THREADS, 2) k_elasticity_q1_sym(const AsmArgs a)
{
  extern __shared__ __align__(16) unsigned char q2_raw[];
  Q2Smem& S = *reinterpret_cast<Q2Smem*>(q2_raw);
  const int cl = threadIdx.x / 36, u = threadIdx.x % 36;
  const double mu = a.constants[0], lmbda = a.constants[1];
  if (threadIdx.x < 36)
  {
    // u -> (i, j), i <= j, row-major over the upper triangle
    int i = 0, rem = threadIdx.x;
    while (rem >= 8 - i)
    {
      rem -= 8 - i;
      ++i;
    }
    S.pi[threadIdx.x] = (unsigned char)i;
    S.pj[threadIdx.x] = (unsigned char)(i + rem);
  }
  const int64_t n_ent = a.n_dev ? min(a.n, (int64_t)*a.n_dev) : a.n;
  for (int64_t e0 = (int64_t)blockIdx.x * Q2_CELLS; e0 < n_ent; e0 += (int64_t)gridDim.x * Q2_CELLS)
  {
    const int64_t e = e0 + cl;
    const bool active = e < n_ent;
    int32_t cell = 0;
    if (active)
      cell = a.cells ? a.cells[e] : (int32_t)e;
    if (active && u < 24)
    {
      const int node = u / 3, m = u % 3;
      S.xc[cl][node][m] = a.x[3 * (int64_t)a.x_dofmap[(int64_t)cell * 8 + node] + m];
      const int32_t r = a.dofmap0[(int64_t)cell * 8 + node], c = a.dofmap1[(int64_t)cell * 8 + node];
      S.z0[cl][u] = a.bc0 ? (unsigned char)a.bc0[3 * (int64_t)r + m] : 0;
      S.z1[cl][u] = a.bc1 ? (unsigned char)a.bc1[3 * (int64_t)c + m] : 0;
      if (m == 0)
      {
        S.d0[cl][node] = r;
        S.d1[cl][node] = c;
      }
    }
    __syncthreads();
    if (active && u < 8)
    {
      double xc[8][3];
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int m = 0; m < 3; ++m)
          xc[n][m] = S.xc[cl][n][m];
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
      if (affine)
      {
        if (u == 0)
        {
          const double X[3] = {0.5, 0.5, 0.5};
          double K[3][3];
          const double det = el::HexQ1::jacobian_inverse(xc, X, K);
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int m = 0; m < 3; ++m)
              S.K[cl][3 * c + m] = K[c][m];
          S.K[cl][9] = fabs(det);
          S.affine[cl] = 1;
        }
      }
      else
      {
        if (u == 0)
          S.affine[cl] = 0;
        const double g0 = 0.5 - 0.28867513459481287, g1 = 0.5 + 0.28867513459481287;
        const double X[3] = {(u & 1) ? g1 : g0, (u & 2) ? g1 : g0, (u & 4) ? g1 : g0};
        double K[3][3];
        const double det = el::HexQ1::jacobian_inverse(xc, X, K);
        S.w[cl][u] = 0.125 * fabs(det);
#pragma unroll
        for (int n = 0; n < 8; ++n)
        {
          double d[3], phi;
          el::HexQ1::dphi(n, X, d, phi);
#pragma unroll
          for (int m = 0; m < 3; ++m)
            S.g[cl][u][n][m] = d[0] * K[0][m] + d[1] * K[1][m] + d[2] * K[2][m];
        }
      }
    }
    __syncthreads();
    if (active)
    {
      const int i = S.pi[u], j = S.pj[u];
      double D[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      if (S.affine[cl])
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
        double K[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int m = 0; m < 3; ++m)
            K[c][m] = S.K[cl][3 * c + m];
        const double adet = S.K[cl][9];
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
      else
      {
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
          const double wq = S.w[cl][q];
          const double gi[3] = {S.g[cl][q][i][0], S.g[cl][q][i][1], S.g[cl][q][i][2]};
          const double gj[3] = {S.g[cl][q][j][0], S.g[cl][q][j][1], S.g[cl][q][j][2]};
#pragma unroll
          for (int p = 0; p < 3; ++p)
          {
            const double wg = wq * gi[p];
#pragma unroll
            for (int r = 0; r < 3; ++r)
              D[p][r] = fma(wg, gj[r], D[p][r]);
          }
        }
      }
      const double tr = D[0][0] + D[1][1] + D[2][2];
      // positions of block (i,j) and of its transpose (j,i)
      const int32_t ri = S.d0[cl][i], rj = S.d0[cl][j], ci = S.d1[cl][i], cj = S.d1[cl][j];
      int64_t pij, pji;
      if (a.pos)
      {
        const PosT* pp = static_cast<const PosT*>(a.pos) + e * 64;
        pij = a.row_ptr[ri] + pp[i * 8 + j];
        pji = a.row_ptr[rj] + pp[j * 8 + i];
      }
      else
      {
        pij = find_col(a.cols, a.row_ptr[ri], a.row_ptr[ri + 1], cj);
        pji = find_col(a.cols, a.row_ptr[rj], a.row_ptr[rj + 1], ci);
      }
      if (pij < 0 || pji < 0)
        *a.err = 1;
      // (row-range call: blocks in rows of another call's range are dropped)
      const bool out_i = a.row_hi > 0 && (ri < a.row_lo || ri >= a.row_hi);
      const bool out_j = a.row_hi > 0 && (rj < a.row_lo || rj >= a.row_hi);
      S.p[cl][i * 8 + j] = (pij < 0 || out_i) ? -1 : pij * 9;
      S.p[cl][j * 8 + i] = (pji < 0 || out_j) ? -1 : pji * 9;
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = 0; l < 3; ++l)
        {
          // A[(i,k),(j,l)] = mu (delta_kl tr D + D[l][k]) + lambda D[k][l]  =  A[(j,l),(i,k)]
          const double v = mu * ((k == l ? tr : 0.0) + D[l][k]) + lmbda * D[k][l];
          S.blk[cl][(i * 8 + j) * 9 + k * 3 + l] = (S.z0[cl][3 * i + k] || S.z1[cl][3 * j + l]) ? 0.0 : v;
          if (i != j)
            S.blk[cl][(j * 8 + i) * 9 + l * 3 + k] = (S.z0[cl][3 * j + l] || S.z1[cl][3 * i + k]) ? 0.0 : v;
        }
    }
    __syncthreads();
    const int ncell = (int)min((int64_t)Q2_CELLS, n_ent - e0);
#pragma unroll 4
    for (int it = 0; it < (Q2_CELLS * 576) / Q2_THREADS; ++it)
    {
      const int s = it * Q2_THREADS + threadIdx.x;
      const int c2 = s / 576;
      if (c2 >= ncell)
        break;
      const int loc = s - c2 * 576;
      const int b = loc / 9;
      const long long p = S.p[c2][b];
      const double v = S.blk[c2][loc];
      if (p >= 0 && v != 0.0)
        red_add(a.values + p + (loc - 9 * b), v);
    }
    __syncthreads();
  }
}

// has_bc (fem/assemble_matrix_impl.h:27-34) over the whole cell list: the cells with a marked column dof, compacted.
// Lifting touches only those (a boundary layer); the per-cell kernels then run on the short list.
__global__ void k_cells_with_bc(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ dofmap1,
                                int nd1, int bs1, const int8_t* __restrict__ bc1, int32_t* __restrict__ out,
                                unsigned long long* __restrict__ count)
{
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
  {
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    int any = 0;
    for (int j = 0; j < nd1; ++j)
    {
      const int64_t d = (int64_t)bs1 * dofmap1[(int64_t)cell * nd1 + j];
      for (int k = 0; k < bs1; ++k)
        any |= bc1[d + k];
    }
    if (any)
      out[atomicAdd(count, 1ULL)] = cell;
  }
}

// ---------------------------------------------------------------------------------------------
// plan construction: cell -> CSR position map (replaces std::lower_bound of insert_csr)
// ---------------------------------------------------------------------------------------------
__global__ void k_max_row_len(int32_t n, const int64_t* __restrict__ row_ptr, int* __restrict__ out)
{
  int m = 0;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    m = max(m, (int)(row_ptr[i + 1] - row_ptr[i]));
  for (int o = 16; o > 0; o >>= 1)
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0)
    atomicMax(out, m);
}

template <typename PosT>
__global__ void k_build_pos(int64_t ncells, const int32_t* __restrict__ cells, const int32_t* __restrict__ dofmap0,
                            int nd0, const int32_t* __restrict__ dofmap1, int nd1,
                            const int64_t* __restrict__ row_ptr, const int32_t* __restrict__ cols, int32_t n_rows,
                            char* __restrict__ pos, int stride, int* __restrict__ err)
{
  const int64_t total = ncells * nd0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t e = t / nd0;
    const int i = (int)(t - e * nd0);
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    const int32_t r = dofmap0[(int64_t)cell * nd0 + i];
    PosT* out = reinterpret_cast<PosT*>(pos + e * stride) + (int64_t)i * nd1;
    if (r < 0 || r >= n_rows)
    {
      *err = 2;
      continue;
    }
    const int64_t rb = row_ptr[r], re = row_ptr[r + 1];
    for (int j = 0; j < nd1; ++j)
    {
      const int64_t p = find_col(cols, rb, re, dofmap1[(int64_t)cell * nd1 + j]);
      if (p < 0)
      {
        *err = 1;
        out[j] = 0;
      }
      else
        out[j] = (PosT)(p - rb);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pack / bc kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_pack(double* __restrict__ coeffs, int cstride, int offset, const double* __restrict__ v,
                       const int32_t* __restrict__ dofmap, int nd, int bs, const int32_t* __restrict__ cells,
                       const int32_t* __restrict__ entities, int64_t n)
{
  const int64_t total = n * nd * bs;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t e = t / (nd * bs);
    const int rem = (int)(t - e * nd * bs);
    const int i = rem / bs, k = rem - i * bs;
    const int32_t cell = entities ? entities[2 * e] : (cells ? cells[e] : (int32_t)e);
    if (cell < 0)
      continue;
    coeffs[e * cstride + offset + bs * i + k] = v[(int64_t)bs * dofmap[(int64_t)cell * nd + i] + k];
  }
}

__global__ void k_bc_mark(int8_t* __restrict__ markers, const int32_t* __restrict__ dofs, int64_t n)
{
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    markers[dofs[t]] = 1;
}

__global__ void k_bc_set(double* __restrict__ x, int32_t x_size, const int32_t* __restrict__ dofs0,
                         const int32_t* __restrict__ dofs_g, int64_t n, const double* __restrict__ g, int g_kind,
                         int bs, const double* __restrict__ x0, double alpha)
{
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int32_t dof = dofs0[t];
    if (dof >= x_size)
      continue; // DirichletBC.h:506-511
    double v = 0.0;
    if (alpha != 0.0)
    {
      const double gv = g_kind == 0 ? g[dofs_g ? dofs_g[t] : dof] : g[dof % bs];
      v = x0 ? alpha * (gv - x0[dof]) : alpha * gv;
    }
    x[dof] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------
template <class E, int MODE>
int launch_matrix(const bfx_asm* P, const AsmArgs& a, cudaStream_t st)
{
  if (a.n == 0)
    return BFX_OK;
  const unsigned grid = grid_for(a.n, 128, 0);
  if (P->pos_bytes == 2 && a.pos)
    k_matrix_cells<E, uint16_t, MODE><<<grid, 128, 0, st>>>(a);
  else
    k_matrix_cells<E, uint8_t, MODE><<<grid, 128, 0, st>>>(a);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

template <int MODE>
int launch_q1(const bfx_asm* P, const AsmArgs& a, cudaStream_t st)
{
  if (a.n == 0)
    return BFX_OK;
  if constexpr (MODE == 0)
  {
    static bool configured = false;
    if (!configured)
    {
      BFX_CUDA(cudaFuncSetAttribute(k_elasticity_q1_sym<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(Q2Smem)));
      BFX_CUDA(cudaFuncSetAttribute(k_elasticity_q1_sym<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(Q2Smem)));
      configured = true;
    }
    // a device-side count means "few cells" (the non-affine remainder of the row-gather path): small grid
    const unsigned grid = grid_for((a.n + Q2_CELLS - 1) / Q2_CELLS, 1, a.n_dev ? 2 : 0);
    if (P->pos_bytes == 2 && a.pos)
      k_elasticity_q1_sym<uint16_t><<<grid, Q2_THREADS, sizeof(Q2Smem), st>>>(a);
    else
      k_elasticity_q1_sym<uint8_t><<<grid, Q2_THREADS, sizeof(Q2Smem), st>>>(a);
  }
  else
  {
    const unsigned grid = grid_for((a.n + Q1_CELLS - 1) / Q1_CELLS, 1, a.n_dev ? 8 : 0);
    if (P->pos_bytes == 2 && a.pos)
      k_elasticity_q1<uint16_t, MODE><<<grid, 64 * Q1_CELLS, 0, st>>>(a);
    else
      k_elasticity_q1<uint8_t, MODE><<<grid, 64 * Q1_CELLS, 0, st>>>(a);
  }
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

template <class E>
int launch_vector(const AsmArgs& a, cudaStream_t st)
{
  if (a.n == 0)
    return BFX_OK;
  k_vector_cells<E><<<grid_for(a.n, 128, 0), 128, 0, st>>>(a);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

int launch_user(const bfx_asm* P, int kernel_id, const AsmArgs& a, int mode, double* scalar_dev, cudaStream_t st);

template <int MODE>
int dispatch_matrix(const bfx_asm* P, int kernel_id, const AsmArgs& a, cudaStream_t st)
{
  switch (kernel_id)
  {
  case BFX_K_LAPLACE_P1_TRI_A: return launch_matrix<el::LaplaceP1Tri, MODE>(P, a, st);
  case BFX_K_MASS_COEFF_P1_TRI_A: return launch_matrix<el::MassCoeffP1Tri, MODE>(P, a, st);
  case BFX_K_FACET_MASS_P1_TRI_A: return launch_matrix<el::FacetMassP1Tri, MODE>(P, a, st);
  case BFX_K_POISSON_P1_TET_A: return launch_matrix<el::PoissonP1Tet, MODE>(P, a, st);
  case BFX_K_POISSON_P2_TET_A: return launch_matrix<el::PoissonP2Tet, MODE>(P, a, st);
  case BFX_K_FACET_MASS_P1_TET_A: return launch_matrix<el::FacetMassP1Tet, MODE>(P, a, st);
  case BFX_K_AVG_MASS_P1_TRI_DS: return launch_matrix<el::AvgMassP1TriDS, MODE>(P, a, st);
  case BFX_K_ELASTICITY_Q1_HEX_A: return launch_q1<MODE>(P, a, st);
  default:
    if (kernel_id >= BFX_K_USER_BASE)
      return launch_user(P, kernel_id, a, MODE, nullptr, st);
    return fail(BFX_ERR_UNSUPPORTED, "kernel id %d is not a bilinear-form kernel", kernel_id);
  }
}

int dispatch_vector(int kernel_id, const AsmArgs& a, cudaStream_t st)
{
  switch (kernel_id)
  {
  case BFX_K_SOURCE_P1_TRI_L: return launch_vector<el::SourceP1Tri>(a, st);
  case BFX_K_LOAD_COEFF_P1_TRI_L: return launch_vector<el::LoadCoeffP1Tri>(a, st);
  case BFX_K_FACET_CONST_P1_TRI_L: return launch_vector<el::FacetConstP1Tri>(a, st);
  case BFX_K_LOAD_P1_TET_L: return launch_vector<el::LoadP1Tet>(a, st);
  case BFX_K_LOAD_P2_TET_L: return launch_vector<el::LoadP2Tet>(a, st);
  case BFX_K_LOAD_Q1_HEX_L: return launch_vector<el::LoadQ1Hex>(a, st);
  case BFX_K_FACET_LOAD_P1_TET_L: return launch_vector<el::FacetLoadP1Tet>(a, st);
  case BFX_K_ACTION_POISSON_P1_TET_L: return launch_vector<el::ActionOf<el::PoissonP1Tet>>(a, st);
  case BFX_K_ACTION_POISSON_P2_TET_L: return launch_vector<el::ActionOf<el::PoissonP2Tet>>(a, st);
  case BFX_K_AVG_LOAD_P1_TRI_DS_L: return launch_vector<el::AvgLoadP1TriDS>(a, st);
  case BFX_K_LOAD_PROD_P1_TET_L: return launch_vector<el::LoadProdP1Tet>(a, st);
  default:
    if (kernel_id >= BFX_K_USER_BASE)
      return launch_user(nullptr, kernel_id, a, 2, nullptr, st);
    return fail(BFX_ERR_UNSUPPORTED, "kernel id %d is not a linear-form kernel", kernel_id);
  }
}

const bfx_kernel_info_t KINFO[BFX_K_COUNT] = {
    /* nx nd bs rank w c facet */
    {3, 3, 1, 2, 0, 0, 0},  {3, 3, 1, 1, 0, 0, 0},  {3, 3, 1, 2, 3, 0, 0},  {3, 3, 1, 1, 3, 0, 0},
    {3, 3, 1, 2, 0, 0, 1},  {3, 3, 1, 1, 0, 1, 1},  {4, 4, 1, 2, 0, 1, 0},  {4, 4, 1, 1, 4, 0, 0},
    {4, 10, 1, 2, 0, 1, 0}, {4, 10, 1, 1, 10, 0, 0}, {8, 8, 3, 2, 0, 2, 0}, {8, 8, 3, 1, 24, 0, 0},
    {4, 4, 1, 1, 4, 0, 1},  {4, 4, 1, 2, 0, 0, 1},  {0, 0, 0, -1, 0, 0, 0}, {4, 4, 1, 1, 4, 1, 0},
    {4, 10, 1, 1, 10, 1, 0}, {4, 4, 1, 0, 4, 0, 0}, {6, 6, 1, 2, 0, 0, 1}, {6, 6, 1, 1, 0, 0, 1},
    {6, 6, 1, 0, 0, 0, 1},  {6, 6, 1, 0, 6, 0, 1},  {3, 3, 1, 0, 3, 0, 1},  {4, 4, 1, 1, 8, 0, 0}};
// coefficients gathered by the fused path of a kernel (all of the same layout): 1 unless listed
int kernel_ncoef(int kernel_id) { return kernel_id == BFX_K_LOAD_PROD_P1_TET_L ? 2 : 1; }

// kernels registered by the caller (bfx_register_kernel, include/bfx_plugin.cuh)
struct UserKernel
{
  bfx_kernel_info_t info;
  bfx_user_launch_t launch;
};
std::map<int, UserKernel>& user_kernels()
{
  static std::map<int, UserKernel> m;
  return m;
}
const bfx_kernel_info_t* kinfo(int kernel_id)
{
  if (kernel_id >= 0 && kernel_id < BFX_K_COUNT)
    return KINFO[kernel_id].rank >= 0 ? &KINFO[kernel_id] : nullptr;
  auto it = user_kernels().find(kernel_id);
  return it == user_kernels().end() ? nullptr : &it->second.info;
}
int launch_user(const bfx_asm* P, int kernel_id, const AsmArgs& a, int mode, double* scalar_dev, cudaStream_t st)
{
  auto it = user_kernels().find(kernel_id);
  if (it == user_kernels().end())
    return fail(BFX_ERR_UNSUPPORTED, "kernel id %d is neither built in nor registered", kernel_id);
  const int e = it->second.launch(&a, P ? P->pos_bytes : 1, mode, scalar_dev, st);
  if (e != BFX_OK)
    return fail(e, "registered kernel %d: launch failed (mode %d)", kernel_id, mode);
  return BFX_OK;
}

int fill_common(const bfx_asm* P, int kernel_id, int rank, const double* x, const bfx_coeffs_t* coeffs,
                const double* constants, int n_constants, AsmArgs& a, bool need_csr = true)
{
  BFX_REQUIRE(P && x, "assemble: null plan or geometry");
  BFX_REQUIRE(kinfo(kernel_id), "assemble: bad kernel id %d", kernel_id);
  const bfx_kernel_info_t& ki = *kinfo(kernel_id);
  BFX_REQUIRE(ki.rank == rank, "kernel id %d has rank %d, expected %d", kernel_id, ki.rank, rank);
  BFX_REQUIRE(ki.nx == P->nx && ki.nd == P->nd0, "kernel id %d expects nx=%d nd=%d, plan has nx=%d nd=%d", kernel_id,
              ki.nx, ki.nd, P->nx, P->nd0);
  if (rank == 2)
  {
    BFX_REQUIRE(P->nd1 == ki.nd, "bilinear kernel %d needs a plan with a trial dofmap of %d dofs per cell", kernel_id,
                ki.nd);
    if (need_csr)
      BFX_REQUIRE(P->csr && P->csr->bs0 == ki.bs && P->csr->bs1 == ki.bs,
                  "bilinear kernel %d needs a plan with a matching MatrixCSR (bs=%d)", kernel_id, ki.bs);
  }
  BFX_REQUIRE(n_constants >= ki.c_size && n_constants <= 8, "kernel id %d needs %d constants (max 8), got %d",
              kernel_id, ki.c_size, n_constants);
  memset(&a, 0, sizeof(a));
  a.x_dofmap = P->x_dofmap;
  a.dofmap0 = P->dofmap0;
  a.dofmap1 = P->dofmap1 ? P->dofmap1 : P->dofmap0;
  a.x = x;
  for (int k = 0; k < n_constants; ++k)
    a.constants[k] = constants[k];
  if (ki.w_size > 0)
  {
    BFX_REQUIRE(coeffs, "kernel id %d needs coefficients", kernel_id);
    if (coeffs->packed_dev)
    {
      BFX_REQUIRE(coeffs->cstride >= ki.w_size, "packed coefficient stride %d < %d", coeffs->cstride, ki.w_size);
      a.coef.packed = coeffs->packed_dev;
      a.coef.cstride = coeffs->cstride;
      a.coef.f[0].off = coeffs->n_fused > 0 ? coeffs->fused[0].offset : 0;
    }
    else
    {
      // (registered kernels: as many coefficients of one layout as fill w)
      const int nc = kernel_id >= BFX_K_USER_BASE ? (coeffs->n_fused > 0 ? coeffs->n_fused : 1) : kernel_ncoef(kernel_id);
      BFX_REQUIRE(coeffs->n_fused == nc && nc <= 4, "kernel id %d gathers %d coefficient(s), %d given", kernel_id, nc, coeffs->n_fused);
      for (int k = 0; k < nc; ++k)
      {
        BFX_REQUIRE(coeffs->fused[k].values_dev && coeffs->fused[k].dofmap_dev, "fused coefficient %d: null array", k);
        BFX_REQUIRE(coeffs->fused[k].nd * coeffs->fused[k].bs * nc == ki.w_size,
                    "coefficient layout nd*bs=%d does not match kernel w size %d / %d", coeffs->fused[k].nd * coeffs->fused[k].bs,
                    ki.w_size, nc);
        a.coef.f[k].v = coeffs->fused[k].values_dev;
        a.coef.f[k].dm = coeffs->fused[k].dofmap_dev;
      }
    }
  }
  if (P->csr)
  {
    a.row_ptr = P->csr->row_ptr;
    a.cols = P->csr->cols;
    a.err = P->csr->err_flag;
  }
  return BFX_OK;
}

} // namespace

namespace bfx
{
int launch_q1_red(const bfx_asm* P, const AsmArgs& a, cudaStream_t st) { return launch_q1<0>(P, a, st); }
} // namespace bfx

namespace
{
int check_err(const bfx_asm* P, cudaStream_t st, const char* what)
{
  int h = 0;
  BFX_CUDA(cudaMemcpyAsync(&h, P->csr->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  BFX_CUDA(cudaStreamSynchronize(st));
  if (h)
  {
    BFX_CUDA(cudaMemsetAsync(P->csr->err_flag, 0, sizeof(int), st));
    return fail(h == 1 ? BFX_ERR_NOT_IN_SPARSITY : BFX_ERR_INVALID,
                h == 1 ? "%s: Entry not in sparsity" : "%s: dof index out of range", what);
  }
  return BFX_OK;
}
} // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C"
{
int bfx_kernel_info(int kernel_id, bfx_kernel_info_t* info)
{
  BFX_REQUIRE(info && (kinfo(kernel_id) || (kernel_id >= 0 && kernel_id < BFX_K_COUNT)), "bfx_kernel_info: bad kernel id %d", kernel_id);
  *info = kinfo(kernel_id) ? *kinfo(kernel_id) : KINFO[kernel_id];
  return BFX_OK;
}

int bfx_register_kernel(int kernel_id, const bfx_kernel_info_t* info, bfx_user_launch_t launch)
{
  BFX_REQUIRE(kernel_id >= BFX_K_USER_BASE && info && launch, "bfx_register_kernel: ids start at BFX_K_USER_BASE (%d)", BFX_K_USER_BASE);
  BFX_REQUIRE(info->nx > 0 && info->nd > 0 && info->bs > 0 && info->rank >= 0 && info->rank <= 2 && info->w_size >= 0
                  && info->c_size >= 0 && info->c_size <= 8,
              "bfx_register_kernel: inconsistent kernel description");
  user_kernels()[kernel_id] = UserKernel{*info, launch};
  return BFX_OK;
}

int bfx_asm_create(bfx_asm_t** out, const bfx_csr_t* csr, const int32_t* x_dofmap, int nx, const int32_t* dofmap0,
                   int nd0, const int32_t* dofmap1, int nd1, int64_t ncells_all, const int32_t* cells, int64_t ncells,
                   int32_t n_rows_all, int borrow, bfx_stream_t stream)
{
  // (a mesh without cells - an empty rank, an interior-facet domain without facets - has empty arrays: NULL is fine then)
  BFX_REQUIRE(out && nx > 0 && nd0 > 0 && ncells_all >= 0 && ncells >= 0 && ((x_dofmap && dofmap0) || ncells_all == 0),
              "bfx_asm_create: bad arguments");
  cudaStream_t st = S(stream);
  bfx_asm* P = new bfx_asm();
  P->csr = csr;
  P->nx = nx;
  P->nd0 = nd0;
  P->nd1 = dofmap1 ? nd1 : (csr ? nd0 : 0);
  P->ncells_all = ncells_all;
  P->ncells = ncells;
  P->n_rows_all = n_rows_all;
  P->owns = !borrow;
  int e = BFX_OK;
  if (borrow)
  {
    P->x_dofmap = const_cast<int32_t*>(x_dofmap);
    P->dofmap0 = const_cast<int32_t*>(dofmap0);
    P->dofmap1 = dofmap1 && dofmap1 != dofmap0 ? const_cast<int32_t*>(dofmap1) : nullptr;
    P->cells = const_cast<int32_t*>(cells);
  }
  else
  {
    if ((e = upload(&P->x_dofmap, x_dofmap, (size_t)ncells_all * nx, st))
        || (e = upload(&P->dofmap0, dofmap0, (size_t)ncells_all * nd0, st)))
      return e;
    if (dofmap1 && dofmap1 != dofmap0)
    {
      if ((e = upload(&P->dofmap1, dofmap1, (size_t)ncells_all * nd1, st)))
        return e;
    }
    if (cells)
    {
      if ((e = upload(&P->cells, cells, (size_t)ncells, st)))
        return e;
    }
  }
  if (csr && ncells > 0)
  {
    BFX_REQUIRE(n_rows_all == csr->n_rows_all, "bfx_asm_create: plan rows %d != matrix rows %d", n_rows_all,
                csr->n_rows_all);
    int* d_max = nullptr;
    BFX_CUDA(cudaMalloc(&d_max, sizeof(int)));
    BFX_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
    k_max_row_len<<<grid_for(csr->n_rows_all, 256, 8), 256, 0, st>>>(csr->n_rows_all, csr->row_ptr, d_max);
    int h_max = 0;
    BFX_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaStreamSynchronize(st));
    BFX_CUDA(cudaFree(d_max));
    if (h_max > 65536)
      return fail(BFX_ERR_UNSUPPORTED, "rows with more than 65536 block entries are not supported (%d)", h_max);
    P->pos_bytes = h_max <= 256 ? 1 : 2;
    const int count = P->nd0 * P->nd1;
    P->pos_stride = (count * P->pos_bytes + 15) / 16 * 16;
    if (P->nd0 == 8 && P->nd1 == 8 && csr->bs0 == 3)
      P->pos_stride = 64 * P->pos_bytes; // Q1 elasticity kernel reads one entry per thread
    if ((e = dev_alloc(&P->pos, (size_t)ncells * P->pos_stride)))
      return e;
    const int32_t* dm1 = P->dofmap1 ? P->dofmap1 : P->dofmap0;
    const unsigned grid = grid_for(ncells * nd0, 256, 0);
    if (P->pos_bytes == 1)
      k_build_pos<uint8_t><<<grid, 256, 0, st>>>(ncells, P->cells, P->dofmap0, P->nd0, dm1, P->nd1, csr->row_ptr,
                                                 csr->cols, csr->n_rows_all, P->pos, P->pos_stride, csr->err_flag);
    else
      k_build_pos<uint16_t><<<grid, 256, 0, st>>>(ncells, P->cells, P->dofmap0, P->nd0, dm1, P->nd1, csr->row_ptr,
                                                  csr->cols, csr->n_rows_all, P->pos, P->pos_stride, csr->err_flag);
    BFX_CHECK_LAUNCH();
    if ((e = check_err(P, st, "bfx_asm_create")))
    {
      bfx_asm_destroy(P);
      return e;
    }
  }
  BFX_CUDA(cudaStreamSynchronize(st));
  *out = P;
  return BFX_OK;
}

int bfx_asm_destroy(bfx_asm_t* P)
{
  if (!P)
    return BFX_OK;
  if (P->owns)
  {
    cudaFree(P->x_dofmap);
    cudaFree(P->dofmap0);
    cudaFree(P->dofmap1);
    cudaFree(P->cells);
  }
  cudaFree(P->pos);
  cudaFree(P->lift_cells);
  cudaFree(P->lift_count);
  free_chunks(P->chunks);
  free_rowgather(P->rowgather);
  for (auto& h : P->hs)
  {
    cudaFree(h.x);
    cudaFree(h.coeff);
    cudaFree(h.bc0);
    cudaFree(h.bc1);
    cudaFree(h.values);
    if (h.st)
      cudaStreamDestroy(h.st);
    if (h.done)
      cudaEventDestroy(h.done);
    if (h.computed)
      cudaEventDestroy(h.computed);
  }
  delete P;
  return BFX_OK;
}

int bfx_assemble_matrix_cells(const bfx_asm_t* P, int kernel_id, const double* x, const int8_t* bc0,
                              const int8_t* bc1, const bfx_coeffs_t* coeffs, const double* constants, int n_constants,
                              double* values, int strategy, int values_mode, bfx_stream_t stream)
{
  AsmArgs a;
  int e = fill_common(P, kernel_id, 2, x, coeffs, constants, n_constants, a);
  if (e)
    return e;
  BFX_REQUIRE(values, "bfx_assemble_matrix_cells: null values");
  BFX_REQUIRE(!kinfo(kernel_id)->facet, "kernel id %d is a facet kernel", kernel_id);
  if (strategy != BFX_ASM_ATOMIC && strategy != BFX_ASM_CHUNKED && strategy != BFX_ASM_ROWGATHER)
    return fail(BFX_ERR_UNSUPPORTED, "assembly strategy %d not available", strategy);
  a.cells = P->cells;
  a.n = P->ncells;
  a.bc0 = bc0;
  a.bc1 = bc1;
  a.values = values;
  a.pos = P->pos;
  if (strategy == BFX_ASM_CHUNKED)
  {
    // a symmetric chunk plan refuses calls whose rows and columns differ in dofmap or markers: those (rare)
    // calls take the RED kernel, which has no such requirement
    const int st = launch_chunked(P, kernel_id, a, values_mode, S(stream));
    if (st != BFX_ERR_UNSUPPORTED || !P->chunks)
      return st;
    return dispatch_matrix<0>(P, kernel_id, a, S(stream));
  }
  if (strategy == BFX_ASM_ROWGATHER)
  {
    if (kernel_id != BFX_K_ELASTICITY_Q1_HEX_A)
      return fail(BFX_ERR_UNSUPPORTED, "kernel id %d has no row-gather variant", kernel_id);
    return launch_rowgather_q1(P, a, values_mode, S(stream));
  }
  return dispatch_matrix<0>(P, kernel_id, a, S(stream));
}

int bfx_assemble_matrix_cells_part(bfx_asm_t* P, int kernel_id, const double* x, const int8_t* bc0, const int8_t* bc1,
                                   const bfx_coeffs_t* coeffs, const double* constants, int n_constants, double* values,
                                   int values_mode, int part, bfx_stream_t stream)
{
  BFX_REQUIRE(P && P->chunks && part >= 0 && part <= 2, "bfx_assemble_matrix_cells_part: needs a partitioned chunk plan");
  P->chunks->launch_part = part;
  const int e = bfx_assemble_matrix_cells(P, kernel_id, x, bc0, bc1, coeffs, constants, n_constants, values, BFX_ASM_CHUNKED,
                                          values_mode, stream);
  P->chunks->launch_part = 0;
  return e;
}

int bfx_assemble_matrix_rows(const bfx_asm_t* P, int kernel_id, const double* x, const int8_t* bc0, const int8_t* bc1,
                             const bfx_coeffs_t* coeffs, const double* constants, int n_constants, double* values,
                             int32_t row_begin, int32_t row_end, int reuse_records, bfx_stream_t stream)
{
  AsmArgs a;
  int e = fill_common(P, kernel_id, 2, x, coeffs, constants, n_constants, a);
  if (e)
    return e;
  BFX_REQUIRE(values, "bfx_assemble_matrix_rows: null values");
  if (kernel_id != BFX_K_ELASTICITY_Q1_HEX_A || !P->rowgather)
    return fail(BFX_ERR_UNSUPPORTED, "bfx_assemble_matrix_rows: kernel id %d / plan has no row-gather variant", kernel_id);
  const int tr = rowgather_tile_rows(P);
  BFX_REQUIRE(tr > 0 && row_begin >= 0 && row_begin <= row_end && row_end <= P->csr->n_rows_all
                  && row_begin % tr == 0 && (row_end % tr == 0 || row_end == P->csr->n_rows_all),
              "bfx_assemble_matrix_rows: row range [%d, %d) must be cut at multiples of %d rows", row_begin, row_end, tr);
  a.cells = P->cells;
  a.n = P->ncells;
  a.bc0 = bc0;
  a.bc1 = bc1;
  a.values = values;
  a.pos = P->pos;
  return launch_rowgather_q1(P, a, BFX_VALUES_OVERWRITE, S(stream), row_begin, row_end, reuse_records != 0);
}

int bfx_asm_rowgather_tile_rows(const bfx_asm_t* P, int* rows)
{
  BFX_REQUIRE(P && rows, "bfx_asm_rowgather_tile_rows: null argument");
  *rows = rowgather_tile_rows(P);
  return BFX_OK;
}

int bfx_assemble_matrix_facets(const bfx_asm_t* P, int kernel_id, const double* x, const int32_t* entities,
                               int64_t n_entities, const int8_t* bc0, const int8_t* bc1, const bfx_coeffs_t* coeffs,
                               const double* constants, int n_constants, double* values, bfx_stream_t stream)
{
  AsmArgs a;
  int e = fill_common(P, kernel_id, 2, x, coeffs, constants, n_constants, a);
  if (e)
    return e;
  BFX_REQUIRE(values && (entities || n_entities == 0), "bfx_assemble_matrix_facets: null argument");
  BFX_REQUIRE(kinfo(kernel_id)->facet, "kernel id %d is not a facet kernel", kernel_id);
  a.entities = entities;
  a.n = n_entities;
  a.bc0 = bc0;
  a.bc1 = bc1;
  a.values = values;
  a.pos = nullptr; // few entities: locate entries by binary search like insert_csr
  if ((e = dispatch_matrix<0>(P, kernel_id, a, S(stream))))
    return e;
  return check_err(P, S(stream), "assemble exterior facets");
}

int bfx_lift_bc_cells(const bfx_asm_t* P, int kernel_id, const double* x, const bfx_coeffs_t* coeffs,
                      const double* constants, int n_constants, double* b, const double* bc_values1,
                      const int8_t* bc_markers1, const double* x0, double alpha, bfx_stream_t stream)
{
  AsmArgs a;
  int e = fill_common(P, kernel_id, 2, x, coeffs, constants, n_constants, a, false);
  if (e)
    return e;
  BFX_REQUIRE(b && bc_values1 && bc_markers1, "bfx_lift_bc_cells: null argument");
  a.cells = P->cells;
  a.n = P->ncells;
  a.bc1 = bc_markers1;
  a.b = b;
  a.bc_values1 = bc_values1;
  a.x0 = x0;
  a.alpha = alpha;
  if (kernel_id == BFX_K_ELASTICITY_Q1_HEX_A && P->ncells > 0)
  {
    // the 64-threads-per-cell elasticity kernel is too heavy to visit every cell for a test: compact first
    bfx_asm* Pm = const_cast<bfx_asm*>(P);
    if (!Pm->lift_cells)
    {
      BFX_CUDA(cudaMalloc(&Pm->lift_cells, sizeof(int32_t) * (size_t)P->ncells));
      BFX_CUDA(cudaMalloc(&Pm->lift_count, sizeof(unsigned long long)));
    }
    cudaStream_t st = S(stream);
    BFX_CUDA(cudaMemsetAsync(Pm->lift_count, 0, sizeof(unsigned long long), st));
    k_cells_with_bc<<<grid_for(P->ncells, 256, 16), 256, 0, st>>>(P->ncells, P->cells, a.dofmap1, P->nd1,
                                                                  kinfo(kernel_id)->bs, bc_markers1, Pm->lift_cells,
                                                                  Pm->lift_count);
    a.cells = Pm->lift_cells;
    a.n_dev = Pm->lift_count;
  }
  return dispatch_matrix<1>(P, kernel_id, a, S(stream));
}

int bfx_assemble_vector_cells(const bfx_asm_t* P, int kernel_id, const double* x, const bfx_coeffs_t* coeffs,
                              const double* constants, int n_constants, double* b, int strategy, bfx_stream_t stream)
{
  AsmArgs a;
  int e = fill_common(P, kernel_id, 1, x, coeffs, constants, n_constants, a);
  if (e)
    return e;
  BFX_REQUIRE(b, "bfx_assemble_vector_cells: null b");
  BFX_REQUIRE(!kinfo(kernel_id)->facet, "kernel id %d is a facet kernel", kernel_id);
  if (strategy != BFX_ASM_ATOMIC && strategy != BFX_ASM_CHUNKED)
    return fail(BFX_ERR_UNSUPPORTED, "assembly strategy %d not available", strategy);
  a.cells = P->cells;
  a.n = P->ncells;
  a.b = b;
  if (strategy == BFX_ASM_CHUNKED)
  {
    // (a plan reduced to its warp tables refuses calls the table kernel cannot serve - packed coefficients, a
    // coefficient on another dofmap: those take the RED kernel, like the matrix path does)
    const int st = launch_vector_grouped(P, kernel_id, a, S(stream));
    if (st != BFX_ERR_UNSUPPORTED)
      return st;
  }
  return dispatch_vector(kernel_id, a, S(stream));
}

namespace
{
// sum of a functional kernel over the cells / entities of `a`, brought to the host
int run_scalar(int kernel_id, const AsmArgs& a, double* result_host, cudaStream_t st)
{
  double* d_res = nullptr;
  BFX_CUDA(cudaMalloc(&d_res, sizeof(double)));
  BFX_CUDA(cudaMemsetAsync(d_res, 0, sizeof(double), st));
  if (a.n > 0)
  {
    const unsigned grid = grid_for(a.n, 256, 8);
    switch (kernel_id)
    {
    case BFX_K_L2NORM2_P1_TET_M: k_scalar_cells<el::L2Norm2P1Tet><<<grid, 256, 0, st>>>(a, d_res); break;
    case BFX_K_ONE_TRI_DS_M: k_scalar_cells<el::OneTriDS><<<grid, 256, 0, st>>>(a, d_res); break;
    case BFX_K_AVG2_COEFF_P1_TRI_DS_M: k_scalar_cells<el::Avg2CoeffP1TriDS><<<grid, 256, 0, st>>>(a, d_res); break;
    case BFX_K_COEFF2_P1_TRI_FACET_M: k_scalar_cells<el::Coeff2P1TriFacet><<<grid, 256, 0, st>>>(a, d_res); break;
    default:
      if (kernel_id >= BFX_K_USER_BASE)
      {
        const int eu = launch_user(nullptr, kernel_id, a, 3, d_res, st);
        if (eu)
        {
          cudaFree(d_res);
          return eu;
        }
        break;
      }
      cudaFree(d_res);
      return fail(BFX_ERR_UNSUPPORTED, "kernel id %d is not a functional kernel", kernel_id);
    }
  }
  BFX_CUDA(cudaMemcpyAsync(result_host, d_res, sizeof(double), cudaMemcpyDeviceToHost, st));
  BFX_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_res);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}
} // namespace

int bfx_assemble_scalar_cells(const bfx_asm_t* P, int kernel_id, const double* x, const bfx_coeffs_t* coeffs,
                              const double* constants, int n_constants, double* result_host, bfx_stream_t stream)
{
  AsmArgs a;
  int e = fill_common(P, kernel_id, 0, x, coeffs, constants, n_constants, a);
  if (e)
    return e;
  BFX_REQUIRE(result_host, "bfx_assemble_scalar_cells: null result");
  BFX_REQUIRE(!kinfo(kernel_id)->facet, "kernel id %d is a facet kernel", kernel_id);
  a.cells = P->cells;
  a.n = P->ncells;
  return run_scalar(kernel_id, a, result_host, S(stream));
}

int bfx_assemble_scalar_facets(const bfx_asm_t* P, int kernel_id, const double* x, const int32_t* entities,
                               int64_t n_entities, const bfx_coeffs_t* coeffs, const double* constants, int n_constants,
                               double* result_host, bfx_stream_t stream)
{
  AsmArgs a;
  int e = fill_common(P, kernel_id, 0, x, coeffs, constants, n_constants, a);
  if (e)
    return e;
  BFX_REQUIRE(result_host && (entities || n_entities == 0), "bfx_assemble_scalar_facets: null argument");
  BFX_REQUIRE(kinfo(kernel_id)->facet, "kernel id %d is not a facet kernel", kernel_id);
  a.entities = entities;
  a.n = n_entities;
  return run_scalar(kernel_id, a, result_host, S(stream));
}

int bfx_assemble_vector_facets(const bfx_asm_t* P, int kernel_id, const double* x, const int32_t* entities,
                               int64_t n_entities, const bfx_coeffs_t* coeffs, const double* constants,
                               int n_constants, double* b, bfx_stream_t stream)
{
  AsmArgs a;
  int e = fill_common(P, kernel_id, 1, x, coeffs, constants, n_constants, a);
  if (e)
    return e;
  BFX_REQUIRE(b && (entities || n_entities == 0), "bfx_assemble_vector_facets: null argument");
  BFX_REQUIRE(kinfo(kernel_id)->facet, "kernel id %d is not a facet kernel", kernel_id);
  a.entities = entities;
  a.n = n_entities;
  a.b = b;
  return dispatch_vector(kernel_id, a, S(stream));
}

int bfx_pack_coefficient(double* coeffs, int cstride, int offset, const double* values, const int32_t* dofmap, int nd,
                         int bs, const int32_t* cells, const int32_t* entities, int64_t n, bfx_stream_t stream)
{
  BFX_REQUIRE(coeffs && values && dofmap && nd > 0 && bs > 0, "bfx_pack_coefficient: bad arguments");
  if (n == 0)
    return BFX_OK;
  k_pack<<<grid_for(n * nd * bs, 256, 16), 256, 0, S(stream)>>>(coeffs, cstride, offset, values, dofmap, nd, bs, cells,
                                                                 entities, n);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

int bfx_bc_mark(int8_t* markers, const int32_t* dofs0, int64_t n, bfx_stream_t stream)
{
  if (n == 0)
    return BFX_OK;
  BFX_REQUIRE(markers && dofs0, "bfx_bc_mark: null argument");
  k_bc_mark<<<grid_for(n, 256, 8), 256, 0, S(stream)>>>(markers, dofs0, n);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

int bfx_bc_set(double* x, int32_t x_size, const int32_t* dofs0, const int32_t* dofs_g, int64_t n, const double* g,
               int g_kind, int bs, const double* x0, double alpha, bfx_stream_t stream)
{
  if (n == 0)
    return BFX_OK;
  BFX_REQUIRE(x && dofs0 && g && bs > 0, "bfx_bc_set: bad arguments");
  k_bc_set<<<grid_for(n, 256, 8), 256, 0, S(stream)>>>(x, x_size, dofs0, dofs_g, n, g, g_kind, bs, x0, alpha);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

int bfx_assemble_matrix_cells_host_begin(bfx_asm_t* P, int kernel_id, const double* x_host, int64_t n_x_nodes,
                                         const int8_t* bc0_host, const int8_t* bc1_host, int64_t n_bc,
                                         const double* coeff_values_host, int64_t n_coeff_values, int coeff_bs,
                                         const double* constants, int n_constants, double* values_host, int strategy)
{
  BFX_REQUIRE(P && P->csr && x_host && values_host, "bfx_assemble_matrix_cells_host: null argument");
  bfx_asm::HostSlot& h = P->hs[P->hs_next];
  if (h.busy)
    return fail(BFX_ERR_INVALID, "bfx_assemble_matrix_cells_host_begin: two calls are already in flight (call _end)");
  if (!h.st)
  {
    BFX_CUDA(cudaStreamCreateWithFlags(&h.st, cudaStreamNonBlocking));
    BFX_CUDA(cudaEventCreateWithFlags(&h.done, cudaEventDisableTiming));
    BFX_CUDA(cudaEventCreateWithFlags(&h.computed, cudaEventDisableTiming));
  }
  cudaStream_t st = h.st;
  const size_t nval = (size_t)P->csr->nnz * P->csr->bs0 * P->csr->bs1;
  // device scratch is allocated on first use and kept by the plan
  if (!h.x || h.x_n < n_x_nodes)
  {
    cudaFree(h.x);
    BFX_CUDA(cudaMalloc(&h.x, sizeof(double) * 3 * (size_t)n_x_nodes));
    h.x_n = n_x_nodes;
  }
  if (!h.values)
    BFX_CUDA(cudaMalloc(&h.values, sizeof(double) * nval));
  if (n_bc > 0 && (!h.bc0 || h.bc_n < n_bc))
  {
    cudaFree(h.bc0);
    cudaFree(h.bc1);
    BFX_CUDA(cudaMalloc(&h.bc0, (size_t)n_bc));
    BFX_CUDA(cudaMalloc(&h.bc1, (size_t)n_bc));
    h.bc_n = n_bc;
  }
  if (n_coeff_values > 0 && (!h.coeff || h.coeff_n < n_coeff_values))
  {
    cudaFree(h.coeff);
    BFX_CUDA(cudaMalloc(&h.coeff, sizeof(double) * (size_t)n_coeff_values));
    h.coeff_n = n_coeff_values;
  }
  BFX_CUDA(cudaMemcpyAsync(h.x, x_host, sizeof(double) * 3 * (size_t)n_x_nodes, cudaMemcpyHostToDevice, st));
  if (bc0_host)
    BFX_CUDA(cudaMemcpyAsync(h.bc0, bc0_host, (size_t)n_bc, cudaMemcpyHostToDevice, st));
  const bool one_marker_array = bc1_host && bc1_host == bc0_host; // same space: rows and columns share the markers
  if (bc1_host && !one_marker_array)
    BFX_CUDA(cudaMemcpyAsync(h.bc1, bc1_host, (size_t)n_bc, cudaMemcpyHostToDevice, st));
  bfx_coeffs_t cf;
  memset(&cf, 0, sizeof(cf));
  if (coeff_values_host && n_coeff_values > 0)
  {
    BFX_CUDA(cudaMemcpyAsync(h.coeff, coeff_values_host, sizeof(double) * (size_t)n_coeff_values,
                             cudaMemcpyHostToDevice, st));
    cf.n_fused = 1;
    cf.fused[0].values_dev = h.coeff;
    cf.fused[0].dofmap_dev = P->dofmap0;
    cf.fused[0].nd = P->nd0;
    cf.fused[0].bs = coeff_bs;
  }
  BFX_CUDA(cudaMemsetAsync(h.values, 0, sizeof(double) * nval, st));
  // the assembly kernels of the two slots share the plan's per-call scratch (packed marker bits, row-gather cell
  // records): they run one after the other; only the copies of one slot overlap the kernels of the other
  const bfx_asm::HostSlot& other = P->hs[P->hs_next ^ 1];
  if (other.computed && other.used)
    BFX_CUDA(cudaStreamWaitEvent(st, other.computed, 0));
  int e = bfx_assemble_matrix_cells(P, kernel_id, h.x, bc0_host ? h.bc0 : nullptr,
                                    bc1_host ? (one_marker_array ? h.bc0 : h.bc1) : nullptr, &cf, constants,
                                    n_constants, h.values, strategy, BFX_VALUES_OVERWRITE,
                                    reinterpret_cast<bfx_stream_t>(st));
  if (e)
    return e;
  BFX_CUDA(cudaEventRecord(h.computed, st));
  h.used = true;
  BFX_CUDA(cudaMemcpyAsync(values_host, h.values, sizeof(double) * nval, cudaMemcpyDeviceToHost, st));
  BFX_CUDA(cudaEventRecord(h.done, st));
  h.busy = true;
  P->hs_next ^= 1;
  return BFX_OK;
}

int bfx_assemble_matrix_cells_host_end(bfx_asm_t* P)
{
  BFX_REQUIRE(P, "bfx_assemble_matrix_cells_host_end: null plan");
  // the oldest call in flight: the slot begin() would use next if it is busy, else the other one
  int s = P->hs[P->hs_next].busy ? P->hs_next : (P->hs_next ^ 1);
  bfx_asm::HostSlot& h = P->hs[s];
  if (!h.busy)
    return fail(BFX_ERR_INVALID, "bfx_assemble_matrix_cells_host_end: no call in flight");
  BFX_CUDA(cudaEventSynchronize(h.done));
  h.busy = false;
  return BFX_OK;
}

int bfx_assemble_matrix_cells_host(bfx_asm_t* P, int kernel_id, const double* x_host, int64_t n_x_nodes,
                                   const int8_t* bc0_host, const int8_t* bc1_host, int64_t n_bc,
                                   const double* coeff_values_host, int64_t n_coeff_values, int coeff_bs,
                                   const double* constants, int n_constants, double* values_host, int strategy,
                                   bfx_stream_t stream)
{
  (void)stream; // the copies and kernels of this entry run on the plan's own streams
  BFX_REQUIRE(P && !P->hs[0].busy && !P->hs[1].busy, "bfx_assemble_matrix_cells_host: asynchronous calls are in flight");
  int e = bfx_assemble_matrix_cells_host_begin(P, kernel_id, x_host, n_x_nodes, bc0_host, bc1_host, n_bc,
                                               coeff_values_host, n_coeff_values, coeff_bs, constants, n_constants,
                                               values_host, strategy);
  if (e)
    return e;
  return bfx_assemble_matrix_cells_host_end(P);
}
}
