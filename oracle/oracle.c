/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * Plain-C CPU restatement of the DOLFINx hot path (assembly -> MatrixCSR ->
 * SpMV) used as the parity checker for the CUDA path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (dolfinx_b200/) never does.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/cpp/dolfinx).  The element kernels restate what FFCx
 * (third-party, fenics-ffcx >= 0.12.0.dev0, unpinned "main" in the reference's
 * CI) generates for the benchmark forms; they deliberately use a different
 * formulation from the CUDA kernels (numerical quadrature here, closed-form
 * pre-integration there) so that agreement is a real cross-check.
 *
 * Pinning: see tests/test_oracle_golden.py — the oracle reproduces the
 * reference's golden scalars (python/test/unit/fem/test_custom_jit_kernels.py:115-116,
 * test_ghost_mesh_assembly.py:64-66, cpp/test/matrix.cpp:116-119) and its
 * insert_csr / spmv agree with the reference's own la/matrix_csr_impl.h compiled
 * from /root/reference (oracle/_ref, container-local).
 *
 * Build: gcc -std=gnu11 -O2 -ffp-contract=off -fPIC -shared oracle.c -o liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef void (*orc_kernel_t)(double* A, const double* w, const double* c, const double* coordinate_dofs,
                             const int* entity_local_index, const uint8_t* quadrature_permutation, void* custom_data);

/* ------------------------------------------------------------------------- */
/* la/matrix_csr_impl.h                                                       */
/* ------------------------------------------------------------------------- */

static const int32_t* lower_bound_i32(const int32_t* first, const int32_t* last, int32_t v)
{
  /* std::lower_bound */
  long count = last - first;
  while (count > 0)
  {
    long step = count / 2;
    const int32_t* it = first + step;
    if (*it < v)
    {
      first = it + 1;
      count -= step + 1;
    }
    else
      count = step;
  }
  return first;
}

/* insert_csr<BS0,BS1>: la/matrix_csr_impl.h:67-109.  op: 0 = set, 1 = add.
 * Returns 0, or -1 for "Entry not in sparsity" (:93-94). */
int orc_insert_csr(double* data, const int32_t* cols, const int64_t* row_ptr, const double* x, const int32_t* xrows,
                   int nr, const int32_t* xcols, int nc, int BS0, int BS1, int op)
{
  for (int r = 0; r < nr; ++r)
  {
    int32_t row = xrows[r];
    const double* xr = x + (size_t)r * nc * BS0 * BS1;
    const int32_t* cit0 = cols + row_ptr[row];
    const int32_t* cit1 = cols + row_ptr[row + 1];
    for (int c = 0; c < nc; ++c)
    {
      const int32_t* it = lower_bound_i32(cit0, cit1, xcols[c]);
      if (it == cit1 || *it != xcols[c])
        return -1;
      size_t d = (size_t)(it - cols);
      size_t di = d * BS0 * BS1;
      size_t xi = (size_t)c * BS1;
      for (int i = 0; i < BS0; ++i)
      {
        for (int j = 0; j < BS1; ++j)
        {
          if (op)
            data[di + j] += xr[xi + j];
          else
            data[di + j] = xr[xi + j];
        }
        di += BS1;
        xi += (size_t)nc * BS1;
      }
    }
  }
  return 0;
}

/* insert_blocked_csr<BS0,BS1>: la/matrix_csr_impl.h:132-172 (blocked data into a bs=1 matrix) */
int orc_insert_blocked_csr(double* data, const int32_t* cols, const int64_t* row_ptr, const double* x,
                           const int32_t* xrows, int nr, const int32_t* xcols, int nc, int BS0, int BS1, int op)
{
  for (int r = 0; r < nr; ++r)
  {
    int32_t row = xrows[r] * BS0;
    for (int i = 0; i < BS0; ++i)
    {
      const double* xr = x + ((size_t)r * BS0 + i) * nc * BS1;
      const int32_t* cit0 = cols + row_ptr[row + i];
      const int32_t* cit1 = cols + row_ptr[row + i + 1];
      for (int c = 0; c < nc; ++c)
      {
        const int32_t* it = lower_bound_i32(cit0, cit1, xcols[c] * BS1);
        if (it == cit1 || *it != xcols[c] * BS1)
          return -1;
        size_t d = (size_t)(it - cols);
        size_t xi = (size_t)c * BS1;
        for (int j = 0; j < BS1; ++j)
        {
          if (op)
            data[d + j] += xr[xi + j];
          else
            data[d + j] = xr[xi + j];
        }
      }
    }
  }
  return 0;
}

/* insert_nonblocked_csr: la/matrix_csr_impl.h:193-232 (bs=1 data into a blocked matrix) */
int orc_insert_nonblocked_csr(double* data, const int32_t* cols, const int64_t* row_ptr, const double* x,
                              const int32_t* xrows, int nr, const int32_t* xcols, int nc, int bs0, int bs1, int op)
{
  const int nbs = bs0 * bs1;
  for (int r = 0; r < nr; ++r)
  {
    div_t rdiv = div(xrows[r], bs0);
    const double* xr = x + (size_t)r * nc;
    const int32_t* cit0 = cols + row_ptr[rdiv.quot];
    const int32_t* cit1 = cols + row_ptr[rdiv.quot + 1];
    for (int c = 0; c < nc; ++c)
    {
      div_t cdiv = div(xcols[c], bs1);
      const int32_t* it = lower_bound_i32(cit0, cit1, cdiv.quot);
      if (it == cit1 || *it != cdiv.quot)
        return -1;
      size_t d = (size_t)(it - cols);
      size_t di = d * nbs + rdiv.rem * bs1 + cdiv.rem;
      if (op)
        data[di] += xr[c];
      else
        data[di] = xr[c];
    }
  }
  return 0;
}

/* spmv: la/matrix_csr_impl.h:259-286 (y += A x over [row_begin, row_end)) */
void orc_spmv(const double* values, const int64_t* row_begin, const int64_t* row_end, const int32_t* indices,
              const double* x, double* y, int bs0, int bs1, int64_t nrows)
{
  for (int k0 = 0; k0 < bs0; ++k0)
  {
    for (int64_t i = 0; i < nrows; i++)
    {
      double vi = 0;
      for (int64_t j = row_begin[i]; j < row_end[i]; j++)
        for (int k1 = 0; k1 < bs1; ++k1)
          vi += values[j * bs0 * bs1 + k0 * bs1 + k1] * x[(int64_t)indices[j] * bs1 + k1];
      y[i * bs0 + k0] += vi;
    }
  }
}

/* spmvT: la/matrix_csr_impl.h:319-343 (y += A^T x) */
void orc_spmvT(const double* values, const int64_t* row_begin, const int64_t* row_end, const int32_t* indices,
               const double* x, double* y, int bs0, int bs1, int64_t nrows)
{
  for (int k0 = 0; k0 < bs0; ++k0)
    for (int64_t i = 0; i < nrows; i++)
    {
      const double xval = x[i * bs0 + k0];
      for (int64_t j = row_begin[i]; j < row_end[i]; j++)
        for (int k1 = 0; k1 < bs1; ++k1)
          y[(int64_t)indices[j] * bs1 + k1] += values[j * bs0 * bs1 + k0 * bs1 + k1] * xval;
    }
}

/* ------------------------------------------------------------------------- */
/* Element kernels (restating FFCx output for the benchmark forms)            */
/* ------------------------------------------------------------------------- */

/* numba kernel of python/test/unit/fem/test_custom_jit_kernels.py:29-46 (P1 triangle Laplace) */
static void k_laplace_p1_tri_A(double* A, const double* w, const double* c, const double* xc, const int* e,
                               const uint8_t* q, void* d)
{
  (void)w; (void)c; (void)e; (void)q; (void)d;
  double x0 = xc[0], y0 = xc[1], x1 = xc[3], y1 = xc[4], x2 = xc[6], y2 = xc[7];
  double Ae = fabs((x0 - x1) * (y2 - y1) - (y0 - y1) * (x2 - x1));
  double B[2][3] = {{y1 - y2, y2 - y0, y0 - y1}, {x2 - x1, x0 - x2, x1 - x0}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      A[3 * i + j] = (B[0][i] * B[0][j] + B[1][i] * B[1][j]) / (2 * Ae);
}

/* test_custom_jit_kernels.py:49-62: b[:] = Ae / 6 */
static void k_source_p1_tri_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                              const uint8_t* q, void* d)
{
  (void)w; (void)c; (void)e; (void)q; (void)d;
  double x0 = xc[0], y0 = xc[1], x1 = xc[3], y1 = xc[4], x2 = xc[6], y2 = xc[7];
  double Ae = fabs((x0 - x1) * (y2 - y1) - (y0 - y1) * (x2 - x1));
  for (int i = 0; i < 3; ++i)
    b[i] = Ae / 6.0;
}

/* 3-point degree-2 rule on the reference triangle (edge midpoints), weights 1/6 */
static const double TRI_Q[3][2] = {{0.5, 0.0}, {0.5, 0.5}, {0.0, 0.5}};

static double tri_detJ(const double* xc)
{
  double J00 = xc[3] - xc[0], J01 = xc[6] - xc[0];
  double J10 = xc[4] - xc[1], J11 = xc[7] - xc[1];
  return J00 * J11 - J01 * J10;
}

/* a = f u v dx, f in P1 (w[0..2]) — degree-3 integrand: 6-point degree-4 rule (Dunavant) */
static const double TRI_Q6[6][3] = {
    {0.445948490915965, 0.445948490915965, 0.223381589678011}, {0.445948490915965, 0.108103018168070, 0.223381589678011},
    {0.108103018168070, 0.445948490915965, 0.223381589678011}, {0.091576213509771, 0.091576213509771, 0.109951743655322},
    {0.091576213509771, 0.816847572980459, 0.109951743655322}, {0.816847572980459, 0.091576213509771, 0.109951743655322}};

static void k_mass_coeff_p1_tri_A(double* A, const double* w, const double* c, const double* xc, const int* e,
                                  const uint8_t* q, void* d)
{
  (void)c; (void)e; (void)q; (void)d;
  double det = fabs(tri_detJ(xc));
  for (int p = 0; p < 6; ++p)
  {
    double phi[3] = {1.0 - TRI_Q6[p][0] - TRI_Q6[p][1], TRI_Q6[p][0], TRI_Q6[p][1]};
    double f = w[0] * phi[0] + w[1] * phi[1] + w[2] * phi[2];
    double wt = 0.5 * TRI_Q6[p][2] * det;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        A[3 * i + j] += wt * f * phi[i] * phi[j];
  }
}

/* L = f v dx, f in P1 (w[0..2]) */
static void k_load_coeff_p1_tri_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                                  const uint8_t* q, void* d)
{
  (void)c; (void)e; (void)q; (void)d;
  double det = fabs(tri_detJ(xc));
  for (int p = 0; p < 3; ++p)
  {
    double phi[3] = {1.0 - TRI_Q[p][0] - TRI_Q[p][1], TRI_Q[p][0], TRI_Q[p][1]};
    double f = w[0] * phi[0] + w[1] * phi[1] + w[2] * phi[2];
    for (int i = 0; i < 3; ++i)
      b[i] += det / 6.0 * f * phi[i];
  }
}

/* Triangle facet i is opposite vertex i (Basix). 2-point Gauss on the edge. */
static void tri_facet(const double* xc, int lf, int* a, int* b, double* len)
{
  static const int FV[3][2] = {{1, 2}, {0, 2}, {0, 1}};
  *a = FV[lf][0];
  *b = FV[lf][1];
  double dx = xc[3 * *b] - xc[3 * *a], dy = xc[3 * *b + 1] - xc[3 * *a + 1], dz = xc[3 * *b + 2] - xc[3 * *a + 2];
  *len = sqrt(dx * dx + dy * dy + dz * dz);
}

/* a = u v ds on local facet entity_local_index[0] */
static void k_facet_mass_p1_tri_A(double* A, const double* w, const double* c, const double* xc, const int* e,
                                  const uint8_t* q, void* d)
{
  (void)w; (void)c; (void)q; (void)d;
  int a, b;
  double len;
  tri_facet(xc, e[0], &a, &b, &len);
  const double g[2] = {0.5 - 0.5 / sqrt(3.0), 0.5 + 0.5 / sqrt(3.0)};
  for (int p = 0; p < 2; ++p)
  {
    double phi[3] = {0, 0, 0};
    phi[a] = 1.0 - g[p];
    phi[b] = g[p];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        A[3 * i + j] += 0.5 * len * phi[i] * phi[j];
  }
}

/* L = c0 v ds (constant c[0]) on local facet */
static void k_facet_const_p1_tri_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                                   const uint8_t* q, void* d)
{
  (void)w; (void)q; (void)d;
  int a, bb;
  double len;
  tri_facet(xc, e[0], &a, &bb, &len);
  b[a] += 0.5 * len * c[0];
  b[bb] += 0.5 * len * c[0];
}

/* a = inner(avg(u), avg(v)) dS, P1 triangles (python/test/unit/fem/test_ghost_mesh_assembly.py:104-122).
 * Macro element of an interior facet: coordinate_dofs = [cell0 (3 x 3), cell1 (3 x 3)], A is (6 x 6) with the
 * 2 x 2 block layout cell0cell0 | cell0cell1 / cell1cell0 | cell1cell1 (fem/assemble_matrix_impl.h:581-587),
 * e[0], e[1] = local facet of the edge in cell0 / cell1.  avg(w) = (w+ + w-)/2; on the edge only the basis
 * functions of the two edge vertices are non-zero, and the trace of a P1 basis function is the 1-D hat:
 * int_edge phi_a phi_b = len (1 + delta_ab) / 6.  The shared vertices are matched through their coordinates. */
static void k_avg_mass_p1_tri_dS(double* A, const double* w, const double* c, const double* xc, const int* e,
                                 const uint8_t* q, void* d)
{
  (void)w; (void)c; (void)q; (void)d;
  int v[2][2];
  double len = 0.0;
  for (int s = 0; s < 2; ++s)
  {
    double l;
    tri_facet(xc + 9 * s, e[s], &v[s][0], &v[s][1], &l);
    len = l;
  }
  /* orient cell1's edge like cell0's: v[1][k] must be the same point as v[0][k] */
  const double* p0 = xc + 3 * v[0][0];
  const double* q0 = xc + 9 + 3 * v[1][0];
  if (fabs(p0[0] - q0[0]) + fabs(p0[1] - q0[1]) + fabs(p0[2] - q0[2]) > 1e-12 * (1.0 + len))
  {
    int t = v[1][0];
    v[1][0] = v[1][1];
    v[1][1] = t;
  }
  for (int s = 0; s < 2; ++s)
    for (int t = 0; t < 2; ++t)
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
          A[6 * (3 * s + v[s][a]) + 3 * t + v[t][b]] += 0.25 * len * (a == b ? 2.0 : 1.0) / 6.0;
}

/* Vertices of the shared edge in the two cells of an interior facet, cell1's pair ordered like cell0's (matched
 * through the coordinates); returns the edge length. */
static double tri_dS_vertices(const double* xc, const int* e, int v[2][2])
{
  double len = 0.0;
  for (int s = 0; s < 2; ++s)
  {
    double l;
    tri_facet(xc + 9 * s, e[s], &v[s][0], &v[s][1], &l);
    len = l;
  }
  const double* p0 = xc + 3 * v[0][0];
  const double* q0 = xc + 9 + 3 * v[1][0];
  if (fabs(p0[0] - q0[0]) + fabs(p0[1] - q0[1]) + fabs(p0[2] - q0[2]) > 1e-12 * (1.0 + len))
  {
    int t = v[1][0];
    v[1][0] = v[1][1];
    v[1][1] = t;
  }
  return len;
}

/* L = conj(avg(v)) dS, P1 triangles (python/test/unit/fem/test_assembler.py:1003): macro element vector of 6 entries
 * [cell0 | cell1]; int_edge phi_a = len / 2 for the two edge vertices of each cell, avg halves it. */
static void k_avg_load_p1_tri_dS_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                                   const uint8_t* q, void* d)
{
  (void)w; (void)c; (void)q; (void)d;
  int v[2][2];
  const double len = tri_dS_vertices(xc, e, v);
  for (int s = 0; s < 2; ++s)
    for (int a = 0; a < 2; ++a)
      b[3 * s + v[s][a]] += 0.5 * (0.5 * len);
}

/* M = 1 dS: the length of the interior facet (python/test/unit/fem/test_assemble_domains.py:203-210) */
static void k_one_tri_dS_M(double* A, const double* w, const double* c, const double* xc, const int* e,
                           const uint8_t* q, void* d)
{
  (void)w; (void)c; (void)q; (void)d;
  int a, b;
  double len;
  tri_facet(xc, e[0], &a, &b, &len);
  A[0] += len;
}

/* two-point Gauss rule on [0, 1] */
static const double GAUSS2[2] = {0.21132486540518713, 0.7886751345948129};

/* M = inner(avg(f), avg(f)) dS, f in P1 (python/test/unit/fem/test_assemble_domains.py:225): w = [f on cell0 (3),
 * f on cell1 (3)] (the reference packs the two restrictions one after the other, fem/pack.h:196-226) */
static void k_avg2_coeff_p1_tri_dS_M(double* A, const double* w, const double* c, const double* xc, const int* e,
                                     const uint8_t* q, void* d)
{
  (void)c; (void)q; (void)d;
  int v[2][2];
  const double len = tri_dS_vertices(xc, e, v);
  for (int g = 0; g < 2; ++g)
  {
    const double t = GAUSS2[g];
    const double h = 0.5 * ((1.0 - t) * w[v[0][0]] + t * w[v[0][1]]) + 0.5 * ((1.0 - t) * w[3 + v[1][0]] + t * w[3 + v[1][1]]);
    A[0] += 0.5 * len * h * h;
  }
}

/* M = inner(f, f) ds on an exterior facet, f in P1 (python/test/unit/fem/test_assemble_domains.py:224) */
static void k_coeff2_p1_tri_ds_M(double* A, const double* w, const double* c, const double* xc, const int* e,
                                 const uint8_t* q, void* d)
{
  (void)c; (void)q; (void)d;
  int a, b;
  double len;
  tri_facet(xc, e[0], &a, &b, &len);
  for (int g = 0; g < 2; ++g)
  {
    const double t = GAUSS2[g];
    const double h = (1.0 - t) * w[a] + t * w[b];
    A[0] += 0.5 * len * h * h;
  }
}

/* --- tetrahedra ---------------------------------------------------------- */

/* J = [x1-x0, x2-x0, x3-x0] (columns), as FFCx builds it (test_custom_jit_kernels.py:173-182) */
static double tet_geometry(const double* xc, double K[3][3])
{
  double J[3][3];
  for (int i = 0; i < 3; ++i)
    for (int a = 0; a < 3; ++a)
      J[i][a] = xc[3 * (a + 1) + i] - xc[i];
  double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
               + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  K[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
  K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
  K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
  K[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
  K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
  K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  K[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
  K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
  K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  return det;
}

/* a = kappa grad u . grad v dx, P1 tets, kappa = c[0] (cpp/demo/poisson/poisson.py form, 3-D) */
static void k_poisson_p1_tet_A(double* A, const double* w, const double* c, const double* xc, const int* e,
                               const uint8_t* q, void* d)
{
  (void)w; (void)e; (void)q; (void)d;
  double K[3][3];
  double det = tet_geometry(xc, K);
  static const double dphi[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double g[4][3]; /* physical gradients: g_i = K^T dphi_i */
  for (int i = 0; i < 4; ++i)
    for (int m = 0; m < 3; ++m)
      g[i][m] = dphi[i][0] * K[0][m] + dphi[i][1] * K[1][m] + dphi[i][2] * K[2][m];
  double scale = c[0] * fabs(det) / 6.0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      A[4 * i + j] += scale * (g[i][0] * g[j][0] + g[i][1] * g[j][1] + g[i][2] * g[j][2]);
}

/* Collapsed (Duffy) Gauss-Legendre rule on the reference tetrahedron, n points per direction.
 * Exact for total degree 2n-3.  Returns number of points; pts[k] = (x,y,z,weight). */
static int tet_rule(int n, double pts[][4])
{
  static const double G2[2] = {-0.5773502691896257, 0.5773502691896257}, W2[2] = {1.0, 1.0};
  static const double G3[3] = {-0.7745966692414834, 0.0, 0.7745966692414834};
  static const double W3[3] = {0.5555555555555556, 0.8888888888888888, 0.5555555555555556};
  static const double G4[4] = {-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526};
  static const double W4[4] = {0.3478548451374538, 0.6521451548625461, 0.6521451548625461, 0.3478548451374538};
  const double* G = n == 2 ? G2 : (n == 3 ? G3 : G4);
  const double* W = n == 2 ? W2 : (n == 3 ? W3 : W4);
  int k = 0;
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b)
      for (int cc = 0; cc < n; ++cc)
      {
        double u = 0.5 * (G[a] + 1), v = 0.5 * (G[b] + 1), t = 0.5 * (G[cc] + 1);
        pts[k][0] = u;
        pts[k][1] = v * (1 - u);
        pts[k][2] = t * (1 - u) * (1 - v);
        pts[k][3] = 0.125 * W[a] * W[b] * W[cc] * (1 - u) * (1 - u) * (1 - v);
        ++k;
      }
  return k;
}

/* L = f v dx, P1 tets, f in P1 (w[0..3]) */
static void k_load_p1_tet_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                            const uint8_t* q, void* d)
{
  (void)c; (void)e; (void)q; (void)d;
  double K[3][3];
  double det = fabs(tet_geometry(xc, K));
  double pts[64][4];
  int np = tet_rule(3, pts);
  for (int p = 0; p < np; ++p)
  {
    double phi[4] = {1 - pts[p][0] - pts[p][1] - pts[p][2], pts[p][0], pts[p][1], pts[p][2]};
    double f = 0;
    for (int j = 0; j < 4; ++j)
      f += w[j] * phi[j];
    for (int i = 0; i < 4; ++i)
      b[i] += pts[p][3] * det * f * phi[i];
  }
}

/* L = f g v dx, P1 tets, f and g in P1: TWO coefficients in one integral, w = [f (4), g (4)] at the form's
 * coefficient offsets (fem/Form.h:593-604, fem/pack.h:265-330) */
static void k_load_prod_p1_tet_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                                 const uint8_t* q, void* d)
{
  (void)c; (void)e; (void)q; (void)d;
  double K[3][3];
  double det = fabs(tet_geometry(xc, K));
  double pts[64][4];
  int np = tet_rule(3, pts);
  for (int p = 0; p < np; ++p)
  {
    double phi[4] = {1 - pts[p][0] - pts[p][1] - pts[p][2], pts[p][0], pts[p][1], pts[p][2]};
    double f = 0, g = 0;
    for (int j = 0; j < 4; ++j)
    {
      f += w[j] * phi[j];
      g += w[4 + j] * phi[j];
    }
    for (int i = 0; i < 4; ++i)
      b[i] += pts[p][3] * det * f * g * phi[i];
  }
}

/* Forms the product does NOT ship: the checker side of the kernel plug point test (tests/cpp/user_kernel_plugin.cu
 * registers the device versions through bfx_register_kernel).  a = inner(u, v) dx, L = c0 v dx, M = 1 dx on P1 tets. */
static void k_mass_p1_tet_A(double* A, const double* w, const double* c, const double* xc, const int* e,
                            const uint8_t* q, void* d)
{
  (void)w; (void)c; (void)e; (void)q; (void)d;
  double K[3][3];
  double det = fabs(tet_geometry(xc, K));
  double pts[64][4];
  int np = tet_rule(3, pts);
  for (int p = 0; p < np; ++p)
  {
    double phi[4] = {1 - pts[p][0] - pts[p][1] - pts[p][2], pts[p][0], pts[p][1], pts[p][2]};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
        A[4 * i + j] += pts[p][3] * det * phi[i] * phi[j];
  }
}

static void k_source_const_p1_tet_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                                    const uint8_t* q, void* d)
{
  (void)w; (void)e; (void)q; (void)d;
  double K[3][3];
  double det = fabs(tet_geometry(xc, K));
  double pts[64][4];
  int np = tet_rule(2, pts);
  for (int p = 0; p < np; ++p)
  {
    double phi[4] = {1 - pts[p][0] - pts[p][1] - pts[p][2], pts[p][0], pts[p][1], pts[p][2]};
    for (int i = 0; i < 4; ++i)
      b[i] += pts[p][3] * det * c[0] * phi[i];
  }
}

static void k_volume_tet_M(double* A, const double* w, const double* c, const double* xc, const int* e, const uint8_t* q,
                           void* d)
{
  (void)w; (void)c; (void)e; (void)q; (void)d;
  double K[3][3];
  A[0] += fabs(tet_geometry(xc, K)) / 6.0;
}

/* P2 basis on the reference tetrahedron: 4 vertex functions then 6 edge functions in Basix edge order */
static const int TET_E[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};

static void p2_tet_basis(const double* X, double phi[10], double dphi[10][3])
{
  double lam[4] = {1 - X[0] - X[1] - X[2], X[0], X[1], X[2]};
  static const double dl[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int v = 0; v < 4; ++v)
  {
    phi[v] = lam[v] * (2 * lam[v] - 1);
    for (int m = 0; m < 3; ++m)
      dphi[v][m] = (4 * lam[v] - 1) * dl[v][m];
  }
  for (int k = 0; k < 6; ++k)
  {
    int a = TET_E[k][0], b = TET_E[k][1];
    phi[4 + k] = 4 * lam[a] * lam[b];
    for (int m = 0; m < 3; ++m)
      dphi[4 + k][m] = 4 * (lam[a] * dl[b][m] + lam[b] * dl[a][m]);
  }
}

/* a = kappa grad u . grad v dx, P2 tets (cpp/test/poisson.py:16-27), by quadrature */
static void k_poisson_p2_tet_A(double* A, const double* w, const double* c, const double* xc, const int* e,
                               const uint8_t* q, void* d)
{
  (void)w; (void)e; (void)q; (void)d;
  double K[3][3];
  double det = fabs(tet_geometry(xc, K));
  double pts[64][4];
  int np = tet_rule(3, pts);
  for (int p = 0; p < np; ++p)
  {
    double phi[10], dphi[10][3], g[10][3];
    p2_tet_basis(pts[p], phi, dphi);
    for (int i = 0; i < 10; ++i)
      for (int m = 0; m < 3; ++m)
        g[i][m] = dphi[i][0] * K[0][m] + dphi[i][1] * K[1][m] + dphi[i][2] * K[2][m];
    double wt = c[0] * pts[p][3] * det;
    for (int i = 0; i < 10; ++i)
      for (int j = 0; j < 10; ++j)
        A[10 * i + j] += wt * (g[i][0] * g[j][0] + g[i][1] * g[j][1] + g[i][2] * g[j][2]);
  }
}

/* L = f v dx, P2 tets, f in P2 (w[0..9]) */
static void k_load_p2_tet_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                            const uint8_t* q, void* d)
{
  (void)c; (void)e; (void)q; (void)d;
  double K[3][3];
  double det = fabs(tet_geometry(xc, K));
  double pts[64][4];
  int np = tet_rule(4, pts);
  for (int p = 0; p < np; ++p)
  {
    double phi[10], dphi[10][3];
    p2_tet_basis(pts[p], phi, dphi);
    double f = 0;
    for (int j = 0; j < 10; ++j)
      f += w[j] * phi[j];
    for (int i = 0; i < 10; ++i)
      b[i] += pts[p][3] * det * f * phi[i];
  }
}

/* Tet facet i is opposite vertex i.  L = g v ds with g in P1 (w[0..3]), degree-2 rule on the facet */
static void k_facet_load_p1_tet_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                                  const uint8_t* q, void* d)
{
  (void)c; (void)q; (void)d;
  static const int FV[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
  const int* fv = FV[e[0]];
  double u[3], v[3];
  for (int m = 0; m < 3; ++m)
  {
    u[m] = xc[3 * fv[1] + m] - xc[3 * fv[0] + m];
    v[m] = xc[3 * fv[2] + m] - xc[3 * fv[0] + m];
  }
  double cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
  double area2 = sqrt(cx * cx + cy * cy + cz * cz); /* 2 * area */
  for (int p = 0; p < 3; ++p)
  {
    double l[3] = {1.0 - TRI_Q[p][0] - TRI_Q[p][1], TRI_Q[p][0], TRI_Q[p][1]};
    double g = w[fv[0]] * l[0] + w[fv[1]] * l[1] + w[fv[2]] * l[2];
    for (int i = 0; i < 3; ++i)
      b[fv[i]] += area2 / 6.0 * g * l[i];
  }
}

/* a = u v ds on a tet facet (used for Robin-type terms; same structure as the 2-D golden form) */
static void k_facet_mass_p1_tet_A(double* A, const double* w, const double* c, const double* xc, const int* e,
                                  const uint8_t* q, void* d)
{
  (void)w; (void)c; (void)q; (void)d;
  static const int FV[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
  const int* fv = FV[e[0]];
  double u[3], v[3];
  for (int m = 0; m < 3; ++m)
  {
    u[m] = xc[3 * fv[1] + m] - xc[3 * fv[0] + m];
    v[m] = xc[3 * fv[2] + m] - xc[3 * fv[0] + m];
  }
  double cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
  double area2 = sqrt(cx * cx + cy * cy + cz * cz);
  for (int p = 0; p < 3; ++p)
  {
    double l[3] = {1.0 - TRI_Q[p][0] - TRI_Q[p][1], TRI_Q[p][0], TRI_Q[p][1]};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        A[4 * fv[i] + fv[j]] += area2 / 6.0 * l[i] * l[j];
  }
}

/* action(a, ui) of the Poisson forms (cpp/demo/poisson_matrix_free/poisson.py: M = action(a, ui)): the element
 * matrix of `a` applied to the coefficient dofs w, be_i = sum_j A_ij w_j */
static void k_action_poisson_p1_tet_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                                      const uint8_t* q, void* d)
{
  double A[16] = {0};
  k_poisson_p1_tet_A(A, 0, c, xc, e, q, d);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      b[i] += A[4 * i + j] * w[j];
}

static void k_action_poisson_p2_tet_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                                      const uint8_t* q, void* d)
{
  double A[100] = {0};
  k_poisson_p2_tet_A(A, 0, c, xc, e, q, d);
  for (int i = 0; i < 10; ++i)
    for (int j = 0; j < 10; ++j)
      b[i] += A[10 * i + j] * w[j];
}

/* M = w^2 dx, w in P1 (the error functional E = (usol - uexact)^2 dx of poisson_matrix_free/poisson.py with
 * w = usol - uexact), by a degree-3 rule */
static void k_l2norm2_p1_tet_M(double* m, const double* w, const double* c, const double* xc, const int* e,
                               const uint8_t* q, void* d)
{
  (void)c; (void)e; (void)q; (void)d;
  double K[3][3];
  double det = fabs(tet_geometry(xc, K));
  double pts[64][4];
  int np = tet_rule(3, pts);
  for (int p = 0; p < np; ++p)
  {
    const double X = pts[p][0], Y = pts[p][1], Z = pts[p][2];
    const double v = w[0] * (1.0 - X - Y - Z) + w[1] * X + w[2] * Y + w[3] * Z;
    m[0] += pts[p][3] * det * v * v;
  }
}

/* --- hexahedra (Q1, tensor node order x fastest) --------------------------- */

static void q1_hex_basis(const double* X, double phi[8], double dphi[8][3])
{
  for (int n = 0; n < 8; ++n)
  {
    int bx = n & 1, by = (n >> 1) & 1, bz = (n >> 2) & 1;
    double fx = bx ? X[0] : 1 - X[0], fy = by ? X[1] : 1 - X[1], fz = bz ? X[2] : 1 - X[2];
    double dx = bx ? 1.0 : -1.0, dy = by ? 1.0 : -1.0, dz = bz ? 1.0 : -1.0;
    phi[n] = fx * fy * fz;
    dphi[n][0] = dx * fy * fz;
    dphi[n][1] = fx * dy * fz;
    dphi[n][2] = fx * fy * dz;
  }
}

/* physical gradients + |det J| at reference point X of a trilinear hexahedron */
static double q1_hex_geometry(const double* xc, const double* X, double phi[8], double g[8][3])
{
  double dphi[8][3], J[3][3] = {{0}}, K[3][3];
  q1_hex_basis(X, phi, dphi);
  for (int n = 0; n < 8; ++n)
    for (int i = 0; i < 3; ++i)
      for (int a = 0; a < 3; ++a)
        J[i][a] += xc[3 * n + i] * dphi[n][a];
  double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
               + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  K[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
  K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
  K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
  K[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
  K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
  K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  K[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
  K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
  K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  for (int n = 0; n < 8; ++n)
    for (int m = 0; m < 3; ++m)
      g[n][m] = dphi[n][0] * K[0][m] + dphi[n][1] * K[1][m] + dphi[n][2] * K[2][m];
  return fabs(det);
}

/* a = inner(sigma(u), grad(v)) dx, sigma = 2 mu sym(grad u) + lambda tr(sym grad u) I
 * (python/demo/demo_elasticity.py:131-150); constants c = {mu, lambda}; Q1 vector, bs = 3,
 * element dof index = 3*node + component (assemble_matrix_impl.h:168-173). */
static const double GL3[3] = {0.1127016653792583, 0.5, 0.8872983346207417};
static const double GW3[3] = {0.2777777777777778, 0.4444444444444444, 0.2777777777777778};
static const double GL2[2] = {0.21132486540518713, 0.7886751345948129};
static const double GW2[2] = {0.5, 0.5};

static void elasticity_q1_hex(double* A, const double* c, const double* xc, int npts, const double* gl,
                              const double* gw)
{
  const double mu = c[0], lmbda = c[1];
  for (int a = 0; a < npts; ++a)
    for (int b = 0; b < npts; ++b)
      for (int cc = 0; cc < npts; ++cc)
      {
        double X[3] = {gl[a], gl[b], gl[cc]}, phi[8], g[8][3];
        double wt = gw[a] * gw[b] * gw[cc] * q1_hex_geometry(xc, X, phi, g);
        for (int i = 0; i < 8; ++i)
          for (int k = 0; k < 3; ++k)
            for (int j = 0; j < 8; ++j)
              for (int l = 0; l < 3; ++l)
              {
                /* sigma(phi_j e_l)_{kn} grad(phi_i e_k)_{kn} summed over n */
                double v = lmbda * g[j][l] * g[i][k] + mu * g[j][k] * g[i][l];
                if (k == l)
                  v += mu * (g[j][0] * g[i][0] + g[j][1] * g[i][1] + g[j][2] * g[i][2]);
                A[(3 * i + k) * 24 + (3 * j + l)] += wt * v;
              }
      }
}

static void k_elasticity_q1_hex_A(double* A, const double* w, const double* c, const double* xc, const int* e,
                                  const uint8_t* q, void* d)
{
  (void)w; (void)e; (void)q; (void)d;
  elasticity_q1_hex(A, c, xc, 3, GL3, GW3);
}

/* same form with the 2x2x2 Gauss rule (exact on parallelepipeds; the rule the CUDA kernel uses on
 * general trilinear cells) */
static void k_elasticity_q1_hex_A_g2(double* A, const double* w, const double* c, const double* xc, const int* e,
                                     const uint8_t* q, void* d)
{
  (void)w; (void)e; (void)q; (void)d;
  elasticity_q1_hex(A, c, xc, 2, GL2, GW2);
}

/* L = f . v dx, f a Q1 vector coefficient (w[3*node + comp]) */
static void k_load_q1_hex_L(double* b, const double* w, const double* c, const double* xc, const int* e,
                            const uint8_t* q, void* d)
{
  (void)c; (void)e; (void)q; (void)d;
  for (int a = 0; a < 3; ++a)
    for (int bb = 0; bb < 3; ++bb)
      for (int cc = 0; cc < 3; ++cc)
      {
        double X[3] = {GL3[a], GL3[bb], GL3[cc]}, phi[8], g[8][3];
        double wt = GW3[a] * GW3[bb] * GW3[cc] * q1_hex_geometry(xc, X, phi, g);
        for (int k = 0; k < 3; ++k)
        {
          double f = 0;
          for (int j = 0; j < 8; ++j)
            f += w[3 * j + k] * phi[j];
          for (int i = 0; i < 8; ++i)
            b[3 * i + k] += wt * f * phi[i];
        }
      }
}

/* Kernel ids shared with include/bfx.h (BFX_K_*) */
enum
{
  K_LAPLACE_P1_TRI_A = 0,
  K_SOURCE_P1_TRI_L = 1,
  K_MASS_COEFF_P1_TRI_A = 2,
  K_LOAD_COEFF_P1_TRI_L = 3,
  K_FACET_MASS_P1_TRI_A = 4,
  K_FACET_CONST_P1_TRI_L = 5,
  K_POISSON_P1_TET_A = 6,
  K_LOAD_P1_TET_L = 7,
  K_POISSON_P2_TET_A = 8,
  K_LOAD_P2_TET_L = 9,
  K_ELASTICITY_Q1_HEX_A = 10,
  K_LOAD_Q1_HEX_L = 11,
  K_FACET_LOAD_P1_TET_L = 12,
  K_FACET_MASS_P1_TET_A = 13,
  K_ELASTICITY_Q1_HEX_A_G2 = 14, /* oracle-only variant: 2x2x2 Gauss */
  K_ACTION_POISSON_P1_TET_L = 15,
  K_ACTION_POISSON_P2_TET_L = 16,
  K_L2NORM2_P1_TET_M = 17,
  K_AVG_MASS_P1_TRI_DS = 18,
  K_AVG_LOAD_P1_TRI_DS_L = 19,
  K_ONE_TRI_DS_M = 20,
  K_AVG2_COEFF_P1_TRI_DS_M = 21,
  K_COEFF2_P1_TRI_FACET_M = 22,
  K_LOAD_PROD_P1_TET_L = 23,
  K_MASS_P1_TET_A = 24,         /* oracle-only: the plug point test's forms */
  K_SOURCE_CONST_P1_TET_L = 25,
  K_VOLUME_TET_M = 26,
  K_COUNT
};

static orc_kernel_t kernel_table(int id)
{
  switch (id)
  {
  case K_LAPLACE_P1_TRI_A: return k_laplace_p1_tri_A;
  case K_SOURCE_P1_TRI_L: return k_source_p1_tri_L;
  case K_MASS_COEFF_P1_TRI_A: return k_mass_coeff_p1_tri_A;
  case K_LOAD_COEFF_P1_TRI_L: return k_load_coeff_p1_tri_L;
  case K_FACET_MASS_P1_TRI_A: return k_facet_mass_p1_tri_A;
  case K_FACET_CONST_P1_TRI_L: return k_facet_const_p1_tri_L;
  case K_POISSON_P1_TET_A: return k_poisson_p1_tet_A;
  case K_LOAD_P1_TET_L: return k_load_p1_tet_L;
  case K_POISSON_P2_TET_A: return k_poisson_p2_tet_A;
  case K_LOAD_P2_TET_L: return k_load_p2_tet_L;
  case K_ELASTICITY_Q1_HEX_A: return k_elasticity_q1_hex_A;
  case K_LOAD_Q1_HEX_L: return k_load_q1_hex_L;
  case K_FACET_LOAD_P1_TET_L: return k_facet_load_p1_tet_L;
  case K_FACET_MASS_P1_TET_A: return k_facet_mass_p1_tet_A;
  case K_ELASTICITY_Q1_HEX_A_G2: return k_elasticity_q1_hex_A_g2;
  case K_ACTION_POISSON_P1_TET_L: return k_action_poisson_p1_tet_L;
  case K_ACTION_POISSON_P2_TET_L: return k_action_poisson_p2_tet_L;
  case K_L2NORM2_P1_TET_M: return k_l2norm2_p1_tet_M;
  case K_AVG_MASS_P1_TRI_DS: return k_avg_mass_p1_tri_dS;
  case K_AVG_LOAD_P1_TRI_DS_L: return k_avg_load_p1_tri_dS_L;
  case K_ONE_TRI_DS_M: return k_one_tri_dS_M;
  case K_AVG2_COEFF_P1_TRI_DS_M: return k_avg2_coeff_p1_tri_dS_M;
  case K_COEFF2_P1_TRI_FACET_M: return k_coeff2_p1_tri_ds_M;
  case K_LOAD_PROD_P1_TET_L: return k_load_prod_p1_tet_L;
  case K_MASS_P1_TET_A: return k_mass_p1_tet_A;
  case K_SOURCE_CONST_P1_TET_L: return k_source_const_p1_tet_L;
  case K_VOLUME_TET_M: return k_volume_tet_M;
  default: return 0;
  }
}

/* Tabulate one element tensor (exposed so tests can compare element matrices entry-wise) */
int orc_tabulate(int kernel_id, double* A, int nA, const double* w, const double* c, const double* xc,
                 int local_entity)
{
  orc_kernel_t k = kernel_table(kernel_id);
  if (!k)
    return -2;
  memset(A, 0, sizeof(double) * nA);
  uint8_t perm = 0;
  k(A, w, c, xc, &local_entity, &perm, 0);
  return 0;
}

/* The same for an interior-facet (macro cell) kernel: two local facet indices */
int orc_tabulate2(int kernel_id, double* A, int nA, const double* w, const double* c, const double* xc, int lf0, int lf1)
{
  orc_kernel_t k = kernel_table(kernel_id);
  if (!k)
    return -2;
  memset(A, 0, sizeof(double) * nA);
  const int lf[2] = {lf0, lf1};
  uint8_t perm[2] = {0, 0};
  k(A, w, c, xc, lf, perm, 0);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* fem/assemble_matrix_impl.h, fem/assemble_vector_impl.h                     */
/* ------------------------------------------------------------------------- */

/* has_bc: fem/assemble_matrix_impl.h:27-34 */
static int has_bc(const int32_t* dofs, int nd, const int8_t* bc, int bs)
{
  for (int i = 0; i < nd; ++i)
    for (int k = 0; k < bs; ++k)
      if (bc[bs * dofs[i] + k])
        return 1;
  return 0;
}

typedef struct
{
  /* mat_set target: CSR add (MatrixCSR::mat_add_values, la/MatrixCSR.h:157-174) */
  double* data;
  const int32_t* cols;
  const int64_t* row_ptr;
  int bs0, bs1;
  /* lifting target (lift_bc lambda, fem/assemble_vector_impl.h:377-402) */
  double* b;
  const double* bc_values1;
  const int8_t* bc_markers1;
  const double* x0; /* may be NULL (x0.empty()) */
  double alpha;
} orc_sink_t;

static int sink_mat_add(orc_sink_t* s, const int32_t* rows, int nr, const int32_t* cols, int nc, const double* Ae)
{
  return orc_insert_csr(s->data, s->cols, s->row_ptr, Ae, rows, nr, cols, nc, s->bs0, s->bs1, 1);
}

static int sink_lifting(orc_sink_t* s, const int32_t* rows, int nr, const int32_t* cols, int ncols, const double* Ae)
{
  const int bs0 = s->bs0, bs1 = s->bs1;
  const size_t nc = (size_t)ncols * bs1;
  for (int i = 0; i < ncols; ++i)
    for (int k = 0; k < bs1; ++k)
    {
      const int32_t ii = cols[i] * bs1 + k;
      if (s->bc_markers1[ii])
      {
        const double x_bc = s->bc_values1[ii];
        const double _x0 = s->x0 ? s->x0[ii] : 0;
        for (int j = 0; j < nr; ++j)
          for (int m = 0; m < bs0; ++m)
          {
            const int32_t jj = rows[j] * bs0 + m;
            s->b[jj] -= Ae[((size_t)j * bs0 + m) * nc + ((size_t)i * bs1 + k)] * s->alpha * (x_bc - _x0);
          }
      }
    }
  return 0;
}

/* assemble_cells_matrix<LiftingMode> (fem/assemble_matrix_impl.h:92-200) and, when
 * entities != NULL, assemble_entities<LiftingMode> (:264-379): entities are flat
 * (cell, local_entity) pairs and coeffs are per entity.  P0/P1T dof transformations
 * are the identity for the Lagrange elements on this path (FiniteElement.cpp:122-127). */
static int assemble_matrix_loop(int lifting_mode, orc_kernel_t kernel, orc_sink_t* sink, const int32_t* x_dofmap,
                                int nx, const double* x, const int32_t* cells, const int32_t* entities, int64_t n,
                                const int32_t* dmap0, int nd0, int bs0, const int32_t* dmap1, int nd1, int bs1,
                                const int8_t* bc0, const int8_t* bc1, const double* coeffs, int cstride,
                                const double* constants)
{
  const int ndim0 = bs0 * nd0, ndim1 = bs1 * nd1;
  double* Ae = (double*)malloc(sizeof(double) * ndim0 * ndim1);
  double* cdofs = (double*)malloc(sizeof(double) * 3 * nx);
  int err = 0;
  for (int64_t c = 0; c < n && !err; ++c)
  {
    int32_t cell = entities ? entities[2 * c] : cells[c];
    int local_entity = entities ? entities[2 * c + 1] : 0;
    const int32_t* dofs0 = dmap0 + (size_t)cell * nd0;
    const int32_t* dofs1 = dmap1 + (size_t)cell * nd1;
    if (lifting_mode && !has_bc(dofs1, nd1, bc1, bs1))
      continue;
    const int32_t* x_dofs = x_dofmap + (size_t)cell * nx;
    for (int i = 0; i < nx; ++i)
      memcpy(cdofs + 3 * i, x + 3 * (size_t)x_dofs[i], 3 * sizeof(double));
    memset(Ae, 0, sizeof(double) * ndim0 * ndim1);
    uint8_t perm = 0;
    kernel(Ae, coeffs ? coeffs + (size_t)c * cstride : 0, constants, cdofs, entities ? &local_entity : 0,
           entities ? &perm : 0, 0);
    if (!lifting_mode)
    {
      if (bc0)
        for (int i = 0; i < nd0; ++i)
          for (int k = 0; k < bs0; ++k)
            if (bc0[bs0 * dofs0[i] + k])
            {
              const int row = bs0 * i + k;
              for (int j = 0; j < ndim1; ++j)
                Ae[(size_t)ndim1 * row + j] = 0;
            }
      if (bc1)
        for (int j = 0; j < nd1; ++j)
          for (int k = 0; k < bs1; ++k)
            if (bc1[bs1 * dofs1[j] + k])
            {
              const int col = bs1 * j + k;
              for (int row = 0; row < ndim0; ++row)
                Ae[(size_t)row * ndim1 + col] = 0;
            }
      err = sink_mat_add(sink, dofs0, nd0, dofs1, nd1, Ae);
    }
    else
      err = sink_lifting(sink, dofs0, nd0, dofs1, nd1, Ae);
  }
  free(Ae);
  free(cdofs);
  return err;
}

/* fem::assemble_matrix cell/exterior-facet integral into MatrixCSR values (bs0 x bs1 blocks). */
int orc_assemble_matrix(int kernel_id, const int32_t* x_dofmap, int nx, const double* x, const int32_t* cells,
                        const int32_t* entities, int64_t n, const int32_t* dmap0, int nd0, int bs0,
                        const int32_t* dmap1, int nd1, int bs1, const int8_t* bc0, const int8_t* bc1,
                        const double* coeffs, int cstride, const double* constants, double* data,
                        const int32_t* cols, const int64_t* row_ptr)
{
  orc_kernel_t k = kernel_table(kernel_id);
  if (!k)
    return -2;
  orc_sink_t s = {data, cols, row_ptr, bs0, bs1, 0, 0, 0, 0, 0};
  return assemble_matrix_loop(0, k, &s, x_dofmap, nx, x, cells, entities, n, dmap0, nd0, bs0, dmap1, nd1, bs1, bc0,
                              bc1, coeffs, cstride, constants);
}

/* impl::lift_bc (fem/assemble_vector_impl.h:361-414): b -= alpha A (x_bc - x0) through
 * assemble_matrix<true>.  x0 may be NULL. */
int orc_lift_bc(int kernel_id, const int32_t* x_dofmap, int nx, const double* x, const int32_t* cells,
                const int32_t* entities, int64_t n, const int32_t* dmap0, int nd0, int bs0, const int32_t* dmap1,
                int nd1, int bs1, const double* coeffs, int cstride, const double* constants, double* b,
                const double* bc_values1, const int8_t* bc_markers1, const double* x0, double alpha)
{
  orc_kernel_t k = kernel_table(kernel_id);
  if (!k)
    return -2;
  orc_sink_t s = {0, 0, 0, bs0, bs1, b, bc_values1, bc_markers1, x0, alpha};
  return assemble_matrix_loop(1, k, &s, x_dofmap, nx, x, cells, entities, n, dmap0, nd0, bs0, dmap1, nd1, bs1, 0,
                              bc_markers1, coeffs, cstride, constants);
}

/* assemble_cells (fem/assemble_vector_impl.h:72-116) / assemble_entities (:157-215) */
int orc_assemble_vector(int kernel_id, const int32_t* x_dofmap, int nx, const double* x, const int32_t* cells,
                        const int32_t* entities, int64_t n, const int32_t* dmap, int nd, int bs, const double* coeffs,
                        int cstride, const double* constants, double* b)
{
  orc_kernel_t kernel = kernel_table(kernel_id);
  if (!kernel)
    return -2;
  double* be = (double*)malloc(sizeof(double) * bs * nd);
  double* cdofs = (double*)malloc(sizeof(double) * 3 * nx);
  for (int64_t index = 0; index < n; ++index)
  {
    int32_t c = entities ? entities[2 * index] : cells[index];
    int local_entity = entities ? entities[2 * index + 1] : 0;
    const int32_t* x_dofs = x_dofmap + (size_t)c * nx;
    for (int i = 0; i < nx; ++i)
      memcpy(cdofs + 3 * i, x + 3 * (size_t)x_dofs[i], 3 * sizeof(double));
    memset(be, 0, sizeof(double) * bs * nd);
    uint8_t perm = 0;
    kernel(be, coeffs ? coeffs + (size_t)index * cstride : 0, constants, cdofs, entities ? &local_entity : 0,
           entities ? &perm : 0, 0);
    const int32_t* dofs = dmap + (size_t)c * nd;
    for (int i = 0; i < nd; ++i)
      for (int k = 0; k < bs; ++k)
        b[(size_t)bs * dofs[i] + k] += be[bs * i + k];
  }
  free(be);
  free(cdofs);
  return 0;
}

/* impl::assemble_interior_facets (fem/assemble_matrix_impl.h:442-667), the common case of one mesh with a cell
 * on both sides of every facet: facets = (F, 2, 2) [[cell0, local0], [cell1, local1]]; joint dofmaps
 * [dofs(cell0), dofs(cell1)]; bc rows / columns zeroed (:613-645); mat_add of the whole joint block (:653-654). */
int orc_assemble_matrix_interior_facets(int kernel_id, const int32_t* x_dofmap, int nx, const double* x,
                                        const int32_t* facets, int64_t nf, const int32_t* dmap0, int nd0, int bs0,
                                        const int32_t* dmap1, int nd1, int bs1, const int8_t* bc0, const int8_t* bc1,
                                        const double* constants, double* data, const int32_t* cols,
                                        const int64_t* row_ptr)
{
  orc_kernel_t kernel = kernel_table(kernel_id);
  if (!kernel)
    return -2;
  const int nr = 2 * nd0 * bs0, nc = 2 * nd1 * bs1;
  double* Ae = (double*)malloc(sizeof(double) * nr * nc);
  double* cdofs = (double*)malloc(sizeof(double) * 3 * 2 * nx);
  int32_t* j0 = (int32_t*)malloc(sizeof(int32_t) * 2 * nd0);
  int32_t* j1 = (int32_t*)malloc(sizeof(int32_t) * 2 * nd1);
  int err = 0;
  for (int64_t f = 0; f < nf && !err; ++f)
  {
    const int32_t c[2] = {facets[4 * f], facets[4 * f + 2]};
    const int lf[2] = {facets[4 * f + 1], facets[4 * f + 3]};
    for (int s = 0; s < 2; ++s)
    {
      for (int i = 0; i < nx; ++i)
        memcpy(cdofs + 3 * (s * nx + i), x + 3 * (size_t)x_dofmap[(size_t)c[s] * nx + i], 3 * sizeof(double));
      memcpy(j0 + s * nd0, dmap0 + (size_t)c[s] * nd0, sizeof(int32_t) * nd0);
      memcpy(j1 + s * nd1, dmap1 + (size_t)c[s] * nd1, sizeof(int32_t) * nd1);
    }
    memset(Ae, 0, sizeof(double) * nr * nc);
    uint8_t perm[2] = {0, 0};
    kernel(Ae, 0, constants, cdofs, lf, perm, 0);
    if (bc0)
      for (int i = 0; i < 2 * nd0; ++i)
        for (int k = 0; k < bs0; ++k)
          if (bc0[(size_t)bs0 * j0[i] + k])
            memset(Ae + (size_t)nc * (bs0 * i + k), 0, sizeof(double) * nc);
    if (bc1)
      for (int j = 0; j < 2 * nd1; ++j)
        for (int k = 0; k < bs1; ++k)
          if (bc1[(size_t)bs1 * j1[j] + k])
            for (int m = 0; m < nr; ++m)
              Ae[(size_t)m * nc + bs1 * j + k] = 0;
    err = orc_insert_csr(data, cols, row_ptr, Ae, j0, 2 * nd0, j1, 2 * nd1, bs0, bs1, 1);
  }
  free(Ae);
  free(cdofs);
  free(j0);
  free(j1);
  return err;
}

/* impl::assemble_interior_facets of a linear form (fem/assemble_vector_impl.h:249-339): be = [cell0 | cell1], added
 * through the dofmaps of the two cells; coeffs is (nf, 2, cstride) (the two restrictions of every coefficient). */
int orc_assemble_vector_interior_facets(int kernel_id, const int32_t* x_dofmap, int nx, const double* x,
                                        const int32_t* facets, int64_t nf, const int32_t* dmap, int nd, int bs,
                                        const double* coeffs, int cstride, const double* constants, double* b)
{
  orc_kernel_t kernel = kernel_table(kernel_id);
  if (!kernel)
    return -2;
  double* be = (double*)malloc(sizeof(double) * 2 * bs * nd);
  double* cdofs = (double*)malloc(sizeof(double) * 3 * 2 * nx);
  for (int64_t f = 0; f < nf; ++f)
  {
    const int32_t c[2] = {facets[4 * f], facets[4 * f + 2]};
    const int lf[2] = {facets[4 * f + 1], facets[4 * f + 3]};
    for (int s = 0; s < 2; ++s)
      for (int i = 0; i < nx; ++i)
        memcpy(cdofs + 3 * (s * nx + i), x + 3 * (size_t)x_dofmap[(size_t)c[s] * nx + i], 3 * sizeof(double));
    memset(be, 0, sizeof(double) * 2 * bs * nd);
    uint8_t perm[2] = {0, 0};
    kernel(be, coeffs ? coeffs + (size_t)f * 2 * cstride : 0, constants, cdofs, lf, perm, 0);
    for (int s = 0; s < 2; ++s)
    {
      const int32_t* dofs = dmap + (size_t)c[s] * nd;
      for (int i = 0; i < nd; ++i)
        for (int k = 0; k < bs; ++k)
          b[(size_t)bs * dofs[i] + k] += be[bs * (s * nd + i) + k];
    }
  }
  free(be);
  free(cdofs);
  return 0;
}

/* impl::assemble_exterior_facets of a functional (fem/assemble_scalar_impl.h:78-113): entities = (cell, local facet) */
int orc_assemble_scalar_facets(int kernel_id, const int32_t* x_dofmap, int nx, const double* x, const int32_t* entities,
                               int64_t n, const double* coeffs, int cstride, const double* constants, double* value)
{
  orc_kernel_t kernel = kernel_table(kernel_id);
  if (!kernel)
    return -2;
  double* cdofs = (double*)malloc(sizeof(double) * 3 * nx);
  double v = 0.0;
  for (int64_t index = 0; index < n; ++index)
  {
    const int32_t* x_dofs = x_dofmap + (size_t)entities[2 * index] * nx;
    int local_entity = entities[2 * index + 1];
    for (int i = 0; i < nx; ++i)
      memcpy(cdofs + 3 * i, x + 3 * (size_t)x_dofs[i], 3 * sizeof(double));
    uint8_t perm = 0;
    kernel(&v, coeffs ? coeffs + (size_t)index * cstride : 0, constants, cdofs, &local_entity, &perm, 0);
  }
  free(cdofs);
  *value = v;
  return 0;
}

/* impl::assemble_interior_facets of a functional (fem/assemble_scalar_impl.h:122-168) */
int orc_assemble_scalar_interior_facets(int kernel_id, const int32_t* x_dofmap, int nx, const double* x,
                                        const int32_t* facets, int64_t nf, const double* coeffs, int cstride,
                                        const double* constants, double* value)
{
  orc_kernel_t kernel = kernel_table(kernel_id);
  if (!kernel)
    return -2;
  double* cdofs = (double*)malloc(sizeof(double) * 3 * 2 * nx);
  double v = 0.0;
  for (int64_t f = 0; f < nf; ++f)
  {
    const int32_t c[2] = {facets[4 * f], facets[4 * f + 2]};
    const int lf[2] = {facets[4 * f + 1], facets[4 * f + 3]};
    for (int s = 0; s < 2; ++s)
      for (int i = 0; i < nx; ++i)
        memcpy(cdofs + 3 * (s * nx + i), x + 3 * (size_t)x_dofmap[(size_t)c[s] * nx + i], 3 * sizeof(double));
    uint8_t perm[2] = {0, 0};
    kernel(&v, coeffs ? coeffs + (size_t)f * 2 * cstride : 0, constants, cdofs, lf, perm, 0);
  }
  free(cdofs);
  *value = v;
  return 0;
}

/* impl::assemble_cells of a functional (fem/assemble_scalar_impl.h:32-60): value += kernel(...) over the cells */
int orc_assemble_scalar(int kernel_id, const int32_t* x_dofmap, int nx, const double* x, const int32_t* cells, int64_t n,
                        const double* coeffs, int cstride, const double* constants, double* value)
{
  orc_kernel_t kernel = kernel_table(kernel_id);
  if (!kernel)
    return -2;
  double* cdofs = (double*)malloc(sizeof(double) * 3 * nx);
  double v = 0.0;
  for (int64_t index = 0; index < n; ++index)
  {
    const int32_t* x_dofs = x_dofmap + (size_t)cells[index] * nx;
    for (int i = 0; i < nx; ++i)
      memcpy(cdofs + 3 * i, x + 3 * (size_t)x_dofs[i], 3 * sizeof(double));
    kernel(&v, coeffs ? coeffs + (size_t)index * cstride : 0, constants, cdofs, 0, 0, 0);
  }
  free(cdofs);
  *value = v;
  return 0;
}

/* pack_coefficient_entity / pack_impl (fem/pack.h:77-176): coeffs[e, offset + bs*i + k] = v[bs*dofs[i] + k]
 * for cells (entities == NULL) or the cell of each (cell, local) entity (pack.h:356-380). */
void orc_pack_coefficient(double* coeffs, int cstride, int offset, const double* v, const int32_t* dofmap, int nd,
                          int bs, const int32_t* cells, const int32_t* entities, int64_t n)
{
  for (int64_t e = 0; e < n; ++e)
  {
    int32_t cell = entities ? entities[2 * e] : cells[e];
    if (cell < 0)
      continue;
    const int32_t* dofs = dofmap + (size_t)cell * nd;
    double* cc = coeffs + (size_t)e * cstride + offset;
    for (int i = 0; i < nd; ++i)
      for (int k = 0; k < bs; ++k)
        cc[bs * i + k] = v[(size_t)bs * dofs[i] + k];
  }
}

/* DirichletBC::mark_dofs (fem/DirichletBC.h:589-601) */
void orc_bc_mark(int8_t* markers, const int32_t* dofs0, int64_t n)
{
  for (int64_t i = 0; i < n; ++i)
    markers[dofs0[i]] = 1;
}

/* DirichletBC::set (fem/DirichletBC.h:495-578).  g_kind 0: Function values g[dofs_g[i]];
 * 1: Constant value g[dof % bs].  x0 may be NULL.  Skips dofs >= x_size (:506-511). */
void orc_bc_set(double* x, int32_t x_size, const int32_t* dofs0, const int32_t* dofs_g, int64_t n, const double* g,
                int g_kind, int bs, const double* x0, double alpha)
{
  for (int64_t i = 0; i < n; ++i)
  {
    if (dofs0[i] >= x_size)
      continue;
    double v;
    if (alpha == 0.0)
      v = 0;
    else
    {
      double gv = g_kind == 0 ? g[dofs_g ? dofs_g[i] : dofs0[i]] : g[dofs0[i] % bs];
      v = x0 ? alpha * (gv - x0[dofs0[i]]) : alpha * gv;
    }
    x[dofs0[i]] = v;
  }
}

/* fem::set_diagonal (fem/assembler.h:644-686) through MatrixCSR::mat_set_values():
 * rows are UNROLLED dofs; a blocked matrix goes through insert_nonblocked_csr (MatrixCSR.h:286-292). */
int orc_set_diagonal(double* data, const int32_t* cols, const int64_t* row_ptr, int bs0, int bs1,
                     const int32_t* rows, int64_t n, double diagonal)
{
  for (int64_t i = 0; i < n; ++i)
  {
    int err;
    if (bs0 == 1 && bs1 == 1)
      err = orc_insert_csr(data, cols, row_ptr, &diagonal, rows + i, 1, rows + i, 1, 1, 1, 0);
    else
      err = orc_insert_nonblocked_csr(data, cols, row_ptr, &diagonal, rows + i, 1, rows + i, 1, bs0, bs1, 0);
    if (err)
      return err;
  }
  return 0;
}
