// Hand-written sm_100a element kernels: the device-side replacement of the FFCx-generated
// tabulate_tensor functions (ABI: fem/kernel.h:18-20; A row-major (bs*nd x bs*nd), w = all form
// coefficients at coefficient_offsets(), c = flattened constants, coordinate_dofs = nx x 3).
//
// Every element is a struct with
//   NX, ND, BS          geometry nodes / dofs per cell / block size (element dof = BS*node + comp)
//   WSIZE, CSIZE        number of coefficient / constant scalars read
//   RANK, FACET         2 = bilinear, 1 = linear; FACET: uses entity_local_index[0]
//   Geo                 per-cell data kept in registers between prepare() and row()/vec()
//   prepare(g, xc, w, c, local_facet)
//   row(g, i, out)      scalar row i of the element matrix (RANK 2); i is a compile-time constant
//                       after unrolling, so everything stays in registers
//   vec(g, out)         element vector (RANK 1)
// All integrals are exact for affine cells (closed-form pre-integration, like FFCx's
// "preintegrated" blocks, python/test/unit/fem/test_custom_jit_kernels.py:207-216); the oracle
// evaluates the same forms by numerical quadrature.
#pragma once
#include "common.cuh"

namespace bfx
{
namespace el
{
#define BFX_DI __device__ __forceinline__

// ---- triangles ------------------------------------------------------------------------------
struct TriBase
{
  static constexpr int NX = 3, ND = 3, BS = 1;
  // 2 x signed area
  static BFX_DI double det2(const double (&xc)[3][3])
  {
    return (xc[1][0] - xc[0][0]) * (xc[2][1] - xc[0][1]) - (xc[2][0] - xc[0][0]) * (xc[1][1] - xc[0][1]);
  }
  // facet i is the edge opposite vertex i (Basix)
  static BFX_DI void facet(const double (&xc)[3][3], int lf, int& a, int& b, double& len)
  {
    a = lf == 0 ? 1 : 0;
    b = lf == 2 ? 1 : 2;
    const double dx = xc[b][0] - xc[a][0], dy = xc[b][1] - xc[a][1], dz = xc[b][2] - xc[a][2];
    len = sqrt(dx * dx + dy * dy + dz * dz);
  }
};

// test_custom_jit_kernels.py:29-46 — A = B^T B / (2 Ae)
struct LaplaceP1Tri : TriBase
{
  static constexpr int WSIZE = 0, WND = 0, WBS = 1, CSIZE = 0, RANK = 2;
  static constexpr bool FACET = false;
  struct Geo
  {
    double B[2][3];
    double inv;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[3][3], const double*, const double*, int)
  {
    const double x0 = xc[0][0], y0 = xc[0][1], x1 = xc[1][0], y1 = xc[1][1], x2 = xc[2][0], y2 = xc[2][1];
    const double Ae = fabs((x0 - x1) * (y2 - y1) - (y0 - y1) * (x2 - x1));
    g.B[0][0] = y1 - y2, g.B[0][1] = y2 - y0, g.B[0][2] = y0 - y1;
    g.B[1][0] = x2 - x1, g.B[1][1] = x0 - x2, g.B[1][2] = x1 - x0;
    g.inv = 1.0 / (2.0 * Ae);
  }
  static BFX_DI void row(const Geo& g, int i, double (&out)[3])
  {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      out[j] = (g.B[0][i] * g.B[0][j] + g.B[1][i] * g.B[1][j]) * g.inv;
  }
};

// test_custom_jit_kernels.py:49-62 — b[:] = Ae / 6
struct SourceP1Tri : TriBase
{
  static constexpr int WSIZE = 0, WND = 0, WBS = 1, CSIZE = 0, RANK = 1;
  static constexpr bool FACET = false;
  struct Geo
  {
    double v;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[3][3], const double*, const double*, int)
  {
    g.v = fabs(det2(xc)) / 6.0;
  }
  static BFX_DI void vec(const Geo& g, double (&out)[3]) { out[0] = out[1] = out[2] = g.v; }
};

// inner(f*u, v)*dx, f in P1: A_ij = |det| sum_k f_k int(phi_i phi_j phi_k),
// int = 1/20 (i=j=k), 1/60 (two equal), 1/120 (all different)
struct MassCoeffP1Tri : TriBase
{
  static constexpr int WSIZE = 3, WND = 3, WBS = 1, CSIZE = 0, RANK = 2;
  static constexpr bool FACET = false;
  struct Geo
  {
    double f[3], F, det;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[3][3], const double* w, const double*, int)
  {
    g.det = fabs(det2(xc));
    g.f[0] = w[0], g.f[1] = w[1], g.f[2] = w[2];
    g.F = w[0] + w[1] + w[2];
  }
  static BFX_DI void row(const Geo& g, int i, double (&out)[3])
  {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      out[j] = i == j ? g.det * (g.F * (1.0 / 60.0) + g.f[i] * (1.0 / 30.0))
                      : g.det * ((g.f[i] + g.f[j] + g.F) * (1.0 / 120.0));
  }
};

// inner(f, v)*dx, f in P1: b_i = |det| (F + f_i) / 24
struct LoadCoeffP1Tri : TriBase
{
  static constexpr int WSIZE = 3, WND = 3, WBS = 1, CSIZE = 0, RANK = 1;
  static constexpr bool FACET = false;
  struct Geo
  {
    double f[3], F, det;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[3][3], const double* w, const double*, int)
  {
    g.det = fabs(det2(xc));
    g.f[0] = w[0], g.f[1] = w[1], g.f[2] = w[2];
    g.F = w[0] + w[1] + w[2];
  }
  static BFX_DI void vec(const Geo& g, double (&out)[3])
  {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      out[i] = g.det * (g.F + g.f[i]) * (1.0 / 24.0);
  }
};

// inner(u, v)*ds on the local facet: len/6 [[2,1],[1,2]] on the facet's two vertices
struct FacetMassP1Tri : TriBase
{
  static constexpr int WSIZE = 0, WND = 0, WBS = 1, CSIZE = 0, RANK = 2;
  static constexpr bool FACET = true;
  struct Geo
  {
    int a, b;
    double len;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[3][3], const double*, const double*, int lf)
  {
    facet(xc, lf, g.a, g.b, g.len);
  }
  static BFX_DI void row(const Geo& g, int i, double (&out)[3])
  {
    const bool on = (i == g.a) || (i == g.b);
#pragma unroll
    for (int j = 0; j < 3; ++j)
    {
      const bool onj = (j == g.a) || (j == g.b);
      out[j] = (on && onj) ? (i == j ? g.len * (1.0 / 3.0) : g.len * (1.0 / 6.0)) : 0.0;
    }
  }
};

// inner(c0, v)*ds
struct FacetConstP1Tri : TriBase
{
  static constexpr int WSIZE = 0, WND = 0, WBS = 1, CSIZE = 1, RANK = 1;
  static constexpr bool FACET = true;
  struct Geo
  {
    int a, b;
    double v;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[3][3], const double*, const double* c, int lf)
  {
    double len;
    facet(xc, lf, g.a, g.b, len);
    g.v = 0.5 * len * c[0];
  }
  static BFX_DI void vec(const Geo& g, double (&out)[3])
  {
#pragma unroll
    for (int i = 0; i < 3; ++i)
      out[i] = (i == g.a || i == g.b) ? g.v : 0.0;
  }
};

// inner(avg(u), avg(v))*dS, P1 triangles (python/test/unit/fem/test_ghost_mesh_assembly.py:104-122): macro element
// of an interior facet.  coordinate_dofs = [cell0 (3 nodes), cell1 (3 nodes)]; A is 6 x 6 with the 2 x 2 block
// layout of fem/assemble_matrix_impl.h:581-587; entity_local_index[0], [1] arrive packed as lf0 + 8 lf1.
// avg(w) = (w+ + w-)/2, and on the edge only the two edge vertices of each cell contribute:
// A[(s,a),(t,b)] = len (1 + delta_ab) / 24, vertices a, b matched between the cells through their coordinates.
struct AvgMassP1TriDS
{
  static constexpr int NX = 6, ND = 6, BS = 1, WSIZE = 0, WND = 0, WBS = 1, CSIZE = 0, RANK = 2;
  static constexpr bool FACET = true;
  struct Geo
  {
    int v[2][2];
    double len;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[6][3], const double*, const double*, int lf)
  {
    const int lfs[2] = {lf & 7, lf >> 3};
#pragma unroll
    for (int s = 0; s < 2; ++s)
    {
      double c3[3][3];
#pragma unroll
      for (int n = 0; n < 3; ++n)
#pragma unroll
        for (int m = 0; m < 3; ++m)
          c3[n][m] = xc[3 * s + n][m];
      double l;
      TriBase::facet(c3, lfs[s], g.v[s][0], g.v[s][1], l);
      g.len = l;
    }
    const double d = fabs(xc[g.v[0][0]][0] - xc[3 + g.v[1][0]][0]) + fabs(xc[g.v[0][0]][1] - xc[3 + g.v[1][0]][1])
                     + fabs(xc[g.v[0][0]][2] - xc[3 + g.v[1][0]][2]);
    if (d > 1e-12 * (1.0 + g.len))
    {
      const int t = g.v[1][0];
      g.v[1][0] = g.v[1][1];
      g.v[1][1] = t;
    }
  }
  static BFX_DI void row(const Geo& g, int i, double (&out)[6])
  {
    const int s = i / 3, vi = i - 3 * s;
    const int a = vi == g.v[s][0] ? 0 : (vi == g.v[s][1] ? 1 : -1);
#pragma unroll
    for (int j = 0; j < 6; ++j)
    {
      const int t = j / 3, vj = j - 3 * t;
      const int b = vj == g.v[t][0] ? 0 : (vj == g.v[t][1] ? 1 : -1);
      out[j] = (a < 0 || b < 0) ? 0.0 : g.len * (a == b ? 2.0 : 1.0) * (1.0 / 24.0);
    }
  }
};

// Shared edge of the two cells of an interior facet (macro cell [cell0 | cell1]): local vertices of the edge in each
// cell, cell1's pair ordered like cell0's, and the edge length
struct TriDSBase
{
  static constexpr int NX = 6, BS = 1;
  static BFX_DI double edge(const double (&xc)[6][3], int lf, int (&v)[2][2])
  {
    const int lfs[2] = {lf & 7, lf >> 3};
    double len = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s)
    {
      double c3[3][3];
#pragma unroll
      for (int n = 0; n < 3; ++n)
#pragma unroll
        for (int m = 0; m < 3; ++m)
          c3[n][m] = xc[3 * s + n][m];
      TriBase::facet(c3, lfs[s], v[s][0], v[s][1], len);
    }
    // (dynamic row index: select, so that xc stays in registers)
    double p[3] = {0, 0, 0}, q[3] = {0, 0, 0};
#pragma unroll
    for (int n = 0; n < 3; ++n)
#pragma unroll
      for (int m = 0; m < 3; ++m)
      {
        p[m] = n == v[0][0] ? xc[n][m] : p[m];
        q[m] = n == v[1][0] ? xc[3 + n][m] : q[m];
      }
    if (fabs(p[0] - q[0]) + fabs(p[1] - q[1]) + fabs(p[2] - q[2]) > 1e-12 * (1.0 + len))
    {
      const int t = v[1][0];
      v[1][0] = v[1][1];
      v[1][1] = t;
    }
    return len;
  }
};

// conj(avg(v))*dS, P1 triangles (python/test/unit/fem/test_assembler.py:1003): element vector [cell0 | cell1];
// int_edge phi = len / 2 on the two edge vertices of each cell, avg halves it
struct AvgLoadP1TriDS : TriDSBase
{
  static constexpr int ND = 6, WSIZE = 0, WND = 0, WBS = 1, CSIZE = 0, RANK = 1;
  static constexpr bool FACET = true;
  struct Geo
  {
    int v[2][2];
    double len;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[6][3], const double*, const double*, int lf)
  {
    g.len = edge(xc, lf, g.v);
  }
  static BFX_DI void vec(const Geo& g, double (&out)[6])
  {
#pragma unroll
    for (int i = 0; i < 6; ++i)
    {
      const int s = i / 3, vi = i - 3 * s;
      out[i] = (vi == g.v[s][0] || vi == g.v[s][1]) ? 0.25 * g.len : 0.0;
    }
  }
};

// functional 1*dS: the length of the interior facet (python/test/unit/fem/test_assemble_domains.py:203-210)
struct OneTriDS : TriDSBase
{
  static constexpr int ND = 6, WSIZE = 0, WND = 0, WBS = 1, CSIZE = 0, RANK = 0;
  static constexpr bool FACET = true;
  struct Geo
  {
    double v;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[6][3], const double*, const double*, int lf)
  {
    int v[2][2];
    g.v = edge(xc, lf, v);
  }
  static BFX_DI double scalar(const Geo& g) { return g.v; }
};

// functional inner(avg(f), avg(f))*dS, f in P1 (test_assemble_domains.py:225): w = [f on cell0 (3), f on cell1 (3)];
// with h = avg(f) linear along the edge: len (ha^2 + ha hb + hb^2) / 3
struct Avg2CoeffP1TriDS : TriDSBase
{
  static constexpr int ND = 6, WSIZE = 6, WND = 6, WBS = 1, CSIZE = 0, RANK = 0;
  static constexpr bool FACET = true;
  struct Geo
  {
    double v;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[6][3], const double* w, const double*, int lf)
  {
    int v[2][2];
    const double len = edge(xc, lf, v);
    double h[2] = {0, 0};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int n = 0; n < 3; ++n)
        h[a] += 0.5 * ((n == v[0][a] ? w[n] : 0.0) + (n == v[1][a] ? w[3 + n] : 0.0));
    g.v = len * (h[0] * h[0] + h[0] * h[1] + h[1] * h[1]) * (1.0 / 3.0);
  }
  static BFX_DI double scalar(const Geo& g) { return g.v; }
};

// functional inner(f, f)*ds on an exterior facet, f in P1 (test_assemble_domains.py:224)
struct Coeff2P1TriFacet : TriBase
{
  static constexpr int WSIZE = 3, WND = 3, WBS = 1, CSIZE = 0, RANK = 0;
  static constexpr bool FACET = true;
  struct Geo
  {
    double v;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[3][3], const double* w, const double*, int lf)
  {
    int a, b;
    double len;
    facet(xc, lf, a, b, len);
    double fa = 0, fb = 0;
#pragma unroll
    for (int n = 0; n < 3; ++n)
    {
      fa = n == a ? w[n] : fa;
      fb = n == b ? w[n] : fb;
    }
    g.v = len * (fa * fa + fa * fb + fb * fb) * (1.0 / 3.0);
  }
  static BFX_DI double scalar(const Geo& g) { return g.v; }
};

// ---- tetrahedra -------------------------------------------------------------------------------
struct TetBase
{
  static constexpr int NX = 4;
  // Unscaled gradients of the barycentric coordinates: grad(lambda_a) = n[a] / det, with
  // J = [x1-x0, x2-x0, x3-x0] (FFCx convention, test_custom_jit_kernels.py:173-182).
  static BFX_DI double normals(const double (&xc)[4][3], double (&n)[4][3])
  {
    double e1[3], e2[3], e3[3];
#pragma unroll
    for (int m = 0; m < 3; ++m)
    {
      e1[m] = xc[1][m] - xc[0][m];
      e2[m] = xc[2][m] - xc[0][m];
      e3[m] = xc[3][m] - xc[0][m];
    }
    n[1][0] = e2[1] * e3[2] - e2[2] * e3[1];
    n[1][1] = e2[2] * e3[0] - e2[0] * e3[2];
    n[1][2] = e2[0] * e3[1] - e2[1] * e3[0];
    n[2][0] = e3[1] * e1[2] - e3[2] * e1[1];
    n[2][1] = e3[2] * e1[0] - e3[0] * e1[2];
    n[2][2] = e3[0] * e1[1] - e3[1] * e1[0];
    n[3][0] = e1[1] * e2[2] - e1[2] * e2[1];
    n[3][1] = e1[2] * e2[0] - e1[0] * e2[2];
    n[3][2] = e1[0] * e2[1] - e1[1] * e2[0];
#pragma unroll
    for (int m = 0; m < 3; ++m)
      n[0][m] = -(n[1][m] + n[2][m] + n[3][m]);
    return e1[0] * n[1][0] + e1[1] * n[1][1] + e1[2] * n[1][2];
  }
  // facet i is the face opposite vertex i; returns its area
  static BFX_DI double facet_area(const double (&xc)[4][3], int lf)
  {
    const int a = lf == 0 ? 1 : 0, b = lf <= 1 ? 2 : 1, c = lf == 3 ? 2 : 3;
    double u[3], v[3];
#pragma unroll
    for (int m = 0; m < 3; ++m)
    {
      u[m] = xc[b][m] - xc[a][m];
      v[m] = xc[c][m] - xc[a][m];
    }
    const double cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
    return 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
  }
};

// kappa*inner(grad u, grad v)*dx, P1: A_ij = kappa n_i.n_j / (6 |det|)
struct PoissonP1Tet : TetBase
{
  static constexpr int ND = 4, BS = 1, WSIZE = 0, WND = 0, WBS = 1, CSIZE = 1, RANK = 2;
  static constexpr bool FACET = false;
  struct Geo
  {
    double n[4][3];
    double scale;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[4][3], const double*, const double* c, int)
  {
    const double det = normals(xc, g.n);
    g.scale = c[0] / (6.0 * fabs(det));
  }
  static BFX_DI void row(const Geo& g, int i, double (&out)[4])
  {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      out[j] = g.scale * (g.n[i][0] * g.n[j][0] + g.n[i][1] * g.n[j][1] + g.n[i][2] * g.n[j][2]);
  }
};

// inner(f, v)*dx, P1, f in P1: b_i = |det| (F + f_i) / 120
struct LoadP1Tet : TetBase
{
  static constexpr int ND = 4, BS = 1, WSIZE = 4, WND = 4, WBS = 1, CSIZE = 0, RANK = 1;
  static constexpr bool FACET = false;
  struct Geo
  {
    double f[4], F, det;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[4][3], const double* w, const double*, int)
  {
    double n[4][3];
    g.det = fabs(normals(xc, n));
    g.F = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      g.f[i] = w[i];
      g.F += w[i];
    }
  }
  static BFX_DI void vec(const Geo& g, double (&out)[4])
  {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      out[i] = g.det * (g.F + g.f[i]) * (1.0 / 120.0);
  }
};

// Basix edge order of the tetrahedron: e0=(2,3) e1=(1,3) e2=(1,2) e3=(0,3) e4=(0,2) e5=(0,1)
BFX_DI constexpr int tet_ea(int k) { return k == 0 ? 2 : (k <= 2 ? 1 : 0); }
BFX_DI constexpr int tet_eb(int k) { return (k == 0 || k == 1 || k == 3) ? 3 : ((k == 2 || k == 4) ? 2 : 1); }

// kappa*inner(grad u, grad v)*dx, P2 (cpp/test/poisson.py:16-27).  With S_ab = n_a.n_b/|det|
// (6 x the P1 stiffness), phi_v = l_v(2 l_v - 1), phi_e(a,b) = 4 l_a l_b:
//   A_vv' = S_vv' (v == v' ? 1/10 : -1/30)
//   A_v,e(a,b) = (1/30) [ S_vb (v == a ? 3 : -1) + S_va (v == b ? 3 : -1) ]
//   A_e(a,b),e(c,d) = (2/15) [ (1+d_ac) S_bd + (1+d_ad) S_bc + (1+d_bc) S_ad + (1+d_bd) S_ac ]
// inner(f*g, v)*dx with f, g in P1: TWO coefficients in one integral, w = [f (4), g (4)].
// int lambda_i lambda_j lambda_k = |det| m(i,j,k) / 720 with m = 6 (i=j=k), 2 (two equal), 1 (all different)
struct LoadProdP1Tet : TetBase
{
  static constexpr int ND = 4, BS = 1, WSIZE = 8, WND = 4, WBS = 1, NCOEF = 2, CSIZE = 0, RANK = 1;
  static constexpr bool FACET = false;
  struct Geo
  {
    double b[4];
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[4][3], const double* w, const double*, int)
  {
    double n[4][3];
    const double det = fabs(normals(xc, n));
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
          const double m = (i == j && j == k) ? 6.0 : ((i == j || j == k || i == k) ? 2.0 : 1.0);
          acc = fma(m * w[j], w[4 + k], acc);
        }
      g.b[i] = det * acc * (1.0 / 720.0);
    }
  }
  static BFX_DI void vec(const Geo& g, double (&out)[4])
  {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      out[i] = g.b[i];
  }
};

struct PoissonP2Tet : TetBase
{
  static constexpr int ND = 10, BS = 1, WSIZE = 0, WND = 0, WBS = 1, CSIZE = 1, RANK = 2;
  static constexpr bool FACET = false;
  struct Geo
  {
    double S[4][4];
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[4][3], const double*, const double* c, int)
  {
    double n[4][3];
    const double det = normals(xc, n);
    const double s = c[0] / fabs(det);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = a; b < 4; ++b)
      {
        g.S[a][b] = s * (n[a][0] * n[b][0] + n[a][1] * n[b][1] + n[a][2] * n[b][2]);
        g.S[b][a] = g.S[a][b];
      }
  }
  static BFX_DI double entry(const Geo& g, int i, int j)
  {
    if (i < 4 && j < 4)
      return g.S[i][j] * (i == j ? 0.1 : (-1.0 / 30.0));
    if (i < 4 || j < 4)
    {
      const int v = i < 4 ? i : j, e = (i < 4 ? j : i) - 4;
      const int a = tet_ea(e), b = tet_eb(e);
      return (1.0 / 30.0) * (g.S[v][b] * (v == a ? 3.0 : -1.0) + g.S[v][a] * (v == b ? 3.0 : -1.0));
    }
    const int a = tet_ea(i - 4), b = tet_eb(i - 4), cc = tet_ea(j - 4), d = tet_eb(j - 4);
    return (2.0 / 15.0)
           * ((a == cc ? 2.0 : 1.0) * g.S[b][d] + (a == d ? 2.0 : 1.0) * g.S[b][cc]
              + (b == cc ? 2.0 : 1.0) * g.S[a][d] + (b == d ? 2.0 : 1.0) * g.S[a][cc]);
  }
  static BFX_DI void row(const Geo& g, int i, double (&out)[10])
  {
#pragma unroll
    for (int j = 0; j < 10; ++j)
      out[j] = entry(g, i, j);
  }
};

// inner(f, v)*dx, P2, f in P2: b_i = |det| sum_j M_ij f_j with the exact reference mass matrix
//   vv: 1/420 (same) 1/2520 (different); ve: -1/630 (v in e) -1/420 (v not in e);
//   ee: 4/315 (same) 2/315 (share a vertex) 1/315 (disjoint)
struct LoadP2Tet : TetBase
{
  static constexpr int ND = 10, BS = 1, WSIZE = 10, WND = 10, WBS = 1, CSIZE = 0, RANK = 1;
  static constexpr bool FACET = false;
  struct Geo
  {
    double f[10], det;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[4][3], const double* w, const double*, int)
  {
    double n[4][3];
    g.det = fabs(normals(xc, n));
#pragma unroll
    for (int i = 0; i < 10; ++i)
      g.f[i] = w[i];
  }
  static BFX_DI double mass(int i, int j)
  {
    if (i < 4 && j < 4)
      return i == j ? 1.0 / 420.0 : 1.0 / 2520.0;
    if (i < 4 || j < 4)
    {
      const int v = i < 4 ? i : j, e = (i < 4 ? j : i) - 4;
      return (v == tet_ea(e) || v == tet_eb(e)) ? -1.0 / 630.0 : -1.0 / 420.0;
    }
    const int a = tet_ea(i - 4), b = tet_eb(i - 4), c = tet_ea(j - 4), d = tet_eb(j - 4);
    const int shared = (a == c) + (a == d) + (b == c) + (b == d);
    return shared == 2 ? 4.0 / 315.0 : (shared == 1 ? 2.0 / 315.0 : 1.0 / 315.0);
  }
  static BFX_DI void vec(const Geo& g, double (&out)[10])
  {
#pragma unroll
    for (int i = 0; i < 10; ++i)
    {
      double s = 0;
#pragma unroll
      for (int j = 0; j < 10; ++j)
        s += mass(i, j) * g.f[j];
      out[i] = g.det * s;
    }
  }
};

// inner(g, v)*ds on a tetrahedron facet, g in P1: b_i = area (G + g_i) / 12 on the facet's vertices
struct FacetLoadP1Tet : TetBase
{
  static constexpr int ND = 4, BS = 1, WSIZE = 4, WND = 4, WBS = 1, CSIZE = 0, RANK = 1;
  static constexpr bool FACET = true;
  struct Geo
  {
    double f[4], G, area;
    int lf;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[4][3], const double* w, const double*, int lf)
  {
    g.area = facet_area(xc, lf);
    g.lf = lf;
    g.G = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      g.f[i] = w[i];
      g.G += (i == lf) ? 0.0 : w[i];
    }
  }
  static BFX_DI void vec(const Geo& g, double (&out)[4])
  {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      out[i] = (i == g.lf) ? 0.0 : g.area * (g.G + g.f[i]) * (1.0 / 12.0);
  }
};

// inner(u, v)*ds on a tetrahedron facet: area (1 + delta_ij) / 12
struct FacetMassP1Tet : TetBase
{
  static constexpr int ND = 4, BS = 1, WSIZE = 0, WND = 0, WBS = 1, CSIZE = 0, RANK = 2;
  static constexpr bool FACET = true;
  struct Geo
  {
    double area;
    int lf;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[4][3], const double*, const double*, int lf)
  {
    g.area = facet_area(xc, lf);
    g.lf = lf;
  }
  static BFX_DI void row(const Geo& g, int i, double (&out)[4])
  {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      out[j] = (i == g.lf || j == g.lf) ? 0.0 : g.area * (i == j ? 1.0 / 6.0 : 1.0 / 12.0);
  }
};

// action(a, ui): the element matrix of a bilinear kernel E applied to the coefficient dofs,
// be_i = sum_j A_ij w_j (cpp/demo/poisson_matrix_free/poisson.py: M = action(a, ui)).  The matrix-free operator
// of SURVEY.md §8f: one assemble_vector per operator application instead of assembling and reading the matrix.
template <class E>
struct ActionOf
{
  static constexpr int NX = E::NX, ND = E::ND, BS = E::BS, WSIZE = E::ND * E::BS, WND = E::ND, WBS = E::BS,
                       CSIZE = E::CSIZE, RANK = 1;
  static constexpr bool FACET = false;
  struct Geo
  {
    typename E::Geo g;
    double u[E::ND * E::BS];
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[NX][3], const double* w, const double* c, int lf)
  {
    E::prepare(g.g, xc, nullptr, c, lf);
#pragma unroll
    for (int j = 0; j < ND * BS; ++j)
      g.u[j] = w[j];
  }
  static BFX_DI void vec(const Geo& g, double (&out)[E::ND * E::BS])
  {
#pragma unroll
    for (int i = 0; i < ND * BS; ++i)
    {
      double row[E::ND * E::BS];
      E::row(g.g, i, row);
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < ND * BS; ++j)
        acc = fma(row[j], g.u[j], acc);
      out[i] = acc;
    }
  }
};

// M = w^2 dx, w in P1 on tetrahedra (functional, RANK 0): |det| w^T Mhat w, Mhat = (1 + delta_ij) / 120
struct L2Norm2P1Tet : TetBase
{
  static constexpr int ND = 4, BS = 1, WSIZE = 4, WND = 4, WBS = 1, CSIZE = 0, RANK = 0;
  static constexpr bool FACET = false;
  struct Geo
  {
    double v;
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[4][3], const double* w, const double*, int)
  {
    double n[4][3];
    const double det = fabs(normals(xc, n));
    const double s = w[0] + w[1] + w[2] + w[3];
    g.v = det * (s * s + w[0] * w[0] + w[1] * w[1] + w[2] * w[2] + w[3] * w[3]) * (1.0 / 120.0);
  }
  static BFX_DI double scalar(const Geo& g) { return g.v; }
};

// ---- hexahedra (Q1, tensor node order, x fastest) ------------------------------------------------
struct HexQ1
{
  static constexpr int NX = 8;
  // reference gradient of node n at X
  static BFX_DI void dphi(int n, const double (&X)[3], double (&d)[3], double& phi)
  {
    const double fx = (n & 1) ? X[0] : 1.0 - X[0], fy = (n & 2) ? X[1] : 1.0 - X[1], fz = (n & 4) ? X[2] : 1.0 - X[2];
    const double sx = (n & 1) ? 1.0 : -1.0, sy = (n & 2) ? 1.0 : -1.0, sz = (n & 4) ? 1.0 : -1.0;
    phi = fx * fy * fz;
    d[0] = sx * fy * fz;
    d[1] = fx * sy * fz;
    d[2] = fx * fy * sz;
  }
  // K = J^{-1}, returns det J at X for the trilinear map given by xc
  static BFX_DI double jacobian_inverse(const double (&xc)[8][3], const double (&X)[3], double (&K)[3][3])
  {
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int n = 0; n < 8; ++n)
    {
      double d[3], phi;
      dphi(n, X, d, phi);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int a = 0; a < 3; ++a)
          J[i][a] = fma(xc[n][i], d[a], J[i][a]);
    }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    K[0][0] = c00 * id;
    K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    K[1][0] = c01 * id;
    K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    K[2][0] = c02 * id;
    K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    return det;
  }
};

// inner(f, v)*dx, Q1 vector (bs = 3), f Q1 vector coefficient; 3x3x3 Gauss (exact on affine cells)
struct LoadQ1Hex : HexQ1
{
  static constexpr int ND = 8, BS = 3, WSIZE = 24, WND = 8, WBS = 3, CSIZE = 0, RANK = 1;
  static constexpr bool FACET = false;
  struct Geo
  {
    double b[24];
  };
  static BFX_DI void prepare(Geo& g, const double (&xc)[8][3], const double* w, const double*, int)
  {
    const double gl[3] = {0.1127016653792583, 0.5, 0.8872983346207417};
    const double gw[3] = {0.2777777777777778, 0.4444444444444444, 0.2777777777777778};
#pragma unroll
    for (int i = 0; i < 24; ++i)
      g.b[i] = 0;
    for (int q = 0; q < 27; ++q)
    {
      const int qa = q % 3, qb = (q / 3) % 3, qc = q / 9;
      const double X[3] = {gl[qa], gl[qb], gl[qc]};
      double K[3][3];
      const double wt = gw[qa] * gw[qb] * gw[qc] * fabs(jacobian_inverse(xc, X, K));
      double phi[8], f[3] = {0, 0, 0};
#pragma unroll
      for (int n = 0; n < 8; ++n)
      {
        double d[3];
        dphi(n, X, d, phi[n]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          f[k] = fma(w[3 * n + k], phi[n], f[k]);
      }
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int k = 0; k < 3; ++k)
          g.b[3 * n + k] = fma(wt * f[k], phi[n], g.b[3 * n + k]);
    }
  }
  static BFX_DI void vec(const Geo& g, double (&out)[24])
  {
#pragma unroll
    for (int i = 0; i < 24; ++i)
      out[i] = g.b[i];
  }
};

#undef BFX_DI
} // namespace el
} // namespace bfx
