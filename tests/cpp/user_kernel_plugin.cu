// A caller's own forms for the kernel plug point (include/bfx_plugin.cuh): three kernels libbfx.so does not ship,
// written with the element conventions of dolfinx_b200/csrc/elements.cuh, compiled into their own shared library by
// tests/test_plugin.py and registered with bfx_register_kernel under ids >= BFX_K_USER_BASE.
#include "../../include/bfx_plugin.cuh"

using namespace bfx::el;

// a = inner(u, v) dx on P1 tetrahedra: |det| (1 + delta_ij) / 120
struct MassP1Tet : TetBase
{
  static constexpr int ND = 4, BS = 1, WSIZE = 0, WND = 0, WBS = 1, CSIZE = 0, RANK = 2;
  static constexpr bool FACET = false;
  struct Geo
  {
    double s;
  };
  static __device__ __forceinline__ void prepare(Geo& g, const double (&xc)[4][3], const double*, const double*, int)
  {
    double n[4][3];
    g.s = fabs(normals(xc, n)) * (1.0 / 120.0);
  }
  static __device__ __forceinline__ void row(const Geo& g, int i, double (&out)[4])
  {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      out[j] = i == j ? 2.0 * g.s : g.s;
  }
};

// L = c0 v dx: |det| c0 / 24 per vertex
struct SourceConstP1Tet : TetBase
{
  static constexpr int ND = 4, BS = 1, WSIZE = 0, WND = 0, WBS = 1, CSIZE = 1, RANK = 1;
  static constexpr bool FACET = false;
  struct Geo
  {
    double v;
  };
  static __device__ __forceinline__ void prepare(Geo& g, const double (&xc)[4][3], const double*, const double* c, int)
  {
    double n[4][3];
    g.v = fabs(normals(xc, n)) * c[0] * (1.0 / 24.0);
  }
  static __device__ __forceinline__ void vec(const Geo& g, double (&out)[4])
  {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      out[i] = g.v;
  }
};

// M = 1 dx: the volume
struct VolumeTet : TetBase
{
  static constexpr int ND = 4, BS = 1, WSIZE = 0, WND = 0, WBS = 1, CSIZE = 0, RANK = 0;
  static constexpr bool FACET = false;
  struct Geo
  {
    double v;
  };
  static __device__ __forceinline__ void prepare(Geo& g, const double (&xc)[4][3], const double*, const double*, int)
  {
    double n[4][3];
    g.v = fabs(normals(xc, n)) * (1.0 / 6.0);
  }
  static __device__ __forceinline__ double scalar(const Geo& g) { return g.v; }
};

BFX_PLUGIN_KERNEL(MassP1Tet, plug_mass_p1_tet)
BFX_PLUGIN_KERNEL(SourceConstP1Tet, plug_source_const_p1_tet)
BFX_PLUGIN_KERNEL(VolumeTet, plug_volume_tet)
