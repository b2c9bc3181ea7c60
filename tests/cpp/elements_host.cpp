// Host build of dolfinx_b200/csrc/elements.cuh: the hand-written element kernels that run on the GPU, compiled by g++
// (the CUDA qualifiers expand to nothing outside nvcc), so that the CPU suite can compare the very source the device
// executes with the oracle's quadrature kernels (tests/test_elements_host.py).  Test infrastructure only.
#include <math.h> // sqrt, fabs: global names, as nvcc provides them to device code
#include "../../dolfinx_b200/csrc/elements.cuh"
#include <cstring>

using namespace bfx;

namespace
{
template <class E>
int run_matrix(const double* xc_flat, const double* w, const double* c, int lf, double* A)
{
  constexpr int N = E::ND * E::BS;
  double xc[E::NX][3];
  std::memcpy(xc, xc_flat, sizeof(xc));
  typename E::Geo g;
  E::prepare(g, xc, w, c, lf);
  for (int i = 0; i < N; ++i)
  {
    double row[N];
    E::row(g, i, row);
    for (int j = 0; j < N; ++j)
      A[i * N + j] = row[j];
  }
  return N * N;
}

template <class E>
int run_vector(const double* xc_flat, const double* w, const double* c, int lf, double* b)
{
  constexpr int N = E::ND * E::BS;
  double xc[E::NX][3];
  std::memcpy(xc, xc_flat, sizeof(xc));
  typename E::Geo g;
  E::prepare(g, xc, w, c, lf);
  double out[N];
  E::vec(g, out);
  for (int i = 0; i < N; ++i)
    b[i] = out[i];
  return N;
}

template <class E>
int run_scalar(const double* xc_flat, const double* w, const double* c, int lf, double* v)
{
  double xc[E::NX][3];
  std::memcpy(xc, xc_flat, sizeof(xc));
  typename E::Geo g;
  E::prepare(g, xc, w, c, lf);
  v[0] = E::scalar(g);
  return 1;
}
} // namespace

/// Element tensor of kernel `id` for one cell: returns the number of scalars written to out, or -1
extern "C" int elements_host_tabulate(int id, const double* xc, const double* w, const double* c, int lf, double* out)
{
  switch (id)
  {
  case BFX_K_LAPLACE_P1_TRI_A: return run_matrix<el::LaplaceP1Tri>(xc, w, c, lf, out);
  case BFX_K_MASS_COEFF_P1_TRI_A: return run_matrix<el::MassCoeffP1Tri>(xc, w, c, lf, out);
  case BFX_K_FACET_MASS_P1_TRI_A: return run_matrix<el::FacetMassP1Tri>(xc, w, c, lf, out);
  case BFX_K_POISSON_P1_TET_A: return run_matrix<el::PoissonP1Tet>(xc, w, c, lf, out);
  case BFX_K_POISSON_P2_TET_A: return run_matrix<el::PoissonP2Tet>(xc, w, c, lf, out);
  case BFX_K_FACET_MASS_P1_TET_A: return run_matrix<el::FacetMassP1Tet>(xc, w, c, lf, out);
  case BFX_K_AVG_MASS_P1_TRI_DS: return run_matrix<el::AvgMassP1TriDS>(xc, w, c, lf, out);
  case BFX_K_SOURCE_P1_TRI_L: return run_vector<el::SourceP1Tri>(xc, w, c, lf, out);
  case BFX_K_LOAD_COEFF_P1_TRI_L: return run_vector<el::LoadCoeffP1Tri>(xc, w, c, lf, out);
  case BFX_K_FACET_CONST_P1_TRI_L: return run_vector<el::FacetConstP1Tri>(xc, w, c, lf, out);
  case BFX_K_LOAD_P1_TET_L: return run_vector<el::LoadP1Tet>(xc, w, c, lf, out);
  case BFX_K_LOAD_P2_TET_L: return run_vector<el::LoadP2Tet>(xc, w, c, lf, out);
  case BFX_K_LOAD_Q1_HEX_L: return run_vector<el::LoadQ1Hex>(xc, w, c, lf, out);
  case BFX_K_FACET_LOAD_P1_TET_L: return run_vector<el::FacetLoadP1Tet>(xc, w, c, lf, out);
  case BFX_K_ACTION_POISSON_P1_TET_L: return run_vector<el::ActionOf<el::PoissonP1Tet>>(xc, w, c, lf, out);
  case BFX_K_ACTION_POISSON_P2_TET_L: return run_vector<el::ActionOf<el::PoissonP2Tet>>(xc, w, c, lf, out);
  case BFX_K_AVG_LOAD_P1_TRI_DS_L: return run_vector<el::AvgLoadP1TriDS>(xc, w, c, lf, out);
  case BFX_K_LOAD_PROD_P1_TET_L: return run_vector<el::LoadProdP1Tet>(xc, w, c, lf, out);
  case BFX_K_ONE_TRI_DS_M: return run_scalar<el::OneTriDS>(xc, w, c, lf, out);
  case BFX_K_AVG2_COEFF_P1_TRI_DS_M: return run_scalar<el::Avg2CoeffP1TriDS>(xc, w, c, lf, out);
  case BFX_K_COEFF2_P1_TRI_FACET_M: return run_scalar<el::Coeff2P1TriFacet>(xc, w, c, lf, out);
  case BFX_K_L2NORM2_P1_TET_M: return run_scalar<el::L2Norm2P1Tet>(xc, w, c, lf, out);
  default: return -1;
  }
}
