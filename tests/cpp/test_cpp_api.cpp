// C++ API mirror test: the call sequence of cpp/demo/poisson/main.cpp:207-226 (assemble A with bcs,
// set_diagonal, assemble b, apply_lifting, scatter_rev, bc.set) and the known answers of
// cpp/test/matrix.cpp:66-120 (A.1 = 0) / cpp/test/vector.cpp (norms), on a 3-D P1 box, through
// dolfinx_b200.h -> libbfx.so.  Prints "CPP_API_OK" on success.
#include "../../dolfinx_b200/cpp/dolfinx_b200.h"
#include <cstdio>
#include <functional>

using namespace dolfinx_b200;

#define REQUIRE(cond)                                                                                                  \
  do                                                                                                                   \
  {                                                                                                                    \
    if (!(cond))                                                                                                       \
    {                                                                                                                  \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                                                    \
      return 1;                                                                                                        \
    }                                                                                                                  \
  } while (0)

// compile-only: la::transpose of the mirror (its device part is covered by tests/test_gpu_parity.py::test_transpose)
[[maybe_unused]] static double transpose_compiles(const dolfinx_b200::la::MatrixCSR<double>& A)
{
  auto AT = dolfinx_b200::la::transpose(A);
  return AT.squared_norm();
}

int main()
{
  int ndev = 0;
  if (bfx_device_count(&ndev) != BFX_OK or ndev == 0)
  {
    std::printf("no CUDA device: %s\n", bfx_last_error());
    return 77;
  }
  // create_box conventions (mesh/generation.h:333-427): n^3 cubes, 6 tets each
  const int n = 6;
  const int nv1 = n + 1;
  std::vector<double> x;
  for (int k = 0; k <= n; ++k)
    for (int j = 0; j <= n; ++j)
      for (int i = 0; i <= n; ++i)
        x.insert(x.end(), {double(i) / n, double(j) / n, double(k) / n});
  std::vector<std::int32_t> cells;
  const int T[6][4] = {{0, 1, 3, 7}, {0, 1, 7, 5}, {0, 5, 7, 4}, {0, 3, 2, 7}, {0, 6, 4, 7}, {0, 2, 6, 7}};
  for (int k = 0; k < n; ++k)
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i)
      {
        int v[8];
        for (int c = 0; c < 8; ++c)
          v[c] = ((k + ((c >> 2) & 1)) * nv1 + (j + ((c >> 1) & 1))) * nv1 + (i + (c & 1));
        for (auto& t : T)
          for (int a = 0; a < 4; ++a)
            cells.push_back(v[t[a]]);
      }
  const std::int32_t ndofs = nv1 * nv1 * nv1;
  common::Comm comm;
  auto mesh = std::make_shared<fem::Mesh>(comm, x, cells, 4);
  auto imap = std::make_shared<common::IndexMap>(comm, ndofs);
  auto dofmap = std::make_shared<fem::DofMap>(cells, 4, 1, imap);
  auto V = std::make_shared<fem::FunctionSpace>(mesh, dofmap);

  // a = kappa grad u . grad v dx, L = f v dx
  auto kappa = std::make_shared<fem::Constant<double>>(2.0);
  auto f = std::make_shared<fem::Function<double>>(V);
  {
    std::vector<double> fh(ndofs);
    for (int d = 0; d < ndofs; ++d)
      fh[d] = 10.0 * std::exp(-((x[3 * d] - 0.5) * (x[3 * d] - 0.5) + (x[3 * d + 1] - 0.5) * (x[3 * d + 1] - 0.5)) / 0.02);
    f->x()->array().copy_from(fh);
  }
  fem::Form<double> a({V, V}, {{{fem::IntegralType::cell, 0}, {BFX_K_POISSON_P1_TET_A, {}, {}}}}, {}, {kappa});
  fem::Form<double> L({V}, {{{fem::IntegralType::cell, 0}, {BFX_K_LOAD_P1_TET_L, {}, {0}}}}, {f}, {});

  la::SparsityPattern sp = fem::create_sparsity_pattern(a);
  sp.finalize();
  REQUIRE(sp.num_nonzeros() == (std::int64_t)ndofs + 2 * (3 * n * nv1 * nv1 + 3 * n * n * nv1 + n * n * n)); // V + 2E
  la::MatrixCSR<double> A(sp);

  // (1) no bcs: A.1 = 0, symmetric
  fem::assemble_matrix(A, a);
  A.scatter_rev();
  {
    la::Vector<double> one(imap, 1), y(imap, 1);
    one.set(1.0);
    A.mult(one, y);
    REQUIRE(la::norm(y, la::Norm::linf) < 1e-13);
    REQUIRE(std::abs(la::norm(one, la::Norm::l1) - ndofs) < 1e-9);
    REQUIRE(std::abs(la::norm(one, la::Norm::l2) - std::sqrt((double)ndofs)) < 1e-9);
    auto D = A.to_dense();
    double asym = 0;
    for (int i = 0; i < ndofs; ++i)
      for (int j = 0; j < i; ++j)
        asym = std::max(asym, std::abs(D[(std::size_t)i * ndofs + j] - D[(std::size_t)j * ndofs + i]));
    REQUIRE(asym < 1e-14);
  }
  // out-of-pattern insertion throws like the reference
  {
    bool threw = false;
    try
    {
      std::vector<double> v{1.0};
      std::vector<std::int32_t> r{0}, c{ndofs - 1};
      A.add<1, 1>(v, r, c);
    }
    catch (const std::runtime_error& e)
    {
      threw = std::string(e.what()) == "Entry not in sparsity";
    }
    REQUIRE(threw);
  }

  // re-assembly: adds onto the matrix (fem/assembler.h:497-498); after set(0) the aggregated kernel overwrites instead and
  // reproduces the first assembly; the reference's call shape with the insertion functor does the same
  {
    const double n1 = A.squared_norm();
    fem::assemble_matrix(A, a);
    REQUIRE(std::abs(A.squared_norm() - 4 * n1) <= 1e-12 * n1);
    A.set(0.0);
    REQUIRE(A.known_zero() and A.squared_norm() == 0.0);
    fem::assemble_matrix<double>(A.mat_add_values(), a);
    REQUIRE(!A.known_zero() and std::abs(A.squared_norm() - n1) <= 1e-13 * n1);
    A.set(2.5); // MatrixCSR::set(value), filled on the device
    REQUIRE(std::abs(A.squared_norm() - 6.25 * (double)sp.num_nonzeros()) <= 1e-9);
    la::Vector<double> v(imap, 1);
    v.set(-3.0);
    REQUIRE(std::abs(la::norm(v, la::Norm::l1) - 3.0 * ndofs) < 1e-9 and std::abs(la::norm(v, la::Norm::linf) - 3.0) == 0.0);
  }

  // (2) Dirichlet problem: u = g on x0 in {0,1}, g = 1 + 3 x1
  std::vector<std::int32_t> bdofs;
  for (int d = 0; d < ndofs; ++d)
    if (x[3 * d] < 1e-12 or x[3 * d] > 1 - 1e-12)
      bdofs.push_back(d);
  auto g = std::make_shared<fem::Function<double>>(V);
  {
    std::vector<double> gh(ndofs);
    for (int d = 0; d < ndofs; ++d)
      gh[d] = 1.0 + 3.0 * x[3 * d + 1];
    g->x()->array().copy_from(gh);
  }
  fem::DirichletBC<double> bc(std::shared_ptr<const fem::Function<double>>(g), bdofs, V);
  std::vector<std::reference_wrapper<const fem::DirichletBC<double>>> bcs{bc};
  A.set(0.0);
  fem::assemble_matrix(A, a, bcs);
  A.scatter_rev();
  fem::set_diagonal(A, *V, bcs, 1.0);
  la::Vector<double> b(imap, 1);
  fem::assemble_vector(b, L);
  fem::apply_lifting<double>(b, {a}, {bcs}, {}, 1.0);
  b.scatter_rev(std::plus<double>{});
  bc.set(b.array(), nullptr, 1.0);

  // CG on the device (cf. cpp/demo/poisson_matrix_free/main.cpp:84-132)
  la::Vector<double> u(imap, 1), r(imap, 1), p(imap, 1), q(imap, 1);
  const std::int64_t N = ndofs;
  check(bfx_memcpy(r.array().data(), b.array().data(), N * sizeof(double), nullptr));
  check(bfx_memcpy(p.array().data(), b.array().data(), N * sizeof(double), nullptr));
  double rr = la::inner_product(r, r);
  const double rr0 = rr;
  int it = 0;
  for (; it < 500 and rr > 1e-24 * rr0; ++it)
  {
    q.array().fill_zero();
    A.mult(p, q);
    const double alpha = rr / la::inner_product(p, q);
    check(bfx_axpy(N, alpha, p.array().data(), u.array().data(), nullptr));
    check(bfx_axpy(N, -alpha, q.array().data(), r.array().data(), nullptr));
    const double rr_new = la::inner_product(r, r);
    const double beta = rr_new / rr;
    rr = rr_new;
    // p = r + beta p
    std::vector<double> ph = p.array().to_host(), rh = r.array().to_host();
    for (std::int64_t i = 0; i < N; ++i)
      ph[i] = rh[i] + beta * ph[i];
    p.array().copy_from(ph);
  }
  REQUIRE(rr <= 1e-20 * rr0);
  auto uh = u.array().to_host();
  for (auto d : bdofs)
    REQUIRE(std::abs(uh[d] - (1.0 + 3.0 * x[3 * d + 1])) < 1e-9);
  std::printf("CG iterations %d, |r|/|r0| = %.2e\n", it, std::sqrt(rr / rr0));
  std::printf("CPP_API_OK\n");
  return 0;
}
