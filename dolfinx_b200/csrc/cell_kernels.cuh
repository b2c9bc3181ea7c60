// Generic cell-parallel kernels of the assembly path, as templates over the element struct (elements.cuh conventions):
// k_matrix_cells (fem/assemble_matrix_impl.h:92-200, 264-379; lifting :27-34 + assemble_vector_impl.h:361-414),
// k_vector_cells (fem/assemble_vector_impl.h:72-116, 157-215), k_scalar_cells (fem/assemble_scalar_impl.h:32-168).
// libbfx.so instantiates them for the kernels it ships (assemble.cu); a caller's own form instantiates them in its own
// translation unit through include/bfx_plugin.cuh and registers the launcher with bfx_register_kernel - the plug point
// that replaces the tabulate_tensor function pointer of fem::integral_data (fem/kernel.h:18-20, fem/Form.h:76-78).
#pragma once
#include "assemble.cuh"
#include "asm_device.cuh"

namespace bfx
{
// ---------------------------------------------------------------------------------------------
// generic thread-per-cell matrix kernel.  MODE 0: assemble with bc row/col zeroing into CSR
// (fp64 RED); MODE 1: lifting (b -= alpha Ae (g - x0) on marked columns).
// ---------------------------------------------------------------------------------------------
template <class E, typename PosT, int MODE>
__global__ void __launch_bounds__(128) k_matrix_cells(const AsmArgs a)
{
  constexpr int NX = E::NX, ND = E::ND, BS = E::BS, N = ND * BS;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += (int64_t)gridDim.x * blockDim.x)
  {
    int32_t cell;
    int lf = 0;
    if (a.entities)
    {
      cell = a.entities[2 * e];
      lf = a.entities[2 * e + 1];
    }
    else
      cell = a.cells ? a.cells[e] : (int32_t)e;

    int32_t d0[ND], d1[ND];
    load_ints<ND>(a.dofmap0 + (int64_t)cell * ND, d0);
    load_ints<ND>(a.dofmap1 + (int64_t)cell * ND, d1);

    uint32_t m0 = 0, m1 = 0; // bc marker bit per scalar row / column of Ae
    if (a.bc1)
    {
#pragma unroll
      for (int j = 0; j < N; ++j)
        m1 |= (a.bc1[(int64_t)BS * d1[j / BS] + j % BS] ? 1u : 0u) << j;
    }
    if (MODE == 1 && m1 == 0)
      continue; // has_bc (assemble_matrix_impl.h:27-34,139-143)
    if (MODE == 0 && a.bc0)
    {
#pragma unroll
      for (int i = 0; i < N; ++i)
        m0 |= (a.bc0[(int64_t)BS * d0[i / BS] + i % BS] ? 1u : 0u) << i;
    }

    int32_t xd[NX];
    load_ints<NX>(a.x_dofmap + (int64_t)cell * NX, xd);
    double xc[NX][3];
    gather_coords<NX>(a.x, xd, xc);
    double w[E::WSIZE > 0 ? E::WSIZE : 1];
    load_w<E>(a, e, cell, w);
    typename E::Geo g;
    E::prepare(g, xc, w, a.constants, lf);

    if constexpr (MODE == 0)
    {
      PosRegs<PosT, ND * ND> pos;
      if (a.pos)
        pos.load(a.pos, e);
#pragma unroll
      for (int i = 0; i < N; ++i)
      {
        if ((m0 >> i) & 1u)
          continue;
        double row[N];
        E::row(g, i, row);
        const int32_t r = d0[i / BS];
        const int64_t rb = a.row_ptr[r];
        const int64_t re = a.pos ? 0 : a.row_ptr[r + 1];
#pragma unroll
        for (int j = 0; j < N; ++j)
        {
          if ((m1 >> j) & 1u)
            continue;
          int64_t p;
          if (a.pos)
            p = rb + pos.get((i / BS) * ND + j / BS);
          else
          {
            p = find_col(a.cols, rb, re, d1[j / BS]);
            if (p < 0)
            {
              *a.err = 1;
              continue;
            }
          }
          red_add(a.values + p * (BS * BS) + (i % BS) * BS + (j % BS), row[j]);
        }
      }
    }
    else
    {
      double dv[N]; // alpha * (g - x0) on marked columns (assemble_vector_impl.h:377-402)
#pragma unroll
      for (int j = 0; j < N; ++j)
      {
        const int64_t jj = (int64_t)BS * d1[j / BS] + j % BS;
        dv[j] = ((m1 >> j) & 1u) ? a.alpha * (a.bc_values1[jj] - (a.x0 ? a.x0[jj] : 0.0)) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < N; ++i)
      {
        double row[N];
        E::row(g, i, row);
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j)
          acc = fma(row[j], dv[j], acc);
        if (acc != 0.0)
          red_add(a.b + (int64_t)BS * d0[i / BS] + i % BS, -acc);
      }
    }
  }
}

// generic thread-per-cell vector kernel
template <class E>
__global__ void __launch_bounds__(128) k_vector_cells(const AsmArgs a)
{
  constexpr int NX = E::NX, ND = E::ND, BS = E::BS, N = ND * BS;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += (int64_t)gridDim.x * blockDim.x)
  {
    int32_t cell;
    int lf = 0;
    if (a.entities)
    {
      cell = a.entities[2 * e];
      lf = a.entities[2 * e + 1];
    }
    else
      cell = a.cells ? a.cells[e] : (int32_t)e;
    int32_t d0[ND], xd[NX];
    load_ints<ND>(a.dofmap0 + (int64_t)cell * ND, d0);
    load_ints<NX>(a.x_dofmap + (int64_t)cell * NX, xd);
    double xc[NX][3];
    gather_coords<NX>(a.x, xd, xc);
    double w[E::WSIZE > 0 ? E::WSIZE : 1];
    load_w<E>(a, e, cell, w);
    typename E::Geo g;
    E::prepare(g, xc, w, a.constants, lf);
    double out[N];
    E::vec(g, out);
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (!E::FACET || out[i] != 0.0)
        red_add(a.b + (int64_t)BS * d0[i / BS] + i % BS, out[i]);
  }
}

// functional (rank 0): per-thread cell values, block reduction, one fp64 atomic per block
template <class E>
__global__ void __launch_bounds__(256) k_scalar_cells(const AsmArgs a, double* __restrict__ result)
{
  constexpr int NX = E::NX;
  double acc = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += (int64_t)gridDim.x * blockDim.x)
  {
    int32_t cell;
    int lf = 0;
    if (a.entities)
    {
      cell = a.entities[2 * e];
      lf = a.entities[2 * e + 1];
    }
    else
      cell = a.cells ? a.cells[e] : (int32_t)e;
    int32_t xd[NX];
    load_ints<NX>(a.x_dofmap + (int64_t)cell * NX, xd);
    double xc[NX][3];
    gather_coords<NX>(a.x, xd, xc);
    double w[E::WSIZE > 0 ? E::WSIZE : 1];
    load_w<E>(a, e, cell, w);
    typename E::Geo g;
    E::prepare(g, xc, w, a.constants, lf);
    acc += E::scalar(g);
  }
  __shared__ double part[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0)
    part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8)
  {
    double v = part[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1)
      v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0)
      atomicAdd(result, v);
  }
}

} // namespace bfx
