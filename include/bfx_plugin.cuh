/* bfx_plugin.cuh - the kernel plug point of libbfx.so for a form the library does not ship.
 *
 * Replaces the tabulate_tensor function pointer of fem::integral_data (fem/kernel.h:18-20, fem/Form.h:76-78).  A host
 * function pointer cannot run on the device, so the caller writes the element as a struct with the conventions of
 * dolfinx_b200/csrc/elements.cuh (the UFCx argument meaning: A / b, w, c, coordinate_dofs, entity_local_index), compiles
 * ONE translation unit with nvcc for sm_100a that instantiates the library's generic cell-parallel kernels for it, and
 * registers the resulting launcher under a kernel id >= BFX_K_USER_BASE:
 *
 *     #include <bfx_plugin.cuh>
 *     struct MassP1Tet : bfx::el::TetBase { ... prepare(...), row(...) ... };     // see elements.cuh
 *     BFX_PLUGIN_KERNEL(MassP1Tet, my_mass_p1_tet)                                // defines my_mass_p1_tet{,_info}
 *
 *     nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -I include plugin.cu -o libplugin.so
 *
 *     bfx_kernel_info_t info;  my_mass_p1_tet_info(&info);
 *     bfx_register_kernel(BFX_K_USER_BASE + 0, &info, my_mass_p1_tet);
 *
 * after which the id works in bfx_assemble_matrix_cells / _facets (fp64-RED strategy), bfx_lift_bc_cells,
 * bfx_assemble_vector_cells / _facets and bfx_assemble_scalar_cells / _facets like a built-in one.  libbfx.so is not
 * rebuilt.  The launcher runs in the caller's library: it must be compiled against the same headers as the libbfx.so
 * it registers with (bfx_version()). */
#pragma once
#include "../dolfinx_b200/csrc/cell_kernels.cuh"
#include "../dolfinx_b200/csrc/elements.cuh"
#include "bfx.h"

namespace bfx
{
/* mode: 0 assemble matrix (fp64 REDs, bc rows / columns zeroed), 1 lifting, 2 vector, 3 functional (partial sums are
 * added to *scalar_dev).  pos_bytes: width of the plan's position map entries (1 or 2), used when args->pos is set. */
template <class E>
int plugin_launch(const AsmArgs* args, int pos_bytes, int mode, double* scalar_dev, cudaStream_t st)
{
  const AsmArgs& a = *args;
  if (a.n <= 0)
    return BFX_OK;
  const long long want = (a.n + 127) / 128;
  const unsigned grid = (unsigned)(want < 148LL * 64 ? want : 148LL * 64);
  if constexpr (E::RANK == 2)
  {
    if (mode == 0)
    {
      if (pos_bytes == 2 && a.pos)
        k_matrix_cells<E, uint16_t, 0><<<grid, 128, 0, st>>>(a);
      else
        k_matrix_cells<E, uint8_t, 0><<<grid, 128, 0, st>>>(a);
    }
    else if (mode == 1)
      k_matrix_cells<E, uint8_t, 1><<<grid, 128, 0, st>>>(a);
    else
      return BFX_ERR_INVALID;
  }
  else if constexpr (E::RANK == 1)
  {
    if (mode != 2)
      return BFX_ERR_INVALID;
    k_vector_cells<E><<<grid, 128, 0, st>>>(a);
  }
  else
  {
    if (mode != 3 || !scalar_dev)
      return BFX_ERR_INVALID;
    k_scalar_cells<E><<<(grid + 1) / 2, 256, 0, st>>>(a, scalar_dev);
  }
  return cudaGetLastError() == cudaSuccess ? BFX_OK : BFX_ERR_CUDA;
}

template <class E>
void plugin_info(bfx_kernel_info_t* info)
{
  info->nx = E::NX, info->nd = E::ND, info->bs = E::BS, info->rank = E::RANK;
  info->w_size = E::WSIZE, info->c_size = E::CSIZE, info->facet = E::FACET ? 1 : 0;
}
} // namespace bfx

#define BFX_PLUGIN_KERNEL(E, symbol)                                                                                   \
  extern "C" int symbol(const void* args, int pos_bytes, int mode, double* scalar_dev, void* stream)                   \
  {                                                                                                                    \
    return bfx::plugin_launch<E>(static_cast<const bfx::AsmArgs*>(args), pos_bytes, mode, scalar_dev,                  \
                                 static_cast<cudaStream_t>(stream));                                                   \
  }                                                                                                                    \
  extern "C" void symbol##_info(bfx_kernel_info_t* info) { bfx::plugin_info<E>(info); }
