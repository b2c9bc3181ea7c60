/*
 * bfx.h — C-ABI of the B200-native DOLFINx hot path (libbfx.so).
 *
 * One CUDA context per process / rank / GPU.  Plain pointers and sizes only; no C++ or torch
 * types cross this boundary.  Every function returns an int status (0 = BFX_OK); the text of
 * the last error on the calling thread is available from bfx_last_error().  Nothing throws.
 * "dev" pointers are device pointers of the current device; "any" pointers may be host or
 * device (copied with cudaMemcpyDefault at plan-build time).  Hot calls take a stream
 * (cudaStream_t passed as void*; NULL = legacy default stream) and never allocate.
 *
 * Each entry point cites the reference interface it replaces; paths are relative to
 * /root/reference/cpp/dolfinx (DOLFINx 0.12.0.dev0).
 */
#ifndef BFX_H
#define BFX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define BFX_VERSION 100

  /* ---- status ---------------------------------------------------------------------------- */
  enum
  {
    BFX_OK = 0,
    BFX_ERR_CUDA = 1,            /* a CUDA runtime call failed (text in bfx_last_error) */
    BFX_ERR_INVALID = 2,         /* invalid argument */
    BFX_ERR_NOT_IN_SPARSITY = 3, /* "Entry not in sparsity" — la/matrix_csr_impl.h:93-94,161-162,223-224 */
    BFX_ERR_UNSUPPORTED = 4,     /* kernel id / block size / row length not supported */
    BFX_ERR_NCCL = 5,            /* NCCL unavailable or a NCCL call failed */
    BFX_ERR_NO_DEVICE = 6        /* no CUDA device: the product has no CPU fallback */
  };

  typedef void* bfx_stream_t;
  typedef struct bfx_csr bfx_csr_t;         /* la::MatrixCSR structure on device (la/MatrixCSR.h:67-70) */
  typedef struct bfx_asm bfx_asm_t;         /* per-(form integral) assembly plan */
  typedef struct bfx_scatter bfx_scatter_t; /* common::Scatterer plan on device (common/Scatterer.h:46-538) */
  typedef struct bfx_comm bfx_comm_t;       /* NCCL communicator (replaces MPI_Comm on the data path) */

  const char* bfx_last_error(void);
  const char* bfx_status_string(int status);
  int bfx_version(void);
  int bfx_device_count(int* count);
  int bfx_set_device(int device);

  /* ---- memory plumbing for non-torch hosts (C++ API mirror) ------------------------------ */
  int bfx_malloc(void** dev_ptr, size_t bytes);
  int bfx_free(void* dev_ptr);
  int bfx_memcpy(void* dst_any, const void* src_any, size_t bytes, bfx_stream_t stream); /* cudaMemcpyDefault, async */
  int bfx_memset(void* dev_ptr, int value, size_t bytes, bfx_stream_t stream);
  int bfx_stream_sync(bfx_stream_t stream);
  int bfx_host_alloc(void** host_ptr, size_t bytes); /* pinned */
  int bfx_host_free(void* host_ptr);

  /* ---- element kernels: the FFCx tabulate_tensor plug point -------------------------------
   * Replaces the std::function<tabulate_tensor> stored in fem::integral_data::kernel
   * (fem/Form.h:52-87, fem/kernel.h:18-20, fem/traits.h:28-30).  A kernel id selects a
   * hand-written sm_100a device function with the UFCx argument meaning
   * (A, w, c, coordinate_dofs, entity_local_index).                                          */
  enum
  {
    BFX_K_LAPLACE_P1_TRI_A = 0,     /* python/test/unit/fem/test_custom_jit_kernels.py:29-46 */
    BFX_K_SOURCE_P1_TRI_L = 1,      /* test_custom_jit_kernels.py:49-62 */
    BFX_K_MASS_COEFF_P1_TRI_A = 2,  /* inner(f*u, v)*dx, f P1 coefficient (test_ghost_mesh_assembly.py:50) */
    BFX_K_LOAD_COEFF_P1_TRI_L = 3,  /* inner(f, v)*dx */
    BFX_K_FACET_MASS_P1_TRI_A = 4,  /* inner(u, v)*ds */
    BFX_K_FACET_CONST_P1_TRI_L = 5, /* inner(c0, v)*ds */
    BFX_K_POISSON_P1_TET_A = 6,     /* kappa*inner(grad u, grad v)*dx (cpp/demo/poisson/poisson.py, 3-D) */
    BFX_K_LOAD_P1_TET_L = 7,        /* inner(f, v)*dx, f P1 */
    BFX_K_POISSON_P2_TET_A = 8,     /* cpp/test/poisson.py:16-27 */
    BFX_K_LOAD_P2_TET_L = 9,        /* inner(f, v)*dx, f P2 */
    BFX_K_ELASTICITY_Q1_HEX_A = 10, /* inner(sigma(u), grad(v))*dx, python/demo/demo_elasticity.py:131-150; c={mu,lambda} */
    BFX_K_LOAD_Q1_HEX_L = 11,       /* inner(f, v)*dx, f Q1 vector (bs=3) */
    BFX_K_FACET_LOAD_P1_TET_L = 12, /* inner(g, v)*ds, g P1 */
    BFX_K_FACET_MASS_P1_TET_A = 13, /* inner(u, v)*ds */
    /* 14 is reserved (the oracle's 2x2x2-Gauss elasticity variant has this number) */
    BFX_K_ACTION_POISSON_P1_TET_L = 15, /* action(a, ui), a = kappa*inner(grad u, grad v)*dx, P1: the matrix-free operator
                                           of cpp/demo/poisson_matrix_free (poisson.py: M = action(a, ui)); w = ui */
    BFX_K_ACTION_POISSON_P2_TET_L = 16, /* same, P2 */
    BFX_K_L2NORM2_P1_TET_M = 17,        /* functional inner(w, w)*dx, w P1 (the error functional E of that demo) */
    BFX_K_AVG_MASS_P1_TRI_DS = 18,      /* inner(avg(u), avg(v))*dS on interior facets, P1 triangles
                                           (test_ghost_mesh_assembly.py:104-122).  Interior-facet kernels work on MACRO
                                           cells: the plan is built on joint arrays [cell0 | cell1] of geometry nodes and
                                           dofs per facet (fem/assemble_matrix_impl.h:442-667), and the entity list of
                                           bfx_assemble_matrix_facets holds (facet, local_facet0 + 8 * local_facet1) */
    BFX_K_AVG_LOAD_P1_TRI_DS_L = 19,    /* conj(avg(v))*dS, P1 triangles (test_assembler.py:1003): macro-cell vector */
    BFX_K_ONE_TRI_DS_M = 20,            /* functional 1*dS (test_assemble_domains.py:203-210) */
    BFX_K_AVG2_COEFF_P1_TRI_DS_M = 21,  /* functional inner(avg(f), avg(f))*dS, f P1 (test_assemble_domains.py:225);
                                           w = [f on cell0 | f on cell1] (fem/pack.h:196-226) */
    BFX_K_COEFF2_P1_TRI_FACET_M = 22,   /* functional inner(f, f)*ds on exterior facets, f P1 (test_assemble_domains.py:224) */
    BFX_K_LOAD_PROD_P1_TET_L = 23,      /* inner(f*g, v)*dx, f and g P1: TWO coefficients in one integral,
                                           w = [f | g] at the form's coefficient offsets (fem/Form.h:593-604) */
    BFX_K_COUNT = 24
  };

  /* Static description of a kernel id: geometry nodes per cell, dofs per cell of test/trial space,
   * block size, rank (1 = vector, 2 = matrix), length of w and c it reads, 1 if facet kernel. */
  typedef struct
  {
    int nx, nd, bs, rank, w_size, c_size, facet;
  } bfx_kernel_info_t;
  int bfx_kernel_info(int kernel_id, bfx_kernel_info_t* info);

  /* Kernel plug point for forms the library does not ship (include/bfx_plugin.cuh): the caller compiles the library's
   * generic kernels for its own element struct and registers the launcher under an id >= BFX_K_USER_BASE.
   * launch(args, pos_bytes, mode, scalar_dev, stream): args = the library's kernel argument block, mode 0 matrix,
   * 1 lifting, 2 vector, 3 functional.  Registered ids run the cell-parallel fp64-RED strategy. */
  enum
  {
    BFX_K_USER_BASE = 1000
  };
  typedef int (*bfx_user_launch_t)(const void* args, int pos_bytes, int mode, double* scalar_dev, void* stream);
  int bfx_register_kernel(int kernel_id, const bfx_kernel_info_t* info, bfx_user_launch_t launch);

  /* ---- la::MatrixCSR structure ------------------------------------------------------------ */
  /* MatrixCSR(const SparsityPattern&) — la/MatrixCSR.h:628-703: copies graph, off_diag = row_ptr + nnz_diag.
   * row_ptr: int64[n_rows_all+1], cols: int32[nnz] (sorted per row), off_diag: int64[n_rows_all]. */
  int bfx_csr_create(bfx_csr_t** out, int32_t n_rows_all, int32_t n_rows_owned, const int64_t* row_ptr_any,
                     const int32_t* cols_any, const int64_t* off_diag_any, int bs0, int bs1);
  int bfx_csr_destroy(bfx_csr_t* csr);
  int64_t bfx_csr_nnz(const bfx_csr_t* csr);
  /* copy structure out (any of the destinations may be NULL) */
  int bfx_csr_get_structure(const bfx_csr_t* csr, int64_t* row_ptr_any, int32_t* cols_any, int64_t* off_diag_any);
  /* device pointers of the structure (valid for the life of csr) */
  int bfx_csr_device_ptrs(const bfx_csr_t* csr, const int64_t** row_ptr, const int32_t** cols,
                          const int64_t** off_diag);

  /* fem::create_sparsity_pattern + SparsityPattern::finalize, local part, on device.
   * fem/sparsitybuild.h:36-50 (cells) + la/SparsityPattern.cpp:438-478 (per-row dedup + sort,
   * off_diagonal_offsets = #cols < n_cols_owned).  Rows/cols are local block indices; extra
   * (row, col) pairs (entries received from other ranks, :394-423, or insert_diagonal) are merged in.
   * cells_any == NULL means cells [0, ncells). */
  int bfx_sparsity_build(bfx_csr_t** out, int32_t n_rows_all, int32_t n_rows_owned, int32_t n_cols_owned,
                         const int32_t* dofmap0_dev, int nd0, const int32_t* dofmap1_dev, int nd1,
                         const int32_t* cells_dev, int64_t ncells, const int32_t* extra_rows_any,
                         const int32_t* extra_cols_any, int64_t n_extra, int bs0, int bs1, bfx_stream_t stream);
  /* For ghost rows (rows >= n_rows_owned): de-duplicated column lists in INSERTION order (cell order,
   * then cell-local dof order) — what finalize() sends to the owners (SparsityPattern.cpp:326-351);
   * duplicates dropped (first occurrence kept), which leaves the owner-side result unchanged.
   * Two-phase: counts_host[n_ghost_rows] first (cols_host == NULL), then the packed columns. */
  int bfx_sparsity_ghost_rows(int32_t n_rows_all, int32_t n_rows_owned, const int32_t* dofmap0_dev, int nd0,
                              const int32_t* dofmap1_dev, int nd1, const int32_t* cells_dev, int64_t ncells,
                              int64_t* counts_host, int32_t* cols_host, bfx_stream_t stream);

  /* MatrixCSR::add / set with explicit blocks — la/MatrixCSR.h:265-335 → impl::insert_csr (kind 0),
   * insert_blocked_csr (kind 1), insert_nonblocked_csr (kind 2), la/matrix_csr_impl.h:67-232.
   * x: row-major (nr*dbs0 x nc*dbs1) values; op 0 = set, 1 = add. dbs = block size of the DATA. */
  int bfx_csr_insert(const bfx_csr_t* csr, double* values_dev, int kind, int dbs0, int dbs1, const double* x_any,
                     const int32_t* xrows_any, int nr, const int32_t* xcols_any, int nc, int op, bfx_stream_t stream);
  /* fem::set_diagonal — fem/assembler.h:644-686: values[row,row] = diag (SET) for unrolled dof rows. */
  int bfx_csr_set_diagonal(const bfx_csr_t* csr, double* values_dev, const int32_t* rows_unrolled_dev, int64_t n,
                           double diag, bfx_stream_t stream);
  /* MatrixCSR::squared_norm — la/MatrixCSR.h:473-486 (owned rows; caller reduces over ranks). */
  int bfx_csr_squared_norm(const bfx_csr_t* csr, const double* values_dev, double* result_host, bfx_stream_t stream);

  /* MatrixCSR::mult — la/MatrixCSR.h:877-946 → impl::spmv la/matrix_csr_impl.h:259-286.  y[0:n_owned*bs0] += A x.
   * part: 0 = whole rows, 1 = diagonal block [row_ptr, off_diag), 2 = off-diagonal block [off_diag, row_end). */
  enum
  {
    BFX_SPMV_FULL = 0,
    BFX_SPMV_DIAG = 1,
    BFX_SPMV_OFFDIAG = 2
  };
  int bfx_spmv(const bfx_csr_t* csr, const double* values_dev, const double* x_dev, double* y_dev, int part,
               bfx_stream_t stream);
  /* la::impl::local_transpose - la/mattrans.h:47-108: transpose of the block "owned rows x owned columns" (entries
   * row_ptr[i] .. off_diag_offset[i] of the owned rows).  Row j of the result lists the rows of A with an entry in
   * column j in ascending order, every bs0 x bs1 block is stored transposed; bit-exact with the reference loop.
   * row_ptrT (device, n_cols_owned + 1) is always written and the number of entries returned in *nnzT_host; colsT /
   * valsT (device, room for `capacity` entries / blocks; row_ptr[n_rows_owned] of A is always enough) may be NULL to
   * query the size only.  Synchronises the stream. */
  int bfx_csr_transpose_local(const bfx_csr_t* A, const double* values, int32_t n_cols_owned, int64_t* row_ptrT,
                              int32_t* colsT, double* valsT, int64_t capacity, int64_t* nnzT_host, bfx_stream_t stream);
  /* la::matmul, local part - impl::matmul, la/matmul.h:395-536 (block size 1; BFX_ERR_UNSUPPORTED otherwise, the
   * reference throws "Currently matmul only supports block size=1"): C = A B over the owned rows of A.  Columns of B
   * below n_owned_cols_b keep their index in C, its ghost columns are renumbered through b_ghost_remap (device, one
   * entry per ghost column of B); the rows of B behind ghost columns of A come as a CSR of fetched rows (device;
   * ghost_row_ptr has one entry more than A has ghost columns, columns already in the numbering of C) - the output of
   * impl::fetch_ghost_rows (:79-390), which stays host-side (dolfinx_b200.la.matrix_matmul_plan).  The rows are
   * bitwise the reference's: same order of additions, no entry for zero products or exact cancellations, sorted
   * columns.  begin() computes C into a workspace and returns its number of entries; end() copies row pointer
   * (n_rows_owned(A) + 1), per-row count of owned columns, columns and values into the caller's device arrays (NULL to
   * skip) and destroys the handle. */
  typedef struct bfx_matmul bfx_matmul_t;
  int bfx_csr_matmul_begin(const bfx_csr_t* A, const double* a_values, const bfx_csr_t* B, const double* b_values,
                           int32_t n_owned_cols_b, const int32_t* b_ghost_remap, const int64_t* ghost_row_ptr,
                           const int32_t* ghost_cols, const double* ghost_vals, int32_t n_owned_cols_c,
                           bfx_matmul_t** out, int64_t* nnz_host, bfx_stream_t stream);
  int bfx_csr_matmul_end(bfx_matmul_t* handle, int64_t* row_ptr, int32_t* off_diag, int32_t* cols, double* vals,
                         bfx_stream_t stream);
  /* bs = 1 SpMV kernel of this matrix: 0 = entry-consecutive stream, 1 = row per thread out of staged (cols, values),
   * 2 = row per thread fed by a two-stage TMA pipeline (persistent CTAs); -1 (default) = chosen by the average row
   * length on the first bfx_spmv call (the same kernel on every rank and in every run: results are reproducible);
   * -2 = time the three on the first bfx_spmv call and keep the fastest (which one wins depends on the numbering of
   * the matrix; not reproducible from run to run in the last bits). */
  int bfx_csr_set_spmv_variant(bfx_csr_t* csr, int variant);
  /* MatrixCSR::multT local kernels — la/MatrixCSR.h:950-1016 → impl::spmvT la/matrix_csr_impl.h:319-343 */
  int bfx_spmvT(const bfx_csr_t* csr, const double* values_dev, const double* x_dev, double* y_dev, int part,
                bfx_stream_t stream);

  /* ---- assembly ---------------------------------------------------------------------------- */
  /* Coefficient sources.  Either the reference's packed array (fem/pack.h:265-…: ncells x cstride,
   * row e belongs to the e-th entity of the integral) or a fused gather from the coefficient dof
   * vectors (pack_impl, fem/pack.h:77-103, done inside the element kernel): w[offset + bs*i + k] =
   * values[bs*dofmap[cell, i] + k]. */
  typedef struct
  {
    const double* values_dev;
    const int32_t* dofmap_dev;
    int nd, bs, offset;
  } bfx_coeff_src_t;
  typedef struct
  {
    const double* packed_dev; /* or NULL */
    int cstride;
    int n_fused; /* used when packed_dev == NULL */
    bfx_coeff_src_t fused[4];
  } bfx_coeffs_t;

  enum
  {
    BFX_ASM_ATOMIC = 0, /* cell-parallel, fp64 RED atomics into CSR through the cell->nnz map */
    /* 1 was reserved for a row-parallel gather in round 1; BFX_ASM_ROWGATHER took that role */
    BFX_ASM_CHUNKED = 2, /* chunk-aggregated: element matrices staged in shared memory, one update per distinct
                            CSR entry and chunk; needs bfx_asm_build_chunks() */
    BFX_ASM_ROWGATHER = 3 /* row-parallel gather with per-(row, cell) recomputation, every CSR value written once by
                             plain stores, no atomics, bitwise reproducible; BFX_K_ELASTICITY_Q1_HEX_A only;
                             needs bfx_asm_build_rowgather() */
  };
  enum
  {
    BFX_VALUES_ADD = 0,      /* reference semantics: values += contributions (assembler.h:497-498) */
    BFX_VALUES_OVERWRITE = 1 /* caller guarantees values are zero: fuses the zero-fill (gather mode) */
  };

  /* Plan for impl::assemble_cells_matrix / assemble_cells over one cell list
   * (fem/assemble_matrix_impl.h:92-200, fem/assemble_vector_impl.h:72-116): uploads / borrows
   * x_dofmap, dofmap0, dofmap1 and cells, and precomputes the cell -> CSR-position map
   * (replaces the per-entry std::lower_bound of insert_csr, la/matrix_csr_impl.h:92).
   * csr may be NULL for a linear form (then dofmap1 is ignored).  Returns BFX_ERR_NOT_IN_SPARSITY
   * where the reference would throw "Entry not in sparsity".
   * borrow != 0: the dofmap / cells pointers are device pointers that outlive the plan (no copy). */
  int bfx_asm_create(bfx_asm_t** out, const bfx_csr_t* csr, const int32_t* x_dofmap_any, int nx,
                     const int32_t* dofmap0_any, int nd0, const int32_t* dofmap1_any, int nd1,
                     int64_t ncells_all, const int32_t* cells_any, int64_t ncells, int32_t n_rows_all, int borrow,
                     bfx_stream_t stream);
  int bfx_asm_destroy(bfx_asm_t* plan);
  /* Plan of the chunk-aggregated strategy (BFX_ASM_CHUNKED) for the plan's cell list and MatrixCSR: cuts the
   * cell list into chunks, and for every chunk precomputes the distinct CSR destinations and the source
   * lists that sum the staged element matrices into them (the device-side replacement of the per-entry
   * lower_bound + "+=" of insert_csr, la/matrix_csr_impl.h:67-109).  x_dev (optional geometry, (N,3)) orders
   * the cells along a Morton curve of their centroids so that a chunk is a compact patch; NULL keeps the
   * given order.  BFX_ERR_UNSUPPORTED when the element matrix is too large for shared-memory staging. */
  enum
  {
    BFX_CHUNKS_TWO_STAGE = 8, /* symmetric P1-sized plans: the kernel leaves the per-destination sums in shared memory
                                 and writes the CSR values back in ADDRESS order (both entries of a symmetric pair),
                                 so that consecutive lanes update consecutive values; ignored where not applicable */
    BFX_CHUNKS_BANK_ORDER = 64, /* with BFX_CHUNKS_LINEAR_STAGING: order every source list so that the 16 lanes of a half
                                   warp read 16 different shared-memory banks per step where possible (no per-cell
                                   colour bytes, no extra instructions in the kernel) */
    BFX_CHUNKS_LEN_SORT = 128, /* with BFX_CHUNKS_LINEAR_STAGING on symmetric plans: destinations ordered by list length
                                  only (less padding of the 32-way interleaved lists); whether a destination is complete
                                  travels as one bit per lane in the group table; lean kernel only */
    BFX_CHUNKS_VECTOR = 65536, /* internal: the plan of a linear form (set by bfx_asm_build_chunks_vector) */
    BFX_CHUNKS_CB_SOFT = 131072, /* BFX_CHUNKS_CB(cells) is the lean kernel's preference, not a demand: a plan that cannot
                                   take the lean options (incomplete warp tables ...) is built with the element's default
                                   chunk size instead */
    BFX_CHUNKS_LEAN_ONLY = 262144, /* build the plan only if it can take the lean kernel's options; otherwise return
                                      BFX_ERR_UNSUPPORTED right after the warp tables, before any list is built (a caller
                                      with another scheme for that case does not pay the memory of a plan it will drop) */
    BFX_CHUNKS_PAD4 = 32, /* pad the source lists to multiples of 4 entries (no remainder steps in the list walk) */
    BFX_CHUNKS_TWO_STAGE_SPLIT = 16, /* with BFX_CHUNKS_TWO_STAGE: plain stores first, REDs after, each in address order */
    BFX_CHUNKS_SHARED_MATRIX = 4, /* the plan's cell list is a SUBSET of the cells that add to the matrix between its
                                     zero-fill and this plan's launch (boundary/interior split): an entry counts as
                                     complete only if EVERY cell of the dofmap that touches it lies in one chunk of this
                                     plan, so BFX_VALUES_OVERWRITE never overwrites another launch's contribution */
    BFX_CHUNKS_LINEAR_STAGING = 2, /* keep the padded linear staging layout (skip the bank colouring of the plan) */
    BFX_CHUNKS_SYMMETRIC = 1 /* symmetric bilinear form on one space (block size 1): stage the upper triangle of the
                                element matrix only and update the (i,j)/(j,i) CSR entries from one sum; a call whose
                                rows and columns do not share dofmap and bc markers is refused (BFX_ERR_UNSUPPORTED) */
  };
/* cells per chunk other than the element's default (256 for the P1 kernels, 128 for symmetric P2): 96, 128, 192 or 384
 * for the P1 kernels, 64 or 96 for symmetric P2 (other values: the default is used); OR-ed into the flags of
 * bfx_asm_build_chunks */
#define BFX_CHUNKS_CB(cells) ((((cells) / 32) & 0xff) << 8)
  int bfx_asm_build_chunks(bfx_asm_t* plan, const double* x_dev, int flags, bfx_stream_t stream);
  /* Plan of the row-gather strategy (BFX_ASM_ROWGATHER): the transposed dofmap (row -> incident (cell, local
   * node) pairs in ascending cell order, cf. fem::transpose_dofmap, fem/DofMap.h:62-64) and the per-call scratch of
   * the cell records.  BFX_ERR_UNSUPPORTED unless the plan is Q1 hexahedra x block size 3 with one dofmap. */
  int bfx_asm_build_rowgather(bfx_asm_t* plan, bfx_stream_t stream);
  /* Plan of the grouped vector strategy (bfx_assemble_vector_cells with BFX_ASM_CHUNKED) for linear-form plans of
   * P1-sized scalar elements (<= 4 nodes, <= 4 dofs per cell): Morton order of the cells (x_dev, optional), the
   * distinct nodes / dofs of every group of 32 cells and the list of (cell, local dof) pairs per distinct dof:
   * one warp assembles 32 cells and issues one RED per distinct dof instead of one per (cell, local dof). */
  int bfx_asm_build_groups(bfx_asm_t* plan, const double* x_dev, bfx_stream_t stream);
  /* Chunk plan of a LINEAR form for kernel_id (bfx_assemble_vector_cells with BFX_ASM_CHUNKED): element vectors staged
   * in shared memory per chunk of 256 / 384 Morton-ordered cells, one sum per distinct (dof, component) of the chunk, a
   * plain update for dofs whose cells all lie in the chunk and one RED per chunk-boundary dof
   * (fem/assemble_vector_impl.h:72-116 issues one += per (cell, local dof)).  Replaces a bfx_asm_build_groups plan.
   * BFX_ERR_UNSUPPORTED for kernels without the variant (facet kernels). */
  int bfx_asm_build_chunks_vector(bfx_asm_t* plan, const double* x_dev, int kernel_id, bfx_stream_t stream);
  int bfx_asm_chunk_stats(const bfx_asm_t* plan, int64_t* nchunks, int64_t* n_dest, int64_t* n_src_entries,
                          int64_t* plan_bytes);
  /* Staged entries whose shared-memory bank still collides with another entry read in the same half-warp step
   * after the plan's bank colouring (-1: linear staging layout in use). */
  int bfx_asm_chunk_bank_conflicts(const bfx_asm_t* plan, int64_t* n_conflicts);
  /* Kernel variant of a chunk plan (symmetric plans with the element's default chunk size; ignored elsewhere):
   * DEFAULT; OCC5 = compiled for 5 resident CTAs per SM (P1-sized elements; measured slower on B200); DIET = list walk
   * specialised per state space with rolled loops (round-2 experiment).  All give the same values. */
  enum
  {
    BFX_CHUNK_KERNEL_DEFAULT = 0,
    BFX_CHUNK_KERNEL_OCC5 = 1,
    BFX_CHUNK_KERNEL_DIET = 2,
    BFX_CHUNK_KERNEL_LEAN = 3, /* instruction-lean kernel for symmetric P1-sized plans built with BFX_CHUNKS_LINEAR_STAGING
                                  (falls back to DEFAULT where its preconditions do not hold) */
    BFX_CHUNK_KERNEL_WIDE = 4  /* elements with more than 36 staged entries per cell (P2): four threads per cell and the
                                  DIET list walk - twice the resident warps per SM at the same shared memory */
  };
  int bfx_asm_chunk_set_kernel(bfx_asm_t* plan, int variant);
  /* 1 if the chunk plan was built with the two-stage (address-ordered) write-back, see BFX_CHUNKS_TWO_STAGE */
  int bfx_asm_chunk_two_stage(const bfx_asm_t* plan, int* two_stage);

  /* impl::assemble_cells_matrix<false> — fem/assemble_matrix_impl.h:92-200 (+ bc row/col zeroing :161-196).
   * bc0/bc1: int8 markers of length bs*(owned+ghost) or NULL.  constants: host array. */
  int bfx_assemble_matrix_cells(const bfx_asm_t* plan, int kernel_id, const double* x_dev, const int8_t* bc0_dev,
                                const int8_t* bc1_dev, const bfx_coeffs_t* coeffs, const double* constants_host,
                                int n_constants, double* values_dev, int strategy, int values_mode,
                                bfx_stream_t stream);
  /* impl::assemble_cells — fem/assemble_vector_impl.h:72-116: b[bs*dof+k] += be[bs*i+k]. */
  int bfx_assemble_vector_cells(const bfx_asm_t* plan, int kernel_id, const double* x_dev,
                                const bfx_coeffs_t* coeffs, const double* constants_host, int n_constants,
                                double* b_dev, int strategy, bfx_stream_t stream);
  /* fem::assemble_scalar over cells — fem/assembler.h:173-213 -> impl::assemble_cells fem/assemble_scalar_impl.h:32-60:
   * *result_host = sum over the plan's cells of the functional kernel (rank 0); the caller reduces over ranks. */
  int bfx_assemble_scalar_cells(const bfx_asm_t* plan, int kernel_id, const double* x_dev, const bfx_coeffs_t* coeffs,
                                const double* constants_host, int n_constants, double* result_host,
                                bfx_stream_t stream);
  /* impl::lift_bc — fem/assemble_vector_impl.h:361-414 through assemble_cells_matrix<true>
   * (cell skip has_bc, assemble_matrix_impl.h:27-34,139-143): b -= alpha * Ae (bc_values1 - x0) on marked columns. */
  int bfx_lift_bc_cells(const bfx_asm_t* plan, int kernel_id, const double* x_dev, const bfx_coeffs_t* coeffs,
                        const double* constants_host, int n_constants, double* b_dev, const double* bc_values1_dev,
                        const int8_t* bc_markers1_dev, const double* x0_dev, double alpha, bfx_stream_t stream);
  /* impl::assemble_entities (exterior facets) — fem/assemble_matrix_impl.h:264-379,
   * fem/assemble_vector_impl.h:157-215.  entities: flat (cell, local_facet) int32 pairs on device;
   * packed coefficient rows (if any) are per entity.  values_dev / b_dev as above. */
  int bfx_assemble_matrix_facets(const bfx_asm_t* plan, int kernel_id, const double* x_dev,
                                 const int32_t* entities_dev, int64_t n_entities, const int8_t* bc0_dev,
                                 const int8_t* bc1_dev, const bfx_coeffs_t* coeffs, const double* constants_host,
                                 int n_constants, double* values_dev, bfx_stream_t stream);
  /* Chunk plans for the distributed overlap: bfx_asm_chunk_partition moves the chunks that hold a cell touching a row
   * >= n_owned_rows (a ghost row) to the front of the plan; *n_first = their number.  bfx_assemble_matrix_cells_part then
   * runs part 1 (those chunks), part 2 (all the others) or 0 (all) of the SAME plan - the chunk geometry (whole-cube bricks under the
   * Morton order) and the completeness of every destination stay those of the one-launch plan, which a plan over a
   * cell subset loses.  A caller assembles part 1, starts MatrixCSR::scatter_rev and assembles part 2 meanwhile. */
  int bfx_asm_chunk_partition(bfx_asm_t* plan, int32_t n_owned_rows, int64_t* n_first);
  int bfx_assemble_matrix_cells_part(bfx_asm_t* plan, int kernel_id, const double* x_dev, const int8_t* bc0_dev,
                                     const int8_t* bc1_dev, const bfx_coeffs_t* coeffs, const double* constants_host,
                                     int n_constants, double* values_dev, int values_mode, int part, bfx_stream_t stream);
  /* Row-range form of the row-gather strategy (plans with bfx_asm_build_rowgather): writes the CSR block rows
   * [row_begin, row_end) - each one complete from all cells of the plan, BFX_VALUES_OVERWRITE - and nothing else.
   * Calls on disjoint ranges add up to one bfx_assemble_matrix_cells call; a distributed caller assembles its ghost
   * rows first, starts MatrixCSR::scatter_rev (la/MatrixCSR.h:399-468) and assembles the owned rows meanwhile.
   * The range must be cut at multiples of bfx_asm_rowgather_tile_rows().  reuse_records != 0: the per-cell geometry
   * records of the previous call with the same x_dev are reused. */
  int bfx_assemble_matrix_rows(const bfx_asm_t* plan, int kernel_id, const double* x_dev, const int8_t* bc0_dev,
                               const int8_t* bc1_dev, const bfx_coeffs_t* coeffs, const double* constants_host,
                               int n_constants, double* values_dev, int32_t row_begin, int32_t row_end,
                               int reuse_records, bfx_stream_t stream);
  int bfx_asm_rowgather_tile_rows(const bfx_asm_t* plan, int* rows);
  /* impl::assemble_exterior_facets / assemble_interior_facets of a functional - fem/assemble_scalar_impl.h:78-168.
   * entities as above ((cell, local_facet), or (facet, lf0 + 8 lf1) on a macro-cell plan); the sum over the entities
   * is written to *result_host (the caller reduces over ranks). */
  int bfx_assemble_scalar_facets(const bfx_asm_t* plan, int kernel_id, const double* x_dev, const int32_t* entities_dev,
                                 int64_t n_entities, const bfx_coeffs_t* coeffs, const double* constants_host,
                                 int n_constants, double* result_host, bfx_stream_t stream);
  int bfx_assemble_vector_facets(const bfx_asm_t* plan, int kernel_id, const double* x_dev,
                                 const int32_t* entities_dev, int64_t n_entities, const bfx_coeffs_t* coeffs,
                                 const double* constants_host, int n_constants, double* b_dev, bfx_stream_t stream);
  /* pack_coefficient_entity — fem/pack.h:121-176: materialise the reference's packed layout on device
   * (parity with callers that pass packed coefficients). */
  int bfx_pack_coefficient(double* coeffs_dev, int cstride, int offset, const double* values_dev,
                           const int32_t* dofmap_dev, int nd, int bs, const int32_t* cells_dev,
                           const int32_t* entities_dev, int64_t n, bfx_stream_t stream);

  /* Host-buffer entry (drop-in call with the reference's host containers): copies x, the fused
   * coefficient dof vectors and bc markers to the device, assembles, copies the CSR values back.
   * All *_host arrays are host memory (pinned for full speed).  Scratch is owned by the plan. */
  int bfx_assemble_matrix_cells_host(bfx_asm_t* plan, int kernel_id, const double* x_host, int64_t n_x_nodes,
                                     const int8_t* bc0_host, const int8_t* bc1_host, int64_t n_bc,
                                     const double* coeff_values_host, int64_t n_coeff_values, int coeff_bs,
                                     const double* constants_host, int n_constants, double* values_host,
                                     int strategy, bfx_stream_t stream);
  /* The same entry split in two, so that a caller in a time loop can keep two steps in flight: begin() enqueues
   * host-to-device copies, zero-fill, assembly and the device-to-host copy of the values on one of the plan's two
   * internal streams and returns; end() waits for the OLDEST call in flight, after which its values_host is
   * complete.  With begin(k+1) issued before end(k) the device-to-host copy of step k (2 GB at C2, PCIe bound)
   * overlaps the uploads and kernels of step k+1 (the kernels of the two calls themselves run one after the other:
   * they share the plan's per-call scratch).  Host buffers must be pinned and must not be reused before the
   * matching end(); at most two calls may be in flight (a third begin() returns BFX_ERR_INVALID). */
  int bfx_assemble_matrix_cells_host_begin(bfx_asm_t* plan, int kernel_id, const double* x_host, int64_t n_x_nodes,
                                           const int8_t* bc0_host, const int8_t* bc1_host, int64_t n_bc,
                                           const double* coeff_values_host, int64_t n_coeff_values, int coeff_bs,
                                           const double* constants_host, int n_constants, double* values_host,
                                           int strategy);
  int bfx_assemble_matrix_cells_host_end(bfx_asm_t* plan);

  /* ---- Dirichlet conditions ----------------------------------------------------------------- */
  /* DirichletBC::mark_dofs — fem/DirichletBC.h:589-601 */
  int bfx_bc_mark(int8_t* markers_dev, const int32_t* dofs0_dev, int64_t n, bfx_stream_t stream);
  /* DirichletBC::set — fem/DirichletBC.h:495-578.  g_kind 0: Function, x[dofs0[i]] = alpha*(g[dofs_g[i]] - x0[dofs0[i]]);
   * g_kind 1: Constant, g[dofs0[i] % bs].  dofs_g NULL = dofs0; x0 NULL = no x0; dofs >= x_size skipped. */
  int bfx_bc_set(double* x_dev, int32_t x_size, const int32_t* dofs0_dev, const int32_t* dofs_g_dev, int64_t n,
                 const double* g_dev, int g_kind, int bs, const double* x0_dev, double alpha, bfx_stream_t stream);

  /* ---- la::Vector reductions — la/Vector.h:434-514 (local part; caller all-reduces) -------- */
  int bfx_dot(int64_t n, const double* x_dev, const double* y_dev, double* result_host, bfx_stream_t stream);
  int bfx_norm(int64_t n, const double* x_dev, int type /*0 l1, 1 l2^2, 2 linf*/, double* result_host,
               bfx_stream_t stream);
  int bfx_axpy(int64_t n, double alpha, const double* x_dev, double* y_dev, bfx_stream_t stream);
  /* x[i] = value - la::Vector::set / MatrixCSR::set(value) (la/Vector.h:205-208, la/MatrixCSR.h:239-241) */
  int bfx_fill(int64_t n, double value, double* x_dev, bfx_stream_t stream);

  /* ---- communicator (NCCL over NVLink; replaces MPI on the data path) ------------------------ */
  int bfx_comm_unique_id(char id_out[128]);
  int bfx_comm_create(bfx_comm_t** out, const char id[128], int rank, int size);
  int bfx_comm_destroy(bfx_comm_t* comm);
  int bfx_comm_rank(const bfx_comm_t* comm, int* rank, int* size);
  int bfx_comm_allreduce(bfx_comm_t* comm, double* buf_dev, int64_t n, int op /*0 sum, 1 max*/, bfx_stream_t stream);

  /* ---- common::Scatterer / la::Vector ghost exchange -------------------------------------------
   * Plan arrays are exactly the Scatterer members (common/Scatterer.h:65-198): local_inds / remote_inds
   * (already expanded by the block size), sizes/displs per neighbour (x bs), dest = ranks that ghost my
   * owned indices, src = ranks that own my ghosts.  comm may be NULL when n_dest = n_src = 0. */
  int bfx_scatter_create(bfx_scatter_t** out, bfx_comm_t* comm, const int32_t* local_inds_any, int64_t n_local,
                         const int32_t* remote_inds_any, int64_t n_remote, const int32_t* sizes_local,
                         const int32_t* displs_local, const int32_t* dest, int n_dest, const int32_t* sizes_remote,
                         const int32_t* displs_remote, const int32_t* src, int n_src);
  int bfx_scatter_destroy(bfx_scatter_t* sc);
  /* Vector::scatter_fwd_begin/end — la/Vector.h:219-281 → Scatterer::scatter_fwd_begin common/Scatterer.h:251-308:
   * pack owned values (kernel) → grouped ncclSend/ncclRecv on the plan's comm stream → unpack into the ghost tail.
   * x_dev holds [owned | ghosts]; n_owned = bs*size_local.  begin() returns immediately; work queued after
   * `stream` in program order; end() makes `stream` wait for the exchange and runs the unpack on it. */
  int bfx_scatter_fwd_begin(bfx_scatter_t* sc, const double* x_dev, bfx_stream_t stream);
  int bfx_scatter_fwd_end(bfx_scatter_t* sc, double* x_dev, int64_t n_owned, bfx_stream_t stream);
  /* Vector::scatter_rev_begin/end — la/Vector.h:314-379 → common/Scatterer.h:336-396; op 0 = set, 1 = add */
  int bfx_scatter_rev_begin(bfx_scatter_t* sc, const double* x_dev, int64_t n_owned, bfx_stream_t stream);
  int bfx_scatter_rev_end(bfx_scatter_t* sc, double* x_dev, int op, bfx_stream_t stream);

  /* MatrixCSR::scatter_rev_begin/end — la/MatrixCSR.h:399-468 with the plan built by the constructor
   * (:705-849): ghost_row_to_rank[n_ghost_rows] (index into src), val_send_disp[n_src+1], val_recv_disp[n_dest+1]
   * (both already x bs0*bs1), unpack_pos[val_recv_disp[n_dest]/bs2] (block positions in the owner's CSR). */
  typedef struct bfx_csr_scatter bfx_csr_scatter_t;
  int bfx_csr_scatter_create(bfx_csr_scatter_t** out, const bfx_csr_t* csr, bfx_comm_t* comm,
                             const int32_t* ghost_row_to_rank, int32_t n_ghost_rows, const int64_t* val_send_disp,
                             const int32_t* src, int n_src, const int64_t* val_recv_disp, const int32_t* dest,
                             int n_dest, const int64_t* unpack_pos_any);
  int bfx_csr_scatter_destroy(bfx_csr_scatter_t* plan);
  int bfx_csr_scatter_rev_begin(bfx_csr_scatter_t* plan, const double* values_dev, bfx_stream_t stream);
  int bfx_csr_scatter_rev_end(bfx_csr_scatter_t* plan, double* values_dev, bfx_stream_t stream);

  /* ---- host helpers for the synthetic fixtures (not on the timed path) ----------------------- */
  /* first-touch dof numbering, fem/dofmapbuilder.cpp:446-459: new_index[old] for old in [0, ndofs) */
  int bfx_host_first_touch_i32(const int32_t* dofmap, int64_t n, int32_t ndofs, int32_t* new_index);
  int bfx_host_first_touch_i64(const int64_t* dofmap, int64_t n, int64_t ndofs, int64_t* new_index);

#ifdef __cplusplus
}
#endif
#endif /* BFX_H */
