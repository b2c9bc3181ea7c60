// Assembly plan and kernel argument block.
#pragma once
#include "common.cuh"
#include "csr.cuh"

namespace bfx
{
struct CoefArgs
{
  const double* packed; // (n_entities, cstride) reference layout, or NULL
  int cstride;
  struct
  {
    const double* v;   // coefficient dof vector
    const int32_t* dm; // its dofmap (ncells_all x WND)
    int off;           // offset of this coefficient inside w (packed layout)
  } f[4];              // fused gather: the element's NCOEF coefficients in the order of w
};

struct AsmArgs
{
  const int32_t *x_dofmap, *dofmap0, *dofmap1;
  const int32_t* cells;    // cell list or NULL (= identity)
  const int32_t* entities; // (cell, local_facet) pairs or NULL
  int64_t n;               // number of cells / entities
  const unsigned long long* n_dev; // optional device-side count (<= n) produced by an earlier kernel
  const double* x;         // geometry, (N,3) row-major
  const int8_t *bc0, *bc1; // Dirichlet markers or NULL
  CoefArgs coef;
  double constants[8];
  // matrix sink
  double* values;
  const int64_t* row_ptr;
  const int32_t* cols;
  const void* pos; // cell -> position-in-row map, or NULL (binary search)
  // vector sink / lifting
  double* b;
  const double* bc_values1;
  const double* x0;
  double alpha;
  int* err;
  // RED kernels called for a row range of the matrix (bfx_assemble_matrix_rows): contributions to rows outside
  // [row_lo, row_hi) are dropped; row_hi == 0: no filter
  int32_t row_lo, row_hi;
};
} // namespace bfx

// Chunk-aggregated assembly plan (strategy BFX_ASM_CHUNKED, chunked.cu).  The cell list is cut into
// chunks of `cb` cells; the element matrices of a chunk are staged in shared memory and summed per
// DISTINCT CSR destination before anything leaves the SM.
struct ChunkHdr
{
  int64_t src_base32; // first 32-entry group of the chunk's source lists
  int64_t dest_base;  // first destination of the chunk (multiple of 32)
  int32_t n_dest;     // distinct destinations of the chunk
  int32_t n_complete; // the first n_complete receive every contribution from this chunk (plain update)
  int32_t n_src32;    // 32-entry groups of source-list entries
  int32_t pad;
};

struct bfx_chunks
{
  int cb = 0, n2 = 0; // cells per chunk, staged scalars per cell
  bool sym = false;   // symmetric form: destinations are (row i col j, row j col i) pairs
  int kernel_variant = 0; // BFX_CHUNK_KERNEL_*: which instantiation of the assembly kernel runs (bfx_asm_chunk_set_kernel)
  int64_t nchunks = 0, n_dest_pad = 0, n_src32 = 0;
  ChunkHdr* hdr = nullptr;
  uint32_t* winfo = nullptr;  // per group of 32 destinations: (offset of its source lists / 32) << 8 | list length
  void* dest_addr = nullptr;  // scalar index into values per destination
  // two-stage write-back (BFX_CHUNKS_TWO_STAGE): (address, destination rank | incomplete << 15) of every CSR value the
  // chunk updates, sorted by address; replaces dest_addr
  uint32_t* wr_addr = nullptr;
  uint16_t* wr_src = nullptr;
  int addr_bytes = 4;
  uint16_t* src = nullptr;    // source lists: shared-memory slots of the staged entries, 32-way interleaved
  // bank-conflict-free staging (small element matrices): slot of entry k of cell c = (k * cb/16 + c/16) * 16 +
  // colour[c][k]; the colours make the 16 lanes of every list step hit 16 different 8-byte banks
  uint8_t* colour = nullptr;  // ns (padded to a multiple of 4) bytes per cell slot, or NULL (padded linear layout)
  int64_t n_bank_conflicts = 0; // list reads the colouring could not make conflict free (diagnostic)
  bool vector_plan = false;     // chunk plan of a linear form (bfx_asm_build_chunks_vector)
  bool len_sorted = false;      // BFX_CHUNKS_LEN_SORT: winfo holds (info, completeness mask) pairs
  bool bank_ordered = false;    // BFX_CHUNKS_BANK_ORDER: the lists of the linear layout are ordered bank-aware
  int32_t* perm = nullptr;    // locality ordering of the plan's cell list (or NULL)
  // geometry dofmap / dofmaps in chunk order (one row per cell slot), so that phase 1 streams them
  // instead of chasing perm -> cells -> dofmap; NULL = the plan's own arrays are already in order
  int32_t *xdm = nullptr, *dm0 = nullptr, *dm1 = nullptr;
  // Warp tables: the distinct geometry nodes (wv_*) / dofs (wd_*) of each group of 32 consecutive cell slots
  // and, per cell, the positions of its nodes in that table.  When a group has <= 32 distinct nodes (the rule
  // on locality-ordered meshes) one lane loads one node and the cells pick theirs up with warp shuffles:
  // 3 coordinate loads per WARP instead of 3 NX per cell.  cnt == 0: the group gathers directly.
  int32_t *wv_ids = nullptr, *wd_ids = nullptr;
  uint8_t *wv_cnt = nullptr, *wd_cnt = nullptr;
  uint8_t *wv_loc = nullptr, *wd_loc = nullptr; // 4 * ceil(width / 4) bytes per slot
  int lean_dbg = 0;
  bool tables_complete = false; // every group of 32 cells has its node (and dof) table: no direct-gather groups
  // split of the chunks for the distributed overlap (bfx_asm_chunk_partition): the chunks with a cell that touches a row
  // >= part_rows (ghost rows) are MOVED to the front - part 1 = chunks [0, n_part1), part 2 = the others
  int64_t n_part1 = 0;
  uint32_t* part_list = nullptr; // plans of the classic kernels keep their chunk order: the chunks of part 1, ascending
  uint8_t* part_flag = nullptr;  // ... and 1 for every chunk of part 1
  int32_t part_rows = -1;
  int launch_part = 0; // which part the next launch runs (0 = all chunks); set by bfx_assemble_matrix_cells_part
  bool slim = false;            // reduced to what the lean kernel reads (perm, xdm, dm0, dm1, wv_cnt, wd_cnt freed)
  int nx = 0, nd0 = 0;          // geometry nodes / dofs per cell of the plan (table word counts)
  // Vector plans (bfx_asm_build_groups): for every group of 32 cell slots, the (cell, local dof) pairs that
  // land on each distinct dof of the group: goff[g][l] .. goff[g][l+1] index glist[g][], entries = i * 32 + lane
  uint8_t *glist = nullptr, *goff = nullptr;
  int gstride = 0; // bytes of glist per group (32 * nd0)
  // Dirichlet markers of the current call packed to one bit per dof (rebuilt by every call: 1/8 of the
  // marker bytes, so the per-cell lookups of phase 1 stay in L1/L2)
  uint32_t *bits0 = nullptr, *bits1 = nullptr;
  int64_t n_dofs0 = 0, n_dofs1 = 0; // scalar dofs indexed through dofmap0 / dofmap1
};

// Row-gather plan of the Q1 elasticity kernel (rowgather.cu)
struct bfx_rowgather
{
  int64_t* tptr = nullptr;   // transposed dofmap: row -> incident (entity, local node) entries
  uint32_t* tent = nullptr;  // (entity << 3) | local node
  double* rec = nullptr;     // per-call cell records (K, |det J|)
  int32_t* na_cells = nullptr;           // per-call list of cells that are not parallelepipeds
  unsigned long long* na_count = nullptr;
  // block-gather plan (round 2, k_q1_blockgather): tiles of BG_ROWS consecutive block rows; per tile the distinct
  // incident cells, per CSR block the list of its (cell slot, i, j) contributions, blocks ranked by list length
  int32_t* bg_cells = nullptr;  // [tile][BG_CCAP]: distinct cells (entity indices) of the tile, ascending
  void* bg_tiles = nullptr;     // [tile] BGTile header: first block, counts, blocks per contribution level
  uint16_t* bg_perm = nullptr;  // [tile][BG_NBCAP]: rank -> block in tile | row in tile << 10
  uint16_t* bg_ent = nullptr;   // [tile][8 BG_INCCAP]: contributions (slot << 6 | i << 3 | j), level-major
  uint8_t* bg_zcb = nullptr;    // [tile][BG_NBCAP]: per-call Dirichlet mask of every block's column node
  uint8_t *bg_mask0 = nullptr, *bg_mask1 = nullptr; // per-call Dirichlet masks per node (bit k = component k)
  int64_t bg_n_nodes = 0;
  uint32_t* bg_cmask = nullptr; // per-call column masks per plan cell (3 bits per local node)
  int64_t bg_ntiles = 0;
  int bg_max_cells = 0, bg_max_blocks = 0, bg_max_steps = 0, bg_max_q = 0; // over all tiles (sizes the kernel's shared memory)
  bool bg_ok = false;
};

struct bfx_asm
{
  const bfx_csr* csr = nullptr;
  int nx = 0, nd0 = 0, nd1 = 0;
  int64_t ncells_all = 0, ncells = 0;
  int32_t n_rows_all = 0;
  bool owns = true;
  int32_t *x_dofmap = nullptr, *dofmap0 = nullptr, *dofmap1 = nullptr, *cells = nullptr;
  // cell -> CSR position map: per cell nd0*nd1 offsets relative to row_ptr[row], uint8 or uint16
  char* pos = nullptr;
  int pos_bytes = 1, pos_stride = 0;
  // scratch of the host-buffer entry point
  // (two slots: a call may be in flight on each, so that the device-to-host copy of one step overlaps the
  // host-to-device copy and the kernels of the next: bfx_assemble_matrix_cells_host_begin / _end)
  struct HostSlot
  {
    double *x = nullptr, *coeff = nullptr, *values = nullptr;
    int8_t *bc0 = nullptr, *bc1 = nullptr;
    int64_t x_n = 0, coeff_n = 0, bc_n = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr, computed = nullptr; // all work of the call / its kernels only
    bool busy = false, used = false;
  } hs[2];
  int hs_next = 0; // slot the next begin() uses; end() waits for the oldest busy slot
  bfx_chunks* chunks = nullptr;
  bfx_rowgather* rowgather = nullptr;
  // lifting scratch: cells with a Dirichlet column dof (rebuilt by every call)
  int32_t* lift_cells = nullptr;
  unsigned long long* lift_count = nullptr;
};

namespace bfx
{
// chunked.cu
int chunked_supported(int kernel_id);
int launch_chunked(const bfx_asm* P, int kernel_id, const AsmArgs& a, int values_mode, cudaStream_t st);
void free_chunks(bfx_chunks* c);
int grouped_vector_supported(int kernel_id);
int chunked_vector_cells(int kernel_id);
int launch_vector_grouped(const bfx_asm* P, int kernel_id, const AsmArgs& a, cudaStream_t st);
// rowgather.cu
int launch_rowgather_q1(const bfx_asm* P, const AsmArgs& a, int values_mode, cudaStream_t st, int32_t row_begin = 0,
                        int32_t row_end = -1, bool reuse_records = false);
int rowgather_tile_rows(const bfx_asm* P);
void free_rowgather(bfx_rowgather* g);
// assemble.cu: the cell-parallel fp64-RED Q1 elasticity kernel (used for non-affine cells by the row-gather path)
int launch_q1_red(const bfx_asm* P, const AsmArgs& a, cudaStream_t st);
} // namespace bfx
