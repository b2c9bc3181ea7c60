// Chunk-aggregated matrix assembly (strategy BFX_ASM_CHUNKED) for
// fem::impl::assemble_cells_matrix (fem/assemble_matrix_impl.h:92-200) + MatrixCSR::add
// (la/MatrixCSR.h:310-335, la/matrix_csr_impl.h:67-109).
//
// Why: a fp64 RED costs 1.2-1.9 SM-cycles per lane even when it hits L2
// (profiles/r01_microbench_red_bulk_lds.txt), a shared-memory load 0.07-0.2.  A P1 mesh sends 16
// contributions per cell to ~2.5 distinct CSR entries per cell, so the sums are formed on the SM:
//   phase 1  one thread per cell of a chunk of CB cells (coordinates and marker bits through the warp
//            tables: one load per distinct node of 32 cells + shuffles): element matrix in registers,
//            bc rows/columns zeroed, entries stored to shared memory; symmetric forms stage the upper
//            triangle only;
//   phase 2  one thread per DISTINCT destination of the chunk (a pair of entries (i,j)/(j,i) for
//            symmetric forms): sums its contributions from shared memory through a precomputed source
//            list; destinations that receive all their contributions from this chunk are updated with a
//            plain load/add/store (or a plain store when the caller guarantees zeroed values), the
//            others (chunk boundary) with one RED.
// The plan (bfx_asm_build_chunks) orders the cells along a Morton curve of their centroids when the
// geometry is given, so a chunk is a compact patch and most destinations are complete.  Destinations
// are sorted by (complete, list length, address): the 32 lanes of a warp walk lists of equal length,
// the lists are stored 32-way interleaved and reach shared memory by TMA bulk copies while phase 1 runs.
// The staging slots are bank-coloured by the plan (k_chunk_colour): stores and list reads are conflict
// free.  DESIGN.md section 4.2 has the measured effect of every step.
// The same file holds the grouped vector kernel (bfx_asm_build_groups) that reuses the warp tables.
#include "asm_device.cuh"
#include <algorithm>
#include "elements.cuh"
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

using namespace bfx;

namespace
{
constexpr int PLAN_THREADS = 1024;
constexpr int MAX_LIST = 255;

// cells per chunk for ns staged scalars per cell (N^2, or N(N+1)/2 for a symmetric form): the staging
// area fits 40-60 KB of shared memory and the plan kernel's 1024 x ITEMS keys
constexpr int chunk_cb(int ns) { return ns <= 36 ? 256 : (ns <= 64 ? 128 : (ns <= 100 ? 64 : 0)); }
// threads per chunk: one per cell for small element matrices, several per cell (rows dealt round-robin) else
constexpr int chunk_threads(int ns) { return 256; }
// alternative chunk sizes (bfx_asm_build_chunks flags, BFX_CHUNKS_CB(cells)): under the Morton order of a Kuhn box
// (6 tetrahedra per cube, all six contiguous) 256 or 128 cells cut the chunks through the cubes (42.7 / 21.3 cubes);
// 96 / 192 / 384 cells are 2x2x4 / 4x4x2 / 4x4x4 bricks of whole cubes with fewer chunk-boundary destinations per
// cell (P1 RED pairs per cell: 96: 1.61, 128: 1.80, 192: 1.25, 256: 1.38, 384: 0.96; measured times in DESIGN.md 4.2)
constexpr bool chunk_cb_supported(int ns, int cb)
{
  return ns <= 16 ? (cb == 96 || cb == 128 || cb == 192 || cb == 384) : (ns == 55 ? (cb == 64 || cb == 96) : false);
}
constexpr int chunk_threads_for(int ns, int cb) { return ns <= 16 ? cb : (cb == 96 ? 288 : 256); }
constexpr int chunk_min_ctas(int threads)
{
  return threads <= 96 ? 10 : (threads <= 128 ? 8 : (threads <= 192 ? 5 : (threads <= 256 ? 4 : (threads <= 384 ? 3 : 2))));
}
// largest element (staged scalars per cell) that gets the bank-coloured staging layout: the colours of a cell
// travel in registers
constexpr int COLOUR_MAX_NS = 36; // (measured: P2 Poisson, 55 staged entries, is bound by its chunk-boundary REDs, not by bank conflicts)
constexpr int staged_per_cell(int n, bool sym) { return sym ? n * (n + 1) / 2 : n * n; }
// 32-entry groups of list entries a chunk may have to be staged in shared memory (lists are padded per group of
// 32 destinations to the longest list of the group: ~1.4x the entries on tetrahedral meshes)
constexpr int src_group_cap(int nst) { return nst <= 4096 ? (nst / 32) * 5 / 2 + 2 : (nst / 32) * 3 / 2 + 2; }

// ---------------------------------------------------------------------------------------------
// plan construction
// ---------------------------------------------------------------------------------------------
struct ChunkBuildArgs
{
  int64_t n; // entities of the plan
  int cb, nd0, nd1, bs0, bs1, ns; // ns: staged scalars per cell
  int sym;                        // symmetric form: one staged entry and one destination PAIR per {i, j}
  const int32_t *perm, *cells, *dofmap0;
  const int64_t* row_ptr;
  const char* pos;
  int pos_stride, pos_bytes;
  const int32_t* total; // contributions per block entry over the whole cell list
  int addr_bits;
  int64_t *o_ndw, *o_nsrc32;                // pass A: per chunk
  const int64_t *dest_base32, *src_base32;  // pass B: exclusive scans of the above
  ChunkHdr* hdr;
  uint32_t* winfo;
  void* dest_addr;
  int addr_bytes;
  uint16_t* src;
  int* err;
  // two-stage write-back (BFX_CHUNKS_TWO_STAGE): every (address, destination rank) pair of the chunk - both entries
  // of a symmetric pair - sorted by address, so that the lanes of the write-back pass hit consecutive CSR values
  int len_sort;        // BFX_CHUNKS_LEN_SORT: destinations ordered by list length only; completeness travels as a bit mask
  int pad4;            // BFX_CHUNKS_PAD4: list lengths (per group of 32 destinations) padded to multiples of 4
  int two, dcap;       // dcap: destinations of a chunk the kernel holds in shared memory
  uint32_t* wr_addr;   // 2 * n_dest_pad entries, chunk q at 2 * dest_base
  uint16_t* wr_src;    // destination rank | 0x8000 if the destination is incomplete (RED)
};

__device__ __forceinline__ uint32_t pos_at(const char* pos, int pos_stride, int pos_bytes, int64_t e, int t)
{
  const char* row = pos + e * pos_stride;
  return pos_bytes == 1 ? (uint32_t) reinterpret_cast<const uint8_t*>(row)[t]
                        : (uint32_t) reinterpret_cast<const uint16_t*>(row)[t];
}

// scalar index into values of staged entry idx = k * cb + c of chunk q, or -1.  General form: k = I * N1 + J.
// Symmetric form (bs = 1, one dofmap): k ranks the pair i <= j; the key is the smaller of the addresses of
// (row i, col j) and (row j, col i), `other` the larger (== key on the diagonal): both receive the same sum.
__device__ __forceinline__ int64_t contrib_addr(const ChunkBuildArgs& p, int64_t q, int idx, int64_t& blk, int64_t& other)
{
  const int c = idx % p.cb, k = idx / p.cb;
  const int64_t slot = q * p.cb + c;
  if (slot >= p.n)
    return -1;
  const int64_t e = p.perm ? p.perm[slot] : slot;
  const int32_t cell = p.cells ? p.cells[e] : (int32_t)e;
  if (p.sym)
  {
    int i = 0, rem = k;
    while (rem >= p.nd0 - i)
    {
      rem -= p.nd0 - i;
      ++i;
    }
    const int j = i + rem;
    const int64_t a1 = p.row_ptr[p.dofmap0[(int64_t)cell * p.nd0 + i]] + pos_at(p.pos, p.pos_stride, p.pos_bytes, e, i * p.nd1 + j);
    const int64_t a2 = p.row_ptr[p.dofmap0[(int64_t)cell * p.nd0 + j]] + pos_at(p.pos, p.pos_stride, p.pos_bytes, e, j * p.nd1 + i);
    blk = a1 < a2 ? a1 : a2;
    other = a1 < a2 ? a2 : a1;
    return blk;
  }
  const int N1 = p.nd1 * p.bs1;
  const int I = k / N1, J = k - I * N1;
  const int i = I / p.bs0, a = I - i * p.bs0, j = J / p.bs1, b = J - j * p.bs1;
  const int32_t r = p.dofmap0[(int64_t)cell * p.nd0 + i];
  blk = p.row_ptr[r] + pos_at(p.pos, p.pos_stride, p.pos_bytes, e, i * p.nd1 + j);
  other = blk * (p.bs0 * p.bs1) + a * p.bs1 + b;
  return other;
}

// contributions per block entry (how many (cell, i, j) land on it)
__global__ void k_count_contrib(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ dofmap0,
                                int nd0, int nd1, const int64_t* __restrict__ row_ptr, const char* __restrict__ pos,
                                int pos_stride, int pos_bytes, int32_t* __restrict__ total)
{
  const int64_t work = n * nd0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t e = t / nd0;
    const int i = (int)(t - e * nd0);
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    const int64_t rb = row_ptr[dofmap0[(int64_t)cell * nd0 + i]];
    for (int j = 0; j < nd1; ++j)
      atomicAdd(total + rb + pos_at(pos, pos_stride, pos_bytes, e, i * nd1 + j), 1);
  }
}

// the same count over EVERY cell of the dofmap, entries located by binary search (cells outside the plan's list
// have no position map): used when other launches add to the same matrix (BFX_CHUNKS_SHARED_MATRIX)
__global__ void k_count_contrib_all(int64_t ncells_all, const int32_t* __restrict__ dofmap0, int nd0,
                                    const int32_t* __restrict__ dofmap1, int nd1, const int64_t* __restrict__ row_ptr,
                                    const int32_t* __restrict__ cols, int32_t* __restrict__ total)
{
  const int64_t work = ncells_all * nd0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t cell = t / nd0;
    const int i = (int)(t - cell * nd0);
    const int32_t r = dofmap0[cell * nd0 + i];
    const int64_t rb = row_ptr[r], re = row_ptr[r + 1];
    for (int j = 0; j < nd1; ++j)
    {
      const int64_t p = find_col(cols, rb, re, dofmap1[cell * nd1 + j]);
      if (p >= 0)
        atomicAdd(total + p, 1);
    }
  }
}

template <int ITEMS, bool WRITE>
__global__ void __launch_bounds__(PLAN_THREADS) k_chunk_plan(const ChunkBuildArgs p)
{
  using Sort1 = cub::BlockRadixSort<uint64_t, PLAN_THREADS, ITEMS>;
  using Sort2 = cub::BlockRadixSort<uint32_t, PLAN_THREADS, ITEMS, uint16_t>;
  using Scan = cub::BlockScan<int, PLAN_THREADS>;
  constexpr int NK = PLAN_THREADS * ITEMS;
  union Temp
  {
    typename Sort1::TempStorage s1;
    typename Sort2::TempStorage s2;
    typename Scan::TempStorage sc;
  };
  extern __shared__ __align__(16) unsigned char raw[];
  Temp& temp = *reinterpret_cast<Temp*>(raw);
  uint64_t* lastkey = reinterpret_cast<uint64_t*>(raw + ((sizeof(Temp) + 15) / 16) * 16);
  uint32_t* mw = reinterpret_cast<uint32_t*>(lastkey + PLAN_THREADS);
  uint32_t* woff = mw + PLAN_THREADS;
  uint32_t* cmask = woff + PLAN_THREADS;
  uint16_t* sidx = reinterpret_cast<uint16_t*>(cmask + PLAN_THREADS);
  uint16_t* dstart = sidx + NK;

  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;
  const int ncontrib = p.ns * p.cb;
  const uint16_t ZERO = (uint16_t)ncontrib;

  // ---- 1. keys (address << 16 | staged index), sorted by address
  uint64_t keys[ITEMS];
#pragma unroll
  for (int it = 0; it < ITEMS; ++it)
  {
    const int idx = tid * ITEMS + it;
    int64_t blk, other, addr = -1;
    if (idx < ncontrib)
      addr = contrib_addr(p, q, idx, blk, other);
    keys[it] = addr >= 0 ? ((uint64_t)addr << 16) | (uint64_t)idx : ~0ull;
  }
  Sort1(temp.s1).Sort(keys, 16, 16 + p.addr_bits);
  __syncthreads();

  // ---- 2. distinct destinations: heads of equal-address runs
  lastkey[tid] = keys[ITEMS - 1];
  mw[tid] = 0;
  cmask[tid] = 0;
  __syncthreads();
  int nheads = 0, nvalid = 0;
  bool head[ITEMS];
  {
    uint64_t prev = tid ? lastkey[tid - 1] : ~0ull;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it)
    {
      const bool valid = keys[it] != ~0ull;
      head[it] = valid && ((tid == 0 && it == 0) || (keys[it] >> 16) != (prev >> 16));
      nheads += head[it];
      nvalid += valid;
      prev = keys[it];
    }
  }
  int dbase, n_dest, vbase, n_valid;
  Scan(temp.sc).ExclusiveSum(nheads, dbase, n_dest);
  __syncthreads();
  Scan(temp.sc).ExclusiveSum(nvalid, vbase, n_valid);
  __syncthreads();
  {
    int d = dbase;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it)
    {
      const int at = tid * ITEMS + it;
      if (keys[it] != ~0ull)
        sidx[at] = (uint16_t)(keys[it] & 0xffffu);
      if (head[it])
        dstart[d++] = (uint16_t)at;
    }
    if (tid == 0)
      dstart[n_dest] = (uint16_t)n_valid; // NK <= 65535 is checked on the host
  }
  __syncthreads();

  // ---- 3. order destinations by (incomplete, list length descending); radix sort is stable, so
  //         the address order survives inside each class
  uint32_t k2[ITEMS];
  uint16_t v2[ITEMS];
  int ncomp = 0;
#pragma unroll
  for (int it = 0; it < ITEMS; ++it)
  {
    const int d = tid * ITEMS + it;
    k2[it] = 0x3ffu;
    v2[it] = 0xffffu;
    if (d < n_dest)
    {
      const int st = dstart[d];
      int cnt = (int)dstart[d + 1] - st;
      if (cnt > MAX_LIST)
      {
        *p.err = 3;
        cnt = MAX_LIST;
      }
      int64_t blk, other;
      contrib_addr(p, q, sidx[st], blk, other);
      const bool complete = p.total[blk] == cnt;
      ncomp += complete;
      k2[it] = ((complete && !p.len_sort ? 0u : 1u) << 8) | (uint32_t)(MAX_LIST - cnt);
      v2[it] = (uint16_t)d;
    }
  }
  Sort2(temp.s2).Sort(k2, v2, 0, 10);
  __syncthreads();
  int cb_, n_complete;
  Scan(temp.sc).ExclusiveSum(ncomp, cb_, n_complete);
  __syncthreads();

  // ---- 4. list length of every group of 32 destinations and the offsets of their lists
#pragma unroll
  for (int it = 0; it < ITEMS; ++it)
  {
    const int r = tid * ITEMS + it;
    if (r < n_dest)
      atomicMax(&mw[r >> 5], (uint32_t)(MAX_LIST - (int)(k2[it] & 0xffu)));
  }
  __syncthreads();
  const int n_dw = (n_dest + 31) >> 5;
  int my_m = tid < n_dw ? (int)mw[tid] : 0, my_off, n_src32;
  if (p.pad4 && tid < n_dw && my_m <= 252)
  {
    my_m = (my_m + 3) & ~3; // the padding entries point at the zero slot; 252: the length travels in 8 bits
    mw[tid] = (uint32_t)my_m; // (read by the list writer after the next barrier)
  }
  if (n_dw > PLAN_THREADS) // cannot happen: NK / 32 <= 1024 for ITEMS <= 32
    *p.err = 4;
  Scan(temp.sc).ExclusiveSum(my_m, my_off, n_src32);
  woff[tid] = (uint32_t)my_off;
  __syncthreads();

  if constexpr (!WRITE)
  {
    if (tid == 0)
    {
      p.o_ndw[q] = (n_dw + 3) & ~3; // 16-byte aligned winfo / 128-destination aligned chunks (TMA)
      p.o_nsrc32[q] = n_src32;
      if (p.two && (n_dest > p.dcap || 2 * n_dest > NK))
        atomicMax(p.err, 5); // the host drops the two-stage write-back and goes on
    }
    return;
  }
  else
  {
    const int64_t dest_base = p.dest_base32[q] << 5;
    const int64_t src_base32 = p.src_base32[q];
    if (tid == 0)
    {
      ChunkHdr h;
      h.src_base32 = src_base32;
      h.dest_base = dest_base;
      h.n_dest = n_dest;
      h.n_complete = n_complete;
      h.n_src32 = n_src32;
      h.pad = 0;
      p.hdr[q] = h;
    }
    if (tid < n_dw && !p.len_sort)
      p.winfo[(dest_base >> 5) + tid] = (woff[tid] << 8) | mw[tid];
    if (p.len_sort)
    {
      // completeness of every destination as one bit per lane of its group
#pragma unroll
      for (int it = 0; it < ITEMS; ++it)
      {
        const int r = tid * ITEMS + it;
        if (r < n_dest)
        {
          const int d = v2[it];
          const int st = dstart[d];
          const int cnt = min((int)dstart[d + 1] - st, MAX_LIST);
          int64_t blk, other;
          contrib_addr(p, q, sidx[st], blk, other);
          if (p.total[blk] == cnt)
            atomicOr(&cmask[r >> 5], 1u << (r & 31));
        }
      }
      __syncthreads();
      if (tid < n_dw)
      {
        p.winfo[2 * ((dest_base >> 5) + tid)] = (woff[tid] << 8) | mw[tid];
        p.winfo[2 * ((dest_base >> 5) + tid) + 1] = cmask[tid];
      }
    }
#pragma unroll
    for (int it = 0; it < ITEMS; ++it)
    {
      const int r = tid * ITEMS + it;
      if (r >= n_dw * 32)
        continue;
      const int w = r >> 5, lane = r & 31, m = (int)mw[w];
      uint16_t* out = p.src + ((src_base32 + woff[w]) << 5) + lane;
      int st = 0, cnt = 0;
      int64_t addr = 0, other = 0;
      if (r < n_dest)
      {
        const int d = v2[it];
        st = dstart[d];
        cnt = min((int)dstart[d + 1] - st, MAX_LIST);
        int64_t blk;
        addr = contrib_addr(p, q, sidx[st], blk, other);
      }
      const int64_t at = (dest_base + r) * (p.sym ? 2 : 1);
      if (p.two)
        ; // the addresses go to the sorted write-back lists below
      else if (p.addr_bytes == 4)
      {
        static_cast<uint32_t*>(p.dest_addr)[at] = (uint32_t)addr;
        if (p.sym)
          static_cast<uint32_t*>(p.dest_addr)[at + 1] = (uint32_t)other;
      }
      else
      {
        static_cast<uint64_t*>(p.dest_addr)[at] = (uint64_t)addr;
        if (p.sym)
          static_cast<uint64_t*>(p.dest_addr)[at + 1] = (uint64_t)other;
      }
      for (int j = 0; j < m; ++j)
        out[j << 5] = j < cnt ? (uint16_t)(sidx[st + j] + sidx[st + j] / p.cb) : (uint16_t)(p.ns * (p.cb + 1));
    }
    if (p.two)
    {
      // ---- 5. write-back lists: (address << 16 | incomplete << 15 | rank) of both entries of every destination,
      //         sorted by address
      uint64_t* wk = reinterpret_cast<uint64_t*>(((uintptr_t)(dstart + NK + 2) + 15) & ~(uintptr_t)15); // see plan_smem
      int* wcnt = reinterpret_cast<int*>(wk + NK);
      if (tid == 0)
        *wcnt = 0;
      __syncthreads();
#pragma unroll
      for (int it = 0; it < ITEMS; ++it)
      {
        const int r = tid * ITEMS + it;
        if (r < n_dest)
        {
          const int d = v2[it];
          int64_t blk, other;
          const int64_t addr = contrib_addr(p, q, sidx[dstart[d]], blk, other);
          const uint64_t inc = p.total[blk] == min((int)dstart[d + 1] - (int)dstart[d], MAX_LIST) ? 0u : 1u;
          // split lists (two == 2): all plain stores first, then all REDs, each in address order
          const uint64_t tag = (inc << 15) | (uint64_t)r | (p.two == 2 ? inc << (16 + p.addr_bits) : 0ull);
          const int at = atomicAdd(wcnt, other != addr ? 2 : 1);
          wk[at] = ((uint64_t)addr << 16) | tag;
          if (other != addr)
            wk[at + 1] = ((uint64_t)other << 16) | tag;
        }
      }
      __syncthreads();
      const int n_wr = *wcnt;
#pragma unroll
      for (int it = 0; it < ITEMS; ++it)
      {
        const int k = tid * ITEMS + it;
        keys[it] = k < n_wr ? wk[k] : ~0ull;
      }
      __syncthreads();
      Sort1(temp.s1).Sort(keys, 16, 16 + p.addr_bits + (p.two == 2 ? 1 : 0));
      const uint64_t amask = (1ull << p.addr_bits) - 1ull;
      const int n_wr_pad = (n_wr + 31) & ~31;
      const int64_t wr_base = 2 * dest_base;
#pragma unroll
      for (int it = 0; it < ITEMS; ++it)
      {
        const int k = tid * ITEMS + it;
        if (k < n_wr_pad)
        {
          p.wr_addr[wr_base + k] = k < n_wr ? (uint32_t)((keys[it] >> 16) & amask) : 0u;
          p.wr_src[wr_base + k] = k < n_wr ? (uint16_t)(keys[it] & 0xffffu) : (uint16_t)0;
        }
      }
      if (tid == 0)
        p.hdr[q].pad = n_wr;
    }
  }
}

template <int ITEMS>
size_t plan_smem()
{
  using Sort1 = cub::BlockRadixSort<uint64_t, PLAN_THREADS, ITEMS>;
  using Sort2 = cub::BlockRadixSort<uint32_t, PLAN_THREADS, ITEMS, uint16_t>;
  using Scan = cub::BlockScan<int, PLAN_THREADS>;
  size_t t = sizeof(typename Sort1::TempStorage);
  t = t > sizeof(typename Sort2::TempStorage) ? t : sizeof(typename Sort2::TempStorage);
  t = t > sizeof(typename Scan::TempStorage) ? t : sizeof(typename Scan::TempStorage);
  t = (t + 15) / 16 * 16;
  // + the staging array of the two-stage write-back lists (NK keys of 8 bytes, a counter, alignment slack)
  return t + PLAN_THREADS * 8 + PLAN_THREADS * 4 * 3 + (size_t)(2 * PLAN_THREADS * ITEMS + 2) * 2 + 16
         + (size_t)PLAN_THREADS * ITEMS * 8 + 64;
}

template <int ITEMS>
int run_plan_pass(bool write, const ChunkBuildArgs& p, int64_t nchunks, cudaStream_t st)
{
  const size_t smem = plan_smem<ITEMS>();
  if (write)
  {
    BFX_CUDA(cudaFuncSetAttribute(k_chunk_plan<ITEMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_chunk_plan<ITEMS, true><<<(unsigned)nchunks, PLAN_THREADS, smem, st>>>(p);
  }
  else
  {
    BFX_CUDA(cudaFuncSetAttribute(k_chunk_plan<ITEMS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_chunk_plan<ITEMS, false><<<(unsigned)nchunks, PLAN_THREADS, smem, st>>>(p);
  }
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

int run_plan_pass_items(int items, bool write, const ChunkBuildArgs& p, int64_t nchunks, cudaStream_t st)
{
  switch (items)
  {
  case 1: return run_plan_pass<1>(write, p, nchunks, st);
  case 2: return run_plan_pass<2>(write, p, nchunks, st);
  case 3: return run_plan_pass<3>(write, p, nchunks, st);
  case 4: return run_plan_pass<4>(write, p, nchunks, st);
  case 5: return run_plan_pass<5>(write, p, nchunks, st);
  case 6: return run_plan_pass<6>(write, p, nchunks, st);
  case 7: return run_plan_pass<7>(write, p, nchunks, st);
  case 8: return run_plan_pass<8>(write, p, nchunks, st);
  case 9: return run_plan_pass<9>(write, p, nchunks, st);
  default: return fail(BFX_ERR_UNSUPPORTED, "chunk plan: %d keys per thread not instantiated", items);
  }
}

// ---- Morton ordering of the cell list ------------------------------------------------------------
__device__ __forceinline__ unsigned long long dkey(double v)
{
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long u)
{
  u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
  return __longlong_as_double((long long)u);
}

__device__ __forceinline__ void centroid(const double* __restrict__ x, const int32_t* __restrict__ xd, int nx,
                                         double (&c)[3])
{
  c[0] = c[1] = c[2] = 0.0;
  for (int k = 0; k < nx; ++k)
  {
    const double* pt = x + 3 * (int64_t)xd[k];
    c[0] += pt[0], c[1] += pt[1], c[2] += pt[2];
  }
  const double s = 1.0 / nx;
  c[0] *= s, c[1] *= s, c[2] *= s;
}

// bb[0..2] = min keys, bb[3..5] = max keys
__global__ void k_centroid_bbox(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ x_dofmap,
                                int nx, const double* __restrict__ x, unsigned long long* __restrict__ bb)
{
  unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0, 0, 0};
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
  {
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    double c[3];
    centroid(x, x_dofmap + (int64_t)cell * nx, nx, c);
#pragma unroll
    for (int m = 0; m < 3; ++m)
    {
      const unsigned long long k = dkey(c[m]);
      lo[m] = k < lo[m] ? k : lo[m];
      hi[m] = k > hi[m] ? k : hi[m];
    }
  }
#pragma unroll
  for (int m = 0; m < 3; ++m)
  {
    for (int o = 16; o > 0; o >>= 1)
    {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo[m], o), b = __shfl_xor_sync(0xffffffffu, hi[m], o);
      lo[m] = a < lo[m] ? a : lo[m];
      hi[m] = b > hi[m] ? b : hi[m];
    }
    if ((threadIdx.x & 31) == 0)
    {
      atomicMin(bb + m, lo[m]);
      atomicMax(bb + 3 + m, hi[m]);
    }
  }
}

__device__ __forceinline__ uint64_t spread21(uint64_t v)
{
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

__global__ void k_morton_keys(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ x_dofmap,
                              int nx, const double* __restrict__ x, const unsigned long long* __restrict__ bb,
                              uint64_t* __restrict__ keys, int32_t* __restrict__ ids)
{
  double lo[3], inv[3];
#pragma unroll
  for (int m = 0; m < 3; ++m)
  {
    lo[m] = dkey_inv(bb[m]);
    const double ext = dkey_inv(bb[3 + m]) - lo[m];
    inv[m] = ext > 0.0 ? 2097151.0 / ext : 0.0;
  }
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
  {
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    double c[3];
    centroid(x, x_dofmap + (int64_t)cell * nx, nx, c);
    uint64_t key = 0;
#pragma unroll
    for (int m = 0; m < 3; ++m)
    {
      double t = (c[m] - lo[m]) * inv[m];
      t = t < 0.0 ? 0.0 : (t > 2097151.0 ? 2097151.0 : t);
      key |= spread21((uint64_t)t) << m;
    }
    keys[e] = key;
    ids[e] = (int32_t)e;
  }
}

__global__ void k_max_dof(int64_t n, const int32_t* __restrict__ cells, const int32_t* __restrict__ dofmap, int nd,
                          int32_t* __restrict__ out)
{
  int32_t m = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
  {
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    for (int k = 0; k < nd; ++k)
      m = max(m, dofmap[(int64_t)cell * nd + k]);
  }
  for (int o = 16; o > 0; o >>= 1)
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0)
    atomicMax(out, m);
}

// Bank colouring of the staging area (one warp per chunk).  Every staged entry is written once (by the half
// warp of 16 consecutive cells c, entry k: "write group") and read once (by a half warp at one step of the
// list walk: "read group").  Giving the 16 entries of a write group the 16 slots of one aligned 128-byte
// line in some order keeps the stores conflict free; the order is chosen so that the entries of a read group
// land in different 8-byte banks too (a proper edge colouring of the bipartite write-group/read-group graph
// exists by Koenig's theorem; this greedy pass finds most of it, what it cannot avoid is counted).
template <int CB>
__global__ void __launch_bounds__(256)
    k_chunk_colour(int64_t nchunks, int ns, const ChunkHdr* __restrict__ hdr, uint16_t* __restrict__ src,
                   uint8_t* __restrict__ colour, int colw, unsigned long long* __restrict__ n_conflicts,
                   int* __restrict__ overflow)
{
  constexpr int WPB = 8;
  extern __shared__ __align__(16) unsigned char craw[];
  const int nst = ns * CB; // staged entries per chunk
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // per warp: pos_of[nst] (uint16: position of the reader in the list, 0xffff = none), used[] masks, col[nst]
  const int max_rg = 2 * src_group_cap(nst);
  const size_t per_warp = ((size_t)nst * 2 + (size_t)max_rg * 2 + (size_t)nst + 15) / 16 * 16;
  uint16_t* pos_of = reinterpret_cast<uint16_t*>(craw + wib * per_warp);
  uint16_t* used = pos_of + nst;
  uint8_t* col = reinterpret_cast<uint8_t*>(used + max_rg);
  const int zero_slot = ns * (CB + 1);
  for (int64_t q = (int64_t)blockIdx.x * WPB + wib; q < nchunks; q += (int64_t)gridDim.x * WPB)
  {
    const ChunkHdr h = hdr[q];
    uint16_t* list = src + (h.src_base32 << 5);
    const int nent = h.n_src32 * 32;
    if (2 * h.n_src32 > max_rg)
    {
      if (lane == 0)
        *overflow = 1; // lists far longer than the entries: the host rebuilds the plan with the linear layout
      continue;
    }
    for (int t = lane; t < nst; t += 32)
      pos_of[t] = 0xffffu;
    for (int t = lane; t < max_rg; t += 32)
      used[t] = 0;
    __syncwarp();
    for (int p = lane; p < nent; p += 32)
    {
      const int a = list[p];
      if (a != zero_slot)
        pos_of[(a / (CB + 1)) * CB + a % (CB + 1)] = (uint16_t)p;
    }
    __syncwarp();
    unsigned long long bad = 0;
    for (int wg = 0; wg < nst / 16; ++wg)
    {
      const int k = wg / (CB / 16), c = (wg % (CB / 16)) * 16 + (lane & 15);
      const int idx = k * CB + c;
      const int pp = lane < 16 ? pos_of[idx] : 0xffff;
      const int rg = pp == 0xffff ? -1 : (pp >> 4);
      const unsigned forb = rg >= 0 ? used[rg] : 0u;
      // sequential greedy over the 16 entries, most constrained (most forbidden banks) first
      unsigned taken = 0, mine = 0;
      unsigned done = 0; // lanes already served (warp-uniform)
      for (int step = 0; step < 16; ++step)
      {
        // pick the unserved lane with the largest number of forbidden colours
        const int score = (lane < 16 && !((done >> lane) & 1u)) ? (__popc((forb | taken) & 0xffffu) << 5) + (31 - lane) : -1;
        int best = score;
        for (int o = 16; o > 0; o >>= 1)
          best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
        const int who = 31 - (best & 31);
        unsigned pick = 0;
        if (lane == who)
        {
          unsigned freec = ~(forb | taken) & 0xffffu;
          if (!freec)
          {
            freec = ~taken & 0xffffu; // unavoidable: two entries of one read group share a bank
            if (rg >= 0)
              ++bad;
          }
          pick = freec & (0u - freec);
          mine = pick;
        }
        pick = __shfl_sync(0xffffffffu, pick, who);
        taken |= pick;
        done |= 1u << who;
        // entries of the same read group inside this write group must see the new bank as forbidden at once
        const int rg_who = __shfl_sync(0xffffffffu, rg, who);
        if (lane == who && rg >= 0)
          used[rg] |= (uint16_t)pick;
        __syncwarp();
        (void)rg_who;
      }
      if (lane < 16)
        col[idx] = (uint8_t)(__ffs(mine) - 1);
      __syncwarp();
      // refresh forb for lanes whose read group was updated by an earlier lane of this write group is not needed:
      // `taken` already excludes every colour used inside the write group
    }
    __syncwarp();
    // rewrite the lists and emit the per-cell colours
    for (int p = lane; p < nent; p += 32)
    {
      const int a = list[p];
      if (a != zero_slot)
      {
        const int k = a / (CB + 1), c = a % (CB + 1);
        list[p] = (uint16_t)((k * (CB / 16) + (c >> 4)) * 16 + col[k * CB + c]);
      }
      else
        list[p] = (uint16_t)nst;
    }
    for (int t = lane; t < CB * colw; t += 32)
    {
      const int c = t / colw, k = t % colw;
      colour[(q * CB + c) * colw + k] = k < ns ? col[k * CB + c] : 0;
    }
    if (lane == 0 && bad)
      atomicAdd(n_conflicts, bad);
    __syncwarp();
  }
}

// Bank-aware ORDER of the source lists (padded linear staging layout, no colours): the 16 lanes of a half warp read
// one list entry each per step; the order of a destination's entries is free (it only fixes the summation order), so
// a greedy pass lets every lane take, at every step, a remaining entry whose 8-byte bank is still unused in that step
// (padding entries of shorter lists float to the steps where nothing fits).  One warp per chunk; lists longer than
// 32 entries keep their order.  What cannot be avoided is counted.
__global__ void __launch_bounds__(256)
    k_chunk_bank_order(int64_t nchunks, const ChunkHdr* __restrict__ hdr, const uint32_t* __restrict__ winfo,
                       uint16_t* __restrict__ src, int zero_slot, int wstride, unsigned long long* __restrict__ n_conflicts)
{
  __shared__ int s_freq[8][2][16]; // per warp and half warp: remaining entries per bank
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int half = lane >> 4, half_base = lane & 16, turn_of_me = lane & 15;
  int* freq = s_freq[wib][half];
  unsigned long long bad = 0;
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nchunks; q += ((int64_t)gridDim.x * blockDim.x) >> 5)
  {
    const ChunkHdr h = hdr[q];
    const int n_dw = (h.n_dest + 31) >> 5;
    const uint32_t* wi = winfo + (h.dest_base >> 5) * wstride;
    for (int dw = 0; dw < n_dw; ++dw)
    {
      const uint32_t info = wi[dw * wstride];
      const int m = (int)(info & 0xffu);
      if (m < 2 || m > 32)
        continue;
      uint16_t* list = src + ((h.src_base32 + (info >> 8)) << 5) + lane;
      uint16_t ent[32], outl[32];
      uint32_t rem = 0; // real entries not yet scheduled
      freq[turn_of_me] = 0;
      __syncwarp();
      for (int j = 0; j < m; ++j)
      {
        ent[j] = list[j << 5];
        if (ent[j] != zero_slot)
        {
          rem |= 1u << j;
          atomicAdd(&freq[ent[j] & 15], 1);
        }
      }
      __syncwarp();
      int pads = m - __popc(rem);
      for (int j = 0; j < m; ++j)
      {
        uint32_t used = 0;
        for (int t0 = 0; t0 < 16; ++t0)
        {
          const int t = (t0 + j) & 15; // rotate the picking order so that no lane is always last
          uint32_t mine = used;
          if (turn_of_me == t)
          {
            // among the remaining entries on a free bank, the one whose bank is the most crowded in this half warp
            int pick = -1, best = -1;
            uint32_t r = rem;
            while (r)
            {
              const int k = __ffs(r) - 1;
              r &= r - 1;
              const int b = ent[k] & 15;
              if (!((used >> b) & 1u) && freq[b] > best)
              {
                best = freq[b];
                pick = k;
              }
            }
            if (pick < 0 && pads > 0)
            {
              --pads;
              outl[j] = (uint16_t)zero_slot;
            }
            else
            {
              if (pick < 0)
              {
                pick = __ffs(rem) - 1; // (rem != 0: real entries + pads == remaining steps)
                ++bad;
              }
              rem &= ~(1u << pick);
              outl[j] = ent[pick];
              mine = used | (1u << (ent[pick] & 15));
              --freq[ent[pick] & 15];
            }
          }
          __syncwarp();
          used = __shfl_sync(0xffffffffu, mine, half_base | t);
        }
      }
      for (int j = 0; j < m; ++j)
        list[j << 5] = outl[j];
      __syncwarp();
    }
  }
  bad = warp_sum(bad);
  if (lane == 0 && bad)
    atomicAdd(n_conflicts, bad);
}

// distinct ids of every group of 32 consecutive slots + per-slot positions in that list (see bfx_chunks)
template <int WMAX>
__global__ void __launch_bounds__(128)
    k_warp_tables(int64_t nslots_pad, int64_t n, int width, const int32_t* __restrict__ rows /* slot-ordered */,
                  int32_t* __restrict__ ids_out, uint8_t* __restrict__ cnt_out, uint8_t* __restrict__ loc_out)
{
  __shared__ int32_t s_id[4][32 * WMAX];
  __shared__ int32_t s_tab[4][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int locw = 4 * ((width + 3) / 4);
  for (int64_t gw = (int64_t)blockIdx.x * 4 + wib; gw * 32 < nslots_pad; gw += (int64_t)gridDim.x * 4)
  {
    const int64_t slot = gw * 32 + lane;
    int32_t* wid = s_id[wib];
    for (int k = 0; k < width; ++k)
      wid[lane * width + k] = slot < n ? rows[slot * width + k] : -1;
    __syncwarp();
    // first occurrences, numbered in index order
    int nfirst = 0;
    unsigned firstmask = 0;
    for (int k = 0; k < width; ++k)
    {
      const int idx = lane * width + k;
      const int32_t id = wid[idx];
      bool first = id >= 0;
      for (int m = 0; m < idx && first; ++m)
        first = wid[m] != id;
      if (first)
      {
        ++nfirst;
        firstmask |= 1u << k;
      }
    }
    int before = nfirst;
    for (int o = 1; o < 32; o <<= 1)
    {
      const int v = __shfl_up_sync(0xffffffffu, before, o);
      if (lane >= o)
        before += v;
    }
    const int total = __shfl_sync(0xffffffffu, before, 31);
    before -= nfirst;
    const bool ok = total >= 1 && total <= 32;
    if (ok)
    {
      int num = before;
      for (int k = 0; k < width; ++k)
        if ((firstmask >> k) & 1u)
          s_tab[wib][num++] = wid[lane * width + k];
    }
    __syncwarp();
    if (lane == 0)
      cnt_out[gw] = ok ? (uint8_t)total : 0;
    ids_out[gw * 32 + lane] = ok ? s_tab[wib][lane < total ? lane : 0] : 0;
    for (int k = 0; k < locw; ++k)
    {
      uint8_t l = 0;
      if (ok && k < width)
      {
        const int32_t id = wid[lane * width + k];
        for (int j = 0; j < total; ++j)
          if (s_tab[wib][j] == id)
            l = (uint8_t)j;
      }
      loc_out[slot * locw + k] = l;
    }
    __syncwarp();
  }
}

__global__ void k_iota64(int64_t n, int64_t* __restrict__ out)
{
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    out[t] = t;
}

__global__ void k_count_zero_bytes(int64_t n, const uint8_t* __restrict__ v, unsigned long long* __restrict__ out)
{
  unsigned long long z = 0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    z += v[t] == 0;
  z = warp_sum(z);
  if ((threadIdx.x & 31) == 0 && z)
    atomicAdd(out, z);
}

__global__ void k_permute_rows(int64_t n, const int32_t* __restrict__ perm, const int32_t* __restrict__ cells,
                               const int32_t* __restrict__ map, int width, int32_t* __restrict__ out)
{
  const int64_t total = n * width;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t slot = t / width;
    const int k = (int)(t - slot * width);
    const int64_t e = perm ? perm[slot] : slot;
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    out[t] = map[(int64_t)cell * width + k];
  }
}

// ---------------------------------------------------------------------------------------------
// assembly kernel
// ---------------------------------------------------------------------------------------------
struct ChunkArgs
{
  const ChunkHdr* hdr;
  const uint32_t* winfo;
  const void* dest_addr;
  const uint16_t* src;
  const int32_t* perm;
  const int32_t *xdm, *dm0, *dm1; // chunk-ordered copies or NULL
  const int32_t *wv_ids, *wd_ids; // warp tables (see bfx_chunks) or NULL
  const uint8_t *wv_cnt, *wd_cnt, *wv_loc, *wd_loc;
  const uint8_t* colour;          // bank colours of the staged entries (per cell slot) or NULL
  const uint32_t *bits0, *bits1;  // bit-packed Dirichlet markers or NULL
  int same_bc;                    // rows and columns share dofmap and markers
  int overwrite;
  const uint32_t* wr_addr;        // two-stage write-back lists (see ChunkBuildArgs) or NULL
  const uint16_t* wr_src;
  int tables_complete;            // every group of 32 cells has its node (and dof) table: no per-group count loads
  // launch over part of the chunks (bfx_assemble_matrix_cells_part on a plan that is not a lean one): chunk_list names
  // the chunk of every CTA (grid = list length), or CTAs of chunks with skip[chunk] != 0 return at once
  const uint32_t* chunk_list;
  const uint8_t* skip;
};

// ---- TMA bulk copy + mbarrier (one chunk's lists are contiguous: two bulk copies per CTA) -----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// L2 prefetch of a contiguous global range (16-byte granules): the data is read later by ordinary loads
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_%=:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra DONE_%=;\n"
               "bra WAIT_%=;\n"
               "DONE_%=:\n"
               "}" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}

// shared-memory budget of the lists of one chunk (chunks over budget read their lists from global)
// destinations of one chunk the two-stage kernel holds in shared memory (typical P1 chunks: 0.22 NS CB)
constexpr int two_stage_dcap(int ns, int cb) { return ns > 36 ? (ns * cb / 2) / 32 * 32 : (ns * cb * 5 / 16) / 32 * 32; }
// two-stage kernels of large elements (P2: 55 staged entries per cell) read the address-ordered write-back lists from
// global memory in phase 3 (coalesced, each entry once) instead of staging them: shared memory keeps two CTAs per SM
constexpr bool two_stage_lists_global(int ns) { return ns > 36; }
// the same for the lean kernel's two-stage variant (tighter: three 384-cell CTAs per SM)
constexpr int lean_two_dcap(int ns, int cb) { return (ns * cb * 3 / 10) / 32 * 32; }

// shared memory of the lean kernel with address-ordered write-back
template <int NS, int CB>
struct LeanTwoSmem
{
  static constexpr int SRC_GROUPS = (NS * CB / 32) * 3 / 2;
  static constexpr int DCAP = lean_two_dcap(NS, CB);
  static constexpr int WINFO = DCAP / 32 * 2;
  static constexpr size_t E_BYTES = (sizeof(double) * ((size_t)NS * (CB + 1) + 2) + 127) / 128 * 128;
  static constexpr size_t SRC_OFF = E_BYTES;
  static constexpr size_t WADDR_OFF = SRC_OFF + (size_t)SRC_GROUPS * 64;
  static constexpr size_t WSRC_OFF = WADDR_OFF + (size_t)DCAP * 2 * 4;
  static constexpr size_t WINFO_OFF = WSRC_OFF + (size_t)DCAP * 2 * 2;
  static constexpr size_t SUM_OFF = WINFO_OFF + (((size_t)WINFO * 4 + 15) / 16) * 16;
  static constexpr size_t BAR_OFF = SUM_OFF + (size_t)DCAP * 8;
  static constexpr size_t TOTAL = BAR_OFF + 16;
};

template <int NS, int CB, int DSTRIDE, bool TWO = false>
struct ChunkSmem
{
  static constexpr int SRC_GROUPS = src_group_cap(NS * CB);        // 32-entry groups of source entries
  static constexpr int DCAP = TWO ? two_stage_dcap(NS, CB) : NS * CB / 2;
  static constexpr bool WRG = TWO && two_stage_lists_global(NS);
  // (round 2, measured and dropped: destination addresses of the large elements read from global memory by the list
  // walk, L2-prefetched at CTA start - 28 KB less shared memory per P2 chunk.  128 cells: still two CTAs per SM, 4.24
  // against 3.87 ms; 96 cells with three CTAs per SM: 3.84 ms, no better than the staged lists - profiles/r02_p2_variants.txt)
  static constexpr int DEST_BYTES = WRG ? 0 : DCAP * 4 * DSTRIDE; // destination addresses / write-back addresses
  static constexpr int WINFO = NS * CB / 32;                      // groups of 32 destinations
  static constexpr size_t E_BYTES = (sizeof(double) * ((size_t)NS * (CB + 1) + 2) + 127) / 128 * 128;
  static constexpr size_t SRC_OFF = E_BYTES;
  static constexpr size_t DEST_OFF = SRC_OFF + (size_t)SRC_GROUPS * 64;
  static constexpr size_t WINFO_OFF = DEST_OFF + DEST_BYTES;
  static constexpr size_t BAR_OFF = WINFO_OFF + (((size_t)WINFO * 4 + 15) / 16) * 16 + 16;
  static constexpr size_t WRSRC_OFF = BAR_OFF + 16;                 // two-stage: ranks of the write-back lists
  static constexpr size_t SUM_OFF = WRSRC_OFF + ((TWO && !WRG) ? (size_t)DCAP * 2 * 2 : 0); // two-stage: one sum per destination
  static constexpr size_t TOTAL = SUM_OFF + (TWO ? (size_t)DCAP * 8 : 0);
};

// Phase 2 of the DIET kernel variant: the same walk as the classic kernel (same summation order, same updates), as a
// function that the kernel inlines TWICE - once on the shared-memory copies of the chunk's lists, once on the global
// arrays - so that each copy addresses ONE state space (LDS / 32-bit addresses in the first) instead of generic loads
// with 64-bit arithmetic, and with the 4-step loop kept rolled: the classic code is unrolled to 16 steps with a cascade
// of 8 / 4 / 3 / 2 / 1-step remainders, 132 instructions per destination for lists of ~6 entries.
// walk of the two-stage variant: the sum of destination t is left in sums[t] (no addresses needed)
template <int THREADS, int WS>
__device__ __forceinline__ void chunk_walk_sums(const uint16_t* __restrict__ srcp, const uint32_t* __restrict__ winfop,
                                                const double* __restrict__ Es, double* __restrict__ sums, int n_dw)
{
  const int lane = threadIdx.x & 31;
  for (int dw = threadIdx.x >> 5; dw < n_dw; dw += THREADS / 32)
  {
    const uint32_t info = winfop[WS * dw];
    const int m = (int)(info & 0xffu);
    const uint16_t* p = srcp + ((info >> 8) << 5) + lane;
    double s0 = 0.0, s1 = 0.0;
    int j = 0;
#pragma unroll 1
    for (; j + 4 <= m; j += 4)
    {
      const uint32_t i0 = p[(j + 0) << 5], i1 = p[(j + 1) << 5], i2 = p[(j + 2) << 5], i3 = p[(j + 3) << 5];
      s0 += Es[i0];
      s1 += Es[i1];
      s0 += Es[i2];
      s1 += Es[i3];
    }
#pragma unroll 1
    for (; j < m; ++j)
      s0 += Es[p[j << 5]];
    sums[(dw << 5) + lane] = s0 + s1;
  }
}

template <bool SYM, int THREADS, typename AddrT, int DBG = 0, bool MASKED = false>
__device__ __forceinline__ void chunk_walk(const uint16_t* __restrict__ srcp, const AddrT* __restrict__ destp,
                                           const uint32_t* __restrict__ winfop, const double* __restrict__ Es,
                                           double* __restrict__ values, int n_dest, int n_complete, int n_dw, int overwrite)
{
  constexpr int DS = SYM ? 2 : 1;
  const int lane = threadIdx.x & 31;
  for (int dw = threadIdx.x >> 5; dw < n_dw; dw += THREADS / 32)
  {
    const uint32_t info = MASKED ? winfop[2 * dw] : winfop[dw];
    const int m = (int)(info & 0xffu);
    const uint16_t* p = srcp + ((info >> 8) << 5) + lane;
    const int t = (dw << 5) + lane;
    double* dst = values + (int64_t)destp[t * DS];
    double* dst2 = SYM ? values + (int64_t)destp[t * DS + (DS - 1)] : dst;
    const bool plain = MASKED ? ((winfop[2 * dw + 1] >> lane) & 1u) != 0u : t < n_complete;
    double old = 0.0, old2 = 0.0;
    if (plain && !overwrite)
    {
      old = *dst;
      if (SYM)
        old2 = *dst2;
    }
    double s0 = 0.0, s1 = 0.0;
    int j = 0;
#pragma unroll 1
    for (; j + 4 <= m; j += 4)
    {
      const uint32_t i0 = p[(j + 0) << 5], i1 = p[(j + 1) << 5], i2 = p[(j + 2) << 5], i3 = p[(j + 3) << 5];
      s0 += Es[i0];
      s1 += Es[i1];
      s0 += Es[i2];
      s1 += Es[i3];
    }
#pragma unroll 1
    for (; j < m; ++j) // (empty when the plan pads the lists to multiples of 4: BFX_CHUNKS_PAD4)
      s0 += Es[p[j << 5]];
    const double sum = s0 + s1;
    if (t < n_dest && (DBG != 3 || sum == 1.2345e300))
    {
      if (plain)
      {
        *dst = old + sum;
        if (SYM && dst2 != dst)
          *dst2 = old2 + sum;
      }
      else
      {
        red_add(dst, sum);
        if (SYM && dst2 != dst)
          red_add(dst2, sum);
      }
    }
  }
}

// TWO (two-stage write-back, symmetric plans with 32-bit addresses): phase 2 leaves the sum of every destination in
// shared memory; a third phase walks the chunk's (address, destination) list in ADDRESS order, so that consecutive
// lanes update consecutive CSR values - both entries of a symmetric pair included - instead of two scattered ones each
// OCC: resident CTAs per SM asked of the register allocator (0 = chunk_min_ctas(THREADS))
// DIET: phase 2 through chunk_walk (round-2 experiment, selected by bfx_asm_chunk_set_kernel; not the default)
// DBG (profiling only, results are wrong): 1 = phase 1 only, 2 = phase 2 only, 3 = phase 2 without the global updates
template <class E, bool SYM, int CB, int THREADS, typename AddrT, bool TWO = false, int OCC = 0, bool DIET = false, int DBG = 0>
__global__ void __launch_bounds__(THREADS, OCC ? OCC : (staged_per_cell(E::ND * E::BS, SYM) > 36 ? 2 : chunk_min_ctas(THREADS)))
    k_matrix_chunked(const AsmArgs a, const ChunkArgs ch)
{
  constexpr int NX = E::NX, ND = E::ND, BS = E::BS, N = ND * BS, NS = staged_per_cell(N, SYM), TPC = THREADS / CB;
  constexpr int DS = SYM ? 2 : 1; // addresses per destination
  static_assert(THREADS % CB == 0 && CB % 32 == 0, "a warp must work on one row residue");
  static_assert(!SYM || BS == 1, "symmetric staging is implemented for block size 1");
  static_assert(!TWO || (SYM && sizeof(AddrT) == 4), "two-stage write-back: symmetric plans, 32-bit addresses");
  using L = ChunkSmem<NS, CB, DS, TWO>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Es = reinterpret_cast<double*>(smem_raw);
  uint16_t* s_src = reinterpret_cast<uint16_t*>(smem_raw + L::SRC_OFF);
  AddrT* s_dest = reinterpret_cast<AddrT*>(smem_raw + L::DEST_OFF);
  uint32_t* s_winfo = reinterpret_cast<uint32_t*>(smem_raw + L::WINFO_OFF);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L::BAR_OFF);

  const int c = threadIdx.x % CB, sub = threadIdx.x / CB;
  int64_t q = blockIdx.x;
  if (ch.chunk_list)
    q = ch.chunk_list[blockIdx.x];
  else if (ch.skip && ch.skip[q])
    return;
  const int64_t slot = q * CB + c;

  // ---- first-level loads of phase 1, all issued before anything waits (header, group tables, positions):
  //      the dependent second level (marker words, coordinates) then costs one more latency, not three
  constexpr int LOCWV = (NX + 3) / 4, LOCWD = (ND + 3) / 4;
  const int lane1 = threadIdx.x & 31;
  const int64_t gw = (q * CB + (c & ~31)) >> 5; // group of 32 consecutive cell slots this warp works on
  const bool use_bits = ch.bits0 || ch.bits1;
  const bool dof_tab = BS == 1 && use_bits && ch.wd_cnt && (ch.same_bc || !ch.bits1);
  const int dcnt = (dof_tab && (DBG < 2 || DBG == 4)) ? ch.wd_cnt[gw] : 0;
  const int vcnt = (ch.wv_cnt && (DBG < 2 || DBG == 4)) ? ch.wv_cnt[gw] : 0;
  const int64_t dof_id = dof_tab ? ch.wd_ids[gw * 32 + lane1] : 0;
  const int32_t vtx_id = ch.wv_cnt ? ch.wv_ids[gw * 32 + lane1] : 0;
  constexpr int COLW = (NS + 3) / 4;
  uint32_t colw[NS <= COLOUR_MAX_NS ? COLW : 1];
  if constexpr (NS <= COLOUR_MAX_NS)
  {
#pragma unroll
    for (int k = 0; k < COLW; ++k)
      colw[k] = ch.colour ? __ldg(reinterpret_cast<const uint32_t*>(ch.colour) + slot * COLW + k) : 0u;
  }
  uint32_t locd[LOCWD], locv[LOCWV];
#pragma unroll
  for (int k = 0; k < LOCWD; ++k)
    locd[k] = dof_tab ? __ldg(reinterpret_cast<const uint32_t*>(ch.wd_loc) + slot * LOCWD + k) : 0u;
#pragma unroll
  for (int k = 0; k < LOCWV; ++k)
    locv[k] = ch.wv_cnt ? __ldg(reinterpret_cast<const uint32_t*>(ch.wv_loc) + slot * LOCWV + k) : 0u;

  // no dof table (more than 32 distinct dofs per 32 cells: P2): the cell's own dofmap row is a first-level load too, so
  // that the marker words travel together with the coordinates instead of after the geometry
  constexpr bool EARLY_D0 = BS == 1 && ND > 4;
  const bool early_bits = EARLY_D0 && use_bits && !dof_tab && ch.bits0 && (ch.same_bc || !ch.bits1) && slot < a.n
                          && (DBG < 2 || DBG == 4) && ch.dm0;
  constexpr int NDE = EARLY_D0 ? ND : 1;
  int32_t d0e[NDE];
  if (early_bits)
    load_ints<NDE>(ch.dm0 + slot * ND, d0e);

  // ---- prefetch of the chunk's lists: TMA bulk copies land while phase 1 computes
  const ChunkHdr h = ch.hdr[q];
  const int n_dw = (h.n_dest + 31) >> 5;
  const uint32_t n_wr_pad = TWO ? ((uint32_t)h.pad + 31u) & ~31u : 0u; // two-stage: (address, rank) pairs of the chunk
  constexpr bool WRG = L::WRG; // two-stage write-back lists read from global memory in phase 3
  const uint32_t src_bytes = (uint32_t)h.n_src32 * 64u,
                 dest_bytes = WRG ? 0u : (TWO ? n_wr_pad * 4u : (uint32_t)n_dw * 32u * DS * (uint32_t)sizeof(AddrT));
  // (a two-stage plan is only built when the destinations of every chunk fit: bfx_asm_build_chunks)
  const bool fits = h.n_src32 <= L::SRC_GROUPS && dest_bytes <= (uint32_t)L::DEST_BYTES && n_dw <= L::WINFO
                    && (!WRG || n_dw * 32 <= L::DCAP);
  const uint16_t* g_src = ch.src + (h.src_base32 << 5);
  const AddrT* g_dest = TWO ? reinterpret_cast<const AddrT*>(ch.wr_addr) + h.dest_base * 2
                            : static_cast<const AddrT*>(ch.dest_addr) + h.dest_base * DS;
  const uint32_t* g_winfo = ch.winfo + (h.dest_base >> 5);
  if (threadIdx.x == 0)
  {
    // one thread arms the barrier and starts the copies; nobody else touches the header before phase 2
    mbar_init(bar, 1);
    Es[ch.colour ? NS * CB : NS * (CB + 1)] = 0.0; // the slot padded list entries point at
    if (fits && n_dw > 0 && DBG != 4)
    {
      const uint32_t winfo_bytes = ((uint32_t)n_dw * 4u + 15u) & ~15u; // chunks start on 4-group boundaries
      mbar_expect_tx(bar, src_bytes + dest_bytes + winfo_bytes + ((TWO && !WRG) ? n_wr_pad * 2u : 0u));
      bulk_g2s(s_src, g_src, src_bytes, bar);
      if (dest_bytes > 0)
        bulk_g2s(s_dest, g_dest, dest_bytes, bar);
      bulk_g2s(s_winfo, g_winfo, winfo_bytes, bar);
      if (TWO && !WRG && n_wr_pad > 0)
        bulk_g2s(smem_raw + L::WRSRC_OFF, ch.wr_src + h.dest_base * 2, n_wr_pad * 2u, bar);
      if (WRG && n_wr_pad > 0) // phase 3 reads these with ordinary loads: have them in L2 by then
      {
        bulk_prefetch_l2(ch.wr_addr + h.dest_base * 2, n_wr_pad * 4u);
        bulk_prefetch_l2(ch.wr_src + h.dest_base * 2, n_wr_pad * 2u);
      }
    }
  }

  // ---- phase 1: element matrices of the chunk -> shared memory (entry-major)
  const bool active = slot < a.n && (DBG < 2 || DBG == 4);
  int64_t e = slot;
  int32_t cell = (int32_t)slot;
  if (active && (E::WSIZE > 0 || !ch.xdm))
  {
    e = ch.perm ? ch.perm[slot] : slot;
    cell = a.cells ? a.cells[e] : (int32_t)e;
  }
  // second level: the marker word of this lane's dof and the coordinates of this lane's node
  uint32_t bword = 0;
  double px = 0.0, py = 0.0, pz = 0.0;
  if (dcnt && ch.bits0)
    bword = __ldg(ch.bits0 + (dof_id >> 5));
  if (vcnt)
  {
    const double* pp = a.x + 3 * (int64_t)vtx_id;
    px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
  }
  uint32_t bwe[NDE];
  if (early_bits)
  {
#pragma unroll
    for (int i = 0; i < NDE; ++i)
      bwe[i] = __ldg(ch.bits0 + (d0e[i] >> 5));
  }
  // -- coordinates
  double xc[NX][3];
  if (vcnt) // warp-uniform
  {
#pragma unroll
    for (int v = 0; v < NX; ++v)
    {
      const int l = (int)((locv[v >> 2] >> (8 * (v & 3))) & 31u);
      xc[v][0] = __shfl_sync(0xffffffffu, px, l);
      xc[v][1] = __shfl_sync(0xffffffffu, py, l);
      xc[v][2] = __shfl_sync(0xffffffffu, pz, l);
    }
  }
  else if (active)
  {
    int32_t xd[NX];
    load_ints<NX>(ch.xdm ? ch.xdm + slot * NX : a.x_dofmap + (int64_t)cell * NX, xd);
    gather_coords<NX>(a.x, xd, xc);
  }
  typename E::Geo g;
  if (active)
  {
    double w[E::WSIZE > 0 ? E::WSIZE : 1];
    load_w<E>(a, e, cell, w);
    E::prepare(g, xc, w, a.constants, 0);
  }
  // -- Dirichlet marker bit per scalar row / column of Ae
  uint32_t m0 = 0, m1 = 0;
  if (use_bits)
  {
    if (dcnt) // warp-uniform: one lane fetched the bit of one distinct dof, the cells pick theirs with shuffles
    {
      const uint32_t b = (bword >> (dof_id & 31)) & 1u;
#pragma unroll
      for (int i = 0; i < ND; ++i)
        m0 |= __shfl_sync(0xffffffffu, b, (locd[i >> 2] >> (8 * (i & 3))) & 31u) << i;
      if (ch.same_bc)
        m1 = m0;
    }
    else if (early_bits)
    {
#pragma unroll
      for (int i = 0; i < NDE; ++i)
        m0 |= ((bwe[i] >> (d0e[i] & 31)) & 1u) << i;
      if (ch.same_bc)
        m1 = m0;
    }
    else if (active)
    {
      int32_t d0[ND];
      load_ints<ND>(ch.dm0 ? ch.dm0 + slot * ND : a.dofmap0 + (int64_t)cell * ND, d0);
      if (ch.bits0)
      {
#pragma unroll
        for (int i = 0; i < N; ++i)
        {
          const int64_t d = (int64_t)BS * d0[i / BS] + i % BS;
          m0 |= ((__ldg(ch.bits0 + (d >> 5)) >> (d & 31)) & 1u) << i;
        }
      }
      if (ch.same_bc)
        m1 = m0;
      else if (ch.bits1)
      {
        int32_t d1[ND];
        load_ints<ND>(ch.dm1 ? ch.dm1 + slot * ND : a.dofmap1 + (int64_t)cell * ND, d1);
#pragma unroll
        for (int j = 0; j < N; ++j)
        {
          const int64_t d = (int64_t)BS * d1[j / BS] + j % BS;
          m1 |= ((__ldg(ch.bits1 + (d >> 5)) >> (d & 31)) & 1u) << j;
        }
      }
    }
  }
  if (active)
  {
#pragma unroll
    for (int i = 0; i < N; ++i)
    {
      if (TPC > 1 && i % TPC != sub) // warp-uniform: the TPC threads of a cell share its rows
        continue;
      double row[N];
      E::row(g, i, row);
      const bool zr = (m0 >> i) & 1u;
#pragma unroll
      for (int j = SYM ? i : 0; j < N; ++j)
      {
        // symmetric form: only j >= i is staged, at rank i N - i (i - 1) / 2 + (j - i)
        const int k = SYM ? i * N - i * (i - 1) / 2 + (j - i) : i * N + j;
        int at = k * (CB + 1) + c;
        if constexpr (NS <= COLOUR_MAX_NS)
        {
          if (ch.colour) // the 16 cells of a half warp permute the 16 slots of one 128-byte line
            at = (k * (CB / 16) + (c >> 4)) * 16 + (int)((colw[k >> 2] >> (8 * (k & 3))) & 15u);
        }
        Es[at] = (zr || ((m1 >> j) & 1u)) ? 0.0 : row[j];
      }
    }
  }
  __syncthreads();
  if (fits && n_dw > 0 && DBG != 4)
    mbar_wait(bar, 0);
  if constexpr (DBG == 1 || DBG == 4)
  {
    double chk = 0.0; // keeps the staged values alive
    for (int k = 0; k < NS; ++k)
      chk += Es[k * CB + ((threadIdx.x * 7 + k) % CB)];
    if (chk == 1.2345e300)
      a.values[0] = chk;
    return;
  }

  if constexpr (DIET && TWO)
  {
    double* sums = reinterpret_cast<double*>(smem_raw + L::SUM_OFF);
    if (fits)
      chunk_walk_sums<THREADS, 1>(s_src, s_winfo, Es, sums, n_dw);
    else
      chunk_walk_sums<THREADS, 1>(g_src, g_winfo, Es, sums, n_dw);
  }
  else if constexpr (DIET)
  {
    if (fits)
      chunk_walk<SYM, THREADS, AddrT, DBG>(s_src, s_dest, s_winfo, Es, a.values, h.n_dest, h.n_complete, n_dw, ch.overwrite);
    else
      chunk_walk<SYM, THREADS, AddrT, DBG>(g_src, g_dest, g_winfo, Es, a.values, h.n_dest, h.n_complete, n_dw, ch.overwrite);
    return;
  }
  // ---- phase 2: one thread per distinct destination
  const int lane = threadIdx.x & 31;
  const uint16_t* srcp = fits ? s_src : g_src;
  const AddrT* destp = (fits && !WRG) ? s_dest : g_dest;
  const uint32_t* winfop = fits ? s_winfo : g_winfo;
  for (int dw = threadIdx.x >> 5; dw < ((DIET && TWO) ? 0 : n_dw); dw += THREADS / 32)
  {
    const uint32_t info = winfop[dw];
    const int m = (int)(info & 0xffu);
    const uint16_t* p = srcp + ((size_t)(info >> 8) << 5) + lane;
    const int t = (dw << 5) + lane;
    double *dst = nullptr, *dst2 = nullptr;
    double old = 0.0, old2 = 0.0;
    const bool plain = t < h.n_complete;
    if constexpr (!TWO)
    {
      dst = a.values + (int64_t)destp[t * DS];
      dst2 = SYM ? a.values + (int64_t)destp[t * DS + (DS - 1)] : dst;
      if (plain && !ch.overwrite)
      {
        old = *dst; // issued before the list walk: its latency hides behind the shared-memory sums
        if (SYM)
          old2 = *dst2;
      }
    }
    double s0 = 0.0, s1 = 0.0;
    int j = 0;
    // the lists hold shared-memory slots: k * (CB + 1) + c (padded linear layout) or the bank-coloured slot
    for (; j + 4 <= m; j += 4)
    {
      const uint32_t i0 = p[(j + 0) << 5], i1 = p[(j + 1) << 5], i2 = p[(j + 2) << 5], i3 = p[(j + 3) << 5];
      s0 += Es[i0];
      s1 += Es[i1];
      s0 += Es[i2];
      s1 += Es[i3];
    }
    for (; j < m; ++j)
    {
      const uint32_t i0 = p[j << 5];
      s0 += Es[i0];
    }
    const double sum = s0 + s1;
    if constexpr (TWO)
    {
      if (t < h.n_dest)
        reinterpret_cast<double*>(smem_raw + L::SUM_OFF)[t] = sum;
    }
    else if (t < h.n_dest)
    {
      if (plain)
      {
        *dst = old + sum;
        if (SYM && dst2 != dst)
          *dst2 = old2 + sum;
      }
      else
      {
        red_add(dst, sum);
        if (SYM && dst2 != dst)
          red_add(dst2, sum);
      }
    }
  }
  if constexpr (TWO)
  {
    // ---- phase 3: write-back in address order
    __syncthreads();
    const double* sums = reinterpret_cast<const double*>(smem_raw + L::SUM_OFF);
    const uint16_t* wsrc = (fits && !WRG) ? reinterpret_cast<const uint16_t*>(smem_raw + L::WRSRC_OFF) : ch.wr_src + h.dest_base * 2;
    const uint32_t* waddr = reinterpret_cast<const uint32_t*>(destp);
#pragma unroll 8
    for (int t = threadIdx.x; t < h.pad; t += THREADS)
    {
      const uint32_t sx = wsrc[t];
      double* dst = a.values + waddr[t];
      const double v = sums[sx & 0x7fffu];
      if (sx & 0x8000u)
        red_add(dst, v);
      else if (ch.overwrite)
        *dst = v;
      else
        *dst += v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LEAN kernel (round 2): the same algorithm as k_matrix_chunked for symmetric plans of P1-sized scalar elements
// (<= 4 nodes, <= 4 dofs per cell) with the padded linear staging layout, 32-bit addresses and complete warp
// tables, rewritten for instruction count - phase 1 of the classic kernel alone runs 1.87 ms at C2 at 80 % issue
// utilisation (profiles/r02_phase_split.txt): 450 instructions per cell of which 104 are fp64.  Here: 32-bit slot
// arithmetic, one table word per cell, unconditional shuffles (no WARPSYNC / collective bookkeeping), Dirichlet
// handling behind one warp vote, staging through immediate offsets.
// ---------------------------------------------------------------------------------------------
// one warp per chunk: does any dof of the chunk's tables lie at or beyond n_owned (a ghost row)?
__global__ void k_chunk_partition(int64_t nchunks, int cb, const int32_t* __restrict__ wd_ids, int64_t ncells,
                                  int32_t n_owned, int32_t* __restrict__ flag)
{
  const int lane = threadIdx.x & 31;
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nchunks; q += ((int64_t)gridDim.x * blockDim.x) >> 5)
  {
    bool hit = false;
    for (int k = lane; k < cb; k += 32)
    {
      // (table words of padding slots repeat a dof of their group, or dof 0 in a group without cells: never a ghost
      // unless everything is)
      hit |= wd_ids[q * cb + k] >= n_owned;
    }
    hit = __any_sync(0xffffffffu, hit);
    if (lane == 0)
      flag[q] = hit ? 1 : 0;
  }
}

// the same from the chunk-ordered dofmap (plans without dof tables: more than 4 dofs per cell): flag + list
__global__ void k_chunk_partition_dofmap(int64_t nchunks, int cb, const int32_t* __restrict__ dm0, int nd0, int64_t ncells,
                                         int32_t n_owned, uint8_t* __restrict__ flag, uint32_t* __restrict__ list,
                                         unsigned long long* __restrict__ count)
{
  const int lane = threadIdx.x & 31;
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nchunks; q += ((int64_t)gridDim.x * blockDim.x) >> 5)
  {
    bool hit = false;
    const int64_t s0 = q * cb, s1 = min(s0 + cb, ncells);
    for (int64_t k = s0 * nd0 + lane; k < s1 * nd0; k += 32)
      hit |= dm0[k] >= n_owned;
    hit = __any_sync(0xffffffffu, hit);
    if (lane == 0)
    {
      flag[q] = hit ? 1 : 0;
      if (hit)
        list[atomicAdd(count, 1ull)] = (uint32_t)q;
    }
  }
}

// new position of every chunk: flagged chunks first, both classes in their old (Morton) order
__global__ void k_chunk_new_index(int64_t nchunks, const int32_t* __restrict__ flag, const int32_t* __restrict__ before,
                                  int32_t n_first, int32_t* __restrict__ to)
{
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nchunks; q += (int64_t)gridDim.x * blockDim.x)
    to[q] = flag[q] ? before[q] : n_first + (int32_t)(q - before[q]);
}

// per-chunk records moved to their new positions: the header and the cb table words of every table
__global__ void k_chunk_move(int64_t nchunks, int cb, const int32_t* __restrict__ to, const ChunkHdr* __restrict__ hdr,
                             ChunkHdr* __restrict__ hdr_new, const uint32_t* __restrict__ t0, uint32_t* __restrict__ n0,
                             const uint32_t* __restrict__ t1, uint32_t* __restrict__ n1, const uint32_t* __restrict__ t2,
                             uint32_t* __restrict__ n2, const uint32_t* __restrict__ t3, uint32_t* __restrict__ n3,
                             int64_t ncells, int already_counted)
{
  const int64_t total = nchunks * cb;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t q = s / cb;
    const int k = (int)(s - q * cb);
    const int64_t d = (int64_t)to[q] * cb + k;
    n0[d] = t0[s], n1[d] = t1[s], n2[d] = t2[s], n3[d] = t3[s];
    if (k == 0)
    {
      ChunkHdr h = hdr[q];
      if (!already_counted) // first partition of this plan: chunk q holds the cells [q cb, min((q + 1) cb, ncells))
        h.pad = (int32_t)max((int64_t)0, min((int64_t)cb, ncells - q * cb));
      hdr_new[to[q]] = h;
    }
  }
}

struct LeanArgs
{
  const ChunkHdr* hdr;
  const uint32_t* winfo;
  const uint32_t* dest; // two addresses per destination
  const uint16_t* src;
  const int32_t *wv_ids, *wd_ids; // one entry per cell slot: the node / dof this LANE fetches for its group of 32 cells
  const uint32_t *wv_loc, *wd_loc; // per cell slot: 4 x uint8 positions of its nodes / dofs in the group's table
  const uint32_t* bits;           // bit-packed Dirichlet markers or NULL
  const double* x;
  double* values;
  double constants[4];
  uint32_t n; // cells of the plan
  int overwrite;
  // launch over part of the chunks: the grid covers the chunks [chunk_begin, chunk_begin + gridDim.x)
  // (bfx_asm_chunk_partition puts the chunks on ghost rows first)
  uint32_t chunk_begin;
  int valid_in_hdr; // cells of chunk q = hdr[q].pad (reordered plans) instead of min(CB, n - q CB)
};

template <class E, int CB, int DBG = 0, int OCC = 0, bool MASKED = false>
__global__ void __launch_bounds__(CB, OCC ? OCC : chunk_min_ctas(CB)) k_matrix_lean(const LeanArgs p)
{
  constexpr int NX = E::NX, ND = E::ND, NS = ND * (ND + 1) / 2, LD = CB + 1;
  static_assert(E::BS == 1 && NX <= 4 && ND <= 4 && E::WSIZE == 0, "lean kernel: P1-sized scalar elements without coefficients");
  using L = ChunkSmem<NS, CB, 2, false>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Es = reinterpret_cast<double*>(smem_raw);
  uint16_t* s_src = reinterpret_cast<uint16_t*>(smem_raw + L::SRC_OFF);
  uint32_t* s_dest = reinterpret_cast<uint32_t*>(smem_raw + L::DEST_OFF);
  uint32_t* s_winfo = reinterpret_cast<uint32_t*>(smem_raw + L::WINFO_OFF);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L::BAR_OFF);

  const uint32_t tid = threadIdx.x;
  const uint32_t q = blockIdx.x + p.chunk_begin;
  const uint32_t slot = q * (uint32_t)CB + tid;
  // ---- first-level loads (coalesced, one word each)
  const int32_t vtx = __ldg(p.wv_ids + slot);
  const uint32_t locv = __ldg(p.wv_loc + slot);
  uint32_t dof = 0, locd = 0;
  if (p.bits)
  {
    dof = (uint32_t)__ldg(p.wd_ids + slot);
    locd = __ldg(p.wd_loc + slot);
  }
  const ChunkHdr h = p.hdr[q];
  const int n_dw = (h.n_dest + 31) >> 5;
  const uint32_t src_bytes = (uint32_t)h.n_src32 * 64u, dest_bytes = (uint32_t)n_dw * 256u;
  constexpr int WS = MASKED ? 2 : 1; // words per group: (list offset, length) [, completeness mask]
  const bool fits = h.n_src32 <= L::SRC_GROUPS && dest_bytes <= (uint32_t)L::DEST_BYTES && WS * n_dw <= L::WINFO;
  const uint16_t* g_src = p.src + (h.src_base32 << 5);
  const uint32_t* g_dest = p.dest + h.dest_base * 2;
  const uint32_t* g_winfo = p.winfo + (h.dest_base >> 5) * WS;
  if (tid == 0)
  {
    mbar_init(bar, 1);
    Es[NS * LD] = 0.0; // the slot padded list entries point at
    if (fits && n_dw > 0)
    {
      const uint32_t winfo_bytes = ((uint32_t)(n_dw * WS) * 4u + 15u) & ~15u;
      mbar_expect_tx(bar, src_bytes + dest_bytes + winfo_bytes);
      bulk_g2s(s_src, g_src, src_bytes, bar);
      bulk_g2s(s_dest, g_dest, dest_bytes, bar);
      bulk_g2s(s_winfo, g_winfo, winfo_bytes, bar);
    }
  }
  // ---- second level: this lane's node (and marker word)
  const double* pp = p.x + 3 * (int64_t)vtx;
  const double px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
  uint32_t bit = 0;
  if (p.bits)
    bit = (__ldg(p.bits + (dof >> 5)) >> (dof & 31u)) & 1u;
  // ---- coordinates of this cell's nodes (the shuffle uses the low 5 bits of the lane operand)
  double xc[NX][3];
#pragma unroll
  for (int v = 0; v < NX; ++v)
  {
    const int l = (int)(locv >> (8 * v));
    xc[v][0] = __shfl_sync(0xffffffffu, px, l);
    xc[v][1] = __shfl_sync(0xffffffffu, py, l);
    xc[v][2] = __shfl_sync(0xffffffffu, pz, l);
  }
  uint32_t m = 0;
  if (__any_sync(0xffffffffu, bit)) // some cell of this warp touches a Dirichlet dof
  {
#pragma unroll
    for (int i = 0; i < ND; ++i)
      m |= __shfl_sync(0xffffffffu, bit, (int)(locd >> (8 * i))) << i;
  }
  // (a partitioned plan has its chunks reordered: the one partly filled chunk is no longer the last, its count of
  // cells travels in the header)
  const bool has_cell = p.valid_in_hdr ? tid < (uint32_t)h.pad : slot < p.n;
  if (has_cell && DBG != 2 && DBG != 3)
  {
    typename E::Geo g;
    E::prepare(g, xc, nullptr, p.constants, 0);
    double* e = Es + tid;
    if (m == 0)
    {
#pragma unroll
      for (int i = 0; i < ND; ++i)
      {
        double row[ND];
        E::row(g, i, row);
#pragma unroll
        for (int j = i; j < ND; ++j)
          e[(i * ND - i * (i - 1) / 2 + (j - i)) * LD] = row[j];
      }
    }
    else
    {
#pragma unroll
      for (int i = 0; i < ND; ++i)
      {
        double row[ND];
        E::row(g, i, row);
#pragma unroll
        for (int j = i; j < ND; ++j)
          e[(i * ND - i * (i - 1) / 2 + (j - i)) * LD] = ((m >> i) | (m >> j)) & 1u ? 0.0 : row[j];
      }
    }
  }
  __syncthreads();
  if (fits && n_dw > 0)
    mbar_wait(bar, 0);
  if constexpr (DBG == 1)
  {
    double chk = 0.0; // keeps the staged values alive
    for (int k = 0; k < NS; ++k)
      chk += Es[k * LD + ((tid * 7 + k) % CB)];
    if (chk == 1.2345e300)
      p.values[0] = chk;
    return;
  }
  if (fits)
    chunk_walk<true, CB, uint32_t, DBG, MASKED>(s_src, s_dest, s_winfo, Es, p.values, h.n_dest, h.n_complete, n_dw, p.overwrite);
  else
    chunk_walk<true, CB, uint32_t, DBG, MASKED>(g_src, g_dest, g_winfo, Es, p.values, h.n_dest, h.n_complete, n_dw, p.overwrite);
}

// Phase 1 of the lean kernels as a function: table words, the lane's node, shuffles, element matrix -> Es
template <class E, int CB>
__device__ __forceinline__ void lean_stage(const LeanArgs& p, double* __restrict__ Es, uint32_t tid, uint32_t slot)
{
  constexpr int NX = E::NX, ND = E::ND, LD = CB + 1;
  const int32_t vtx = __ldg(p.wv_ids + slot);
  const uint32_t locv = __ldg(p.wv_loc + slot);
  uint32_t bit = 0, locd = 0;
  if (p.bits)
  {
    const uint32_t dof = (uint32_t)__ldg(p.wd_ids + slot);
    locd = __ldg(p.wd_loc + slot);
    bit = (__ldg(p.bits + (dof >> 5)) >> (dof & 31u)) & 1u;
  }
  const double* pp = p.x + 3 * (int64_t)vtx;
  const double px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
  double xc[NX][3];
#pragma unroll
  for (int v = 0; v < NX; ++v)
  {
    const int l = (int)(locv >> (8 * v));
    xc[v][0] = __shfl_sync(0xffffffffu, px, l);
    xc[v][1] = __shfl_sync(0xffffffffu, py, l);
    xc[v][2] = __shfl_sync(0xffffffffu, pz, l);
  }
  uint32_t m = 0;
  if (__any_sync(0xffffffffu, bit))
  {
#pragma unroll
    for (int i = 0; i < ND; ++i)
      m |= __shfl_sync(0xffffffffu, bit, (int)(locd >> (8 * i))) << i;
  }
  if (slot < p.n)
  {
    typename E::Geo g;
    E::prepare(g, xc, nullptr, p.constants, 0);
    double* e = Es + tid;
    if (m == 0)
    {
#pragma unroll
      for (int i = 0; i < ND; ++i)
      {
        double row[ND];
        E::row(g, i, row);
#pragma unroll
        for (int j = i; j < ND; ++j)
          e[(i * ND - i * (i - 1) / 2 + (j - i)) * LD] = row[j];
      }
    }
    else
    {
#pragma unroll
      for (int i = 0; i < ND; ++i)
      {
        double row[ND];
        E::row(g, i, row);
#pragma unroll
        for (int j = i; j < ND; ++j)
          e[(i * ND - i * (i - 1) / 2 + (j - i)) * LD] = ((m >> i) | (m >> j)) & 1u ? 0.0 : row[j];
      }
    }
  }
}

// Lean kernel with ADDRESS-ORDERED write-back (plans built with BFX_CHUNKS_TWO_STAGE): the list walk leaves one sum per
// destination in shared memory, then the chunk's (address, destination) list - both entries of every symmetric pair,
// sorted by address at plan time, plain stores first, REDs after - is walked by consecutive lanes, so that a warp
// updates runs of consecutive CSR values (l1tex wavefronts of the scattered updates: 2.9 per cell, r02 ncu capture).
template <class E, int CB, int WS>
__global__ void __launch_bounds__(CB, chunk_min_ctas(CB)) k_matrix_lean_two(const LeanArgs p, const uint32_t* __restrict__ wr_addr,
                                                                             const uint16_t* __restrict__ wr_src)
{
  constexpr int ND = E::ND, NS = ND * (ND + 1) / 2, LD = CB + 1;
  using L = LeanTwoSmem<NS, CB>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Es = reinterpret_cast<double*>(smem_raw);
  uint16_t* s_src = reinterpret_cast<uint16_t*>(smem_raw + L::SRC_OFF);
  uint32_t* s_waddr = reinterpret_cast<uint32_t*>(smem_raw + L::WADDR_OFF);
  uint16_t* s_wsrc = reinterpret_cast<uint16_t*>(smem_raw + L::WSRC_OFF);
  uint32_t* s_winfo = reinterpret_cast<uint32_t*>(smem_raw + L::WINFO_OFF);
  double* s_sum = reinterpret_cast<double*>(smem_raw + L::SUM_OFF);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L::BAR_OFF);

  const uint32_t tid = threadIdx.x;
  const ChunkHdr h = p.hdr[blockIdx.x];
  const int n_dw = (h.n_dest + 31) >> 5;
  const uint32_t n_wr = (uint32_t)h.pad, n_wr_pad = (n_wr + 31u) & ~31u;
  const bool fits = h.n_src32 <= L::SRC_GROUPS && WS * n_dw <= L::WINFO; // (n_dest <= DCAP: checked when the plan was built)
  const uint16_t* g_src = p.src + (h.src_base32 << 5);
  const uint32_t* g_winfo = p.winfo + (h.dest_base >> 5) * WS;
  const uint32_t* g_waddr = wr_addr + h.dest_base * 2;
  const uint16_t* g_wsrc = wr_src + h.dest_base * 2;
  if (tid == 0)
  {
    mbar_init(bar, 1);
    Es[NS * LD] = 0.0;
    if (fits && n_dw > 0)
    {
      const uint32_t winfo_bytes = ((uint32_t)(n_dw * WS) * 4u + 15u) & ~15u;
      mbar_expect_tx(bar, (uint32_t)h.n_src32 * 64u + winfo_bytes + n_wr_pad * 6u);
      bulk_g2s(s_src, g_src, (uint32_t)h.n_src32 * 64u, bar);
      bulk_g2s(s_winfo, g_winfo, winfo_bytes, bar);
      if (n_wr_pad)
      {
        bulk_g2s(s_waddr, g_waddr, n_wr_pad * 4u, bar);
        bulk_g2s(s_wsrc, g_wsrc, n_wr_pad * 2u, bar);
      }
    }
  }
  lean_stage<E, CB>(p, Es, tid, blockIdx.x * (uint32_t)CB + tid);
  __syncthreads();
  if (fits && n_dw > 0)
    mbar_wait(bar, 0);
  if (fits)
    chunk_walk_sums<CB, WS>(s_src, s_winfo, Es, s_sum, n_dw);
  else
    chunk_walk_sums<CB, WS>(g_src, g_winfo, Es, s_sum, n_dw);
  __syncthreads();
  const uint32_t* wa = fits ? s_waddr : g_waddr;
  const uint16_t* ws = fits ? s_wsrc : g_wsrc;
  for (uint32_t t = tid; t < n_wr; t += CB)
  {
    const uint32_t sx = ws[t];
    double* dst = p.values + wa[t];
    const double v = s_sum[sx & 0x7fffu];
    if (sx & 0x8000u)
      red_add(dst, v);
    else if (p.overwrite)
      *dst = v;
    else
      *dst += v;
  }
}

// ---------------------------------------------------------------------------------------------
// LEAN2: the lean kernel as PERSISTENT CTAs with a two-stage pipeline.  The CTA walks chunks q, q + grid, ...; the
// lists of chunk q + grid (source lists, destination addresses, group table: three TMA bulk copies) land in the
// other shared-memory stage and the table words of its cells are fetched into registers while chunk q is computed,
// so that of the three DRAM round trips of a short-lived CTA (tables, lists, coordinates) only the coordinate
// gather stays exposed.  One __syncthreads per chunk; the staging area is handed back through an mbarrier that every
// thread arrives on after its list walk and waits for just before it stages the next chunk.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NS, int CB>
struct Lean2Smem
{
  static constexpr int SRC_GROUPS = (NS * CB / 32) * 19 / 10; // 32-entry groups of list entries per stage
  static constexpr int DCAP = (NS * CB * 13 / 40) / 32 * 32;   // destinations per stage
  static constexpr int WINFO = DCAP / 32;
  static constexpr size_t E_BYTES = (sizeof(double) * ((size_t)NS * (CB + 1) + 2) + 127) / 128 * 128;
  static constexpr size_t SRC_BYTES = (size_t)SRC_GROUPS * 64, DEST_BYTES = (size_t)DCAP * 8;
  static constexpr size_t WINFO_BYTES = ((size_t)WINFO * 4 + 15) / 16 * 16;
  static constexpr size_t STAGE = SRC_BYTES + DEST_BYTES + WINFO_BYTES;
  static constexpr size_t STAGE_OFF = E_BYTES;
  static constexpr size_t HDR_OFF = STAGE_OFF + 2 * STAGE; // 2 x ChunkHdr
  static constexpr size_t BAR_OFF = HDR_OFF + 2 * sizeof(ChunkHdr);
  static constexpr size_t TOTAL = BAR_OFF + 32;
};

template <class E, int CB>
__global__ void __launch_bounds__(CB, chunk_min_ctas(CB)) k_matrix_lean2(const LeanArgs p, const uint32_t nchunks)
{
  constexpr int NX = E::NX, ND = E::ND, NS = ND * (ND + 1) / 2, LD = CB + 1;
  static_assert(E::BS == 1 && NX <= 4 && ND <= 4 && E::WSIZE == 0, "lean kernel: P1-sized scalar elements without coefficients");
  using L = Lean2Smem<NS, CB>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Es = reinterpret_cast<double*>(smem_raw);
  ChunkHdr* s_hdr = reinterpret_cast<ChunkHdr*>(smem_raw + L::HDR_OFF);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + L::BAR_OFF); // [2]
  uint64_t* bar_free = bar_full + 2;

  const uint32_t tid = threadIdx.x;
  // one thread: header of chunk q -> shared memory, TMA of its lists into stage st
  auto fetch_lists = [&](uint32_t q, int st)
  {
    const ChunkHdr h = p.hdr[q];
    s_hdr[st] = h;
    const int n_dw = (h.n_dest + 31) >> 5;
    const uint32_t src_bytes = (uint32_t)h.n_src32 * 64u, dest_bytes = (uint32_t)n_dw * 256u;
    const bool fits = h.n_src32 <= L::SRC_GROUPS && n_dw <= L::WINFO;
    if (fits && n_dw > 0)
    {
      unsigned char* base = smem_raw + L::STAGE_OFF + (size_t)st * L::STAGE;
      const uint32_t winfo_bytes = ((uint32_t)n_dw * 4u + 15u) & ~15u;
      mbar_expect_tx(bar_full + st, src_bytes + dest_bytes + winfo_bytes);
      bulk_g2s(base, p.src + (h.src_base32 << 5), src_bytes, bar_full + st);
      bulk_g2s(base + L::SRC_BYTES, p.dest + h.dest_base * 2, dest_bytes, bar_full + st);
      bulk_g2s(base + L::SRC_BYTES + L::DEST_BYTES, p.winfo + (h.dest_base >> 5), winfo_bytes, bar_full + st);
    }
  };
  uint32_t q = blockIdx.x;
  if (tid == 0)
  {
    mbar_init(bar_full, 1);
    mbar_init(bar_full + 1, 1);
    mbar_init(bar_free, CB);
    Es[NS * LD] = 0.0; // the slot padded list entries point at
    fetch_lists(q, 0);
  }
  // table words of the first chunk
  uint32_t slot = q * (uint32_t)CB + tid;
  int32_t vtx = __ldg(p.wv_ids + slot);
  uint32_t locv = __ldg(p.wv_loc + slot);
  uint32_t dof = 0, locd = 0;
  if (p.bits)
  {
    dof = (uint32_t)__ldg(p.wd_ids + slot);
    locd = __ldg(p.wd_loc + slot);
  }
  __syncthreads(); // barriers initialised before anybody waits on them

  for (uint32_t it = 0; q < nchunks; ++it, q += gridDim.x)
  {
    const int st = (int)(it & 1u);
    // ---- this chunk: second-level loads (the lane's node, its marker word)
    const double* pp = p.x + 3 * (int64_t)vtx;
    const double px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
    uint32_t bit = 0;
    if (p.bits)
      bit = (__ldg(p.bits + (dof >> 5)) >> (dof & 31u)) & 1u;
    const uint32_t locv_c = locv, locd_c = locd;
    const bool active = slot < p.n;
    // ---- next chunk: table words into registers, lists into the other stage
    const uint32_t qn = q + gridDim.x;
    if (qn < nchunks)
    {
      slot = qn * (uint32_t)CB + tid;
      vtx = __ldg(p.wv_ids + slot);
      locv = __ldg(p.wv_loc + slot);
      if (p.bits)
      {
        dof = (uint32_t)__ldg(p.wd_ids + slot);
        locd = __ldg(p.wd_loc + slot);
      }
      if (tid == 0)
      {
        if (it > 0)
          mbar_wait(bar_free, (it - 1) & 1u); // everybody has left the walk that read this stage
        fetch_lists(qn, st ^ 1);
      }
    }
    // ---- coordinates of this cell's nodes
    double xc[NX][3];
#pragma unroll
    for (int v = 0; v < NX; ++v)
    {
      const int l = (int)(locv_c >> (8 * v));
      xc[v][0] = __shfl_sync(0xffffffffu, px, l);
      xc[v][1] = __shfl_sync(0xffffffffu, py, l);
      xc[v][2] = __shfl_sync(0xffffffffu, pz, l);
    }
    uint32_t m = 0;
    if (__any_sync(0xffffffffu, bit))
    {
#pragma unroll
      for (int i = 0; i < ND; ++i)
        m |= __shfl_sync(0xffffffffu, bit, (int)(locd_c >> (8 * i))) << i;
    }
    typename E::Geo g;
    if (active)
      E::prepare(g, xc, nullptr, p.constants, 0);
    if (it > 0)
      mbar_wait(bar_free, (it - 1) & 1u); // the previous walk no longer reads the staging area
    if (active)
    {
      double* e = Es + tid;
#pragma unroll
      for (int i = 0; i < ND; ++i)
      {
        double row[ND];
        E::row(g, i, row);
#pragma unroll
        for (int j = i; j < ND; ++j)
          e[(i * ND - i * (i - 1) / 2 + (j - i)) * LD] = (m && (((m >> i) | (m >> j)) & 1u)) ? 0.0 : row[j];
      }
    }
    __syncthreads();
    const ChunkHdr h = s_hdr[st];
    const int n_dw = (h.n_dest + 31) >> 5;
    const bool fits = h.n_src32 <= L::SRC_GROUPS && n_dw <= L::WINFO;
    if (fits)
    {
      if (n_dw > 0)
        mbar_wait(bar_full + st, (it >> 1) & 1u);
      const unsigned char* base = smem_raw + L::STAGE_OFF + (size_t)st * L::STAGE;
      chunk_walk<true, CB, uint32_t>(reinterpret_cast<const uint16_t*>(base),
                                     reinterpret_cast<const uint32_t*>(base + L::SRC_BYTES),
                                     reinterpret_cast<const uint32_t*>(base + L::SRC_BYTES + L::DEST_BYTES), Es, p.values,
                                     h.n_dest, h.n_complete, n_dw, p.overwrite);
    }
    else
      chunk_walk<true, CB, uint32_t>(p.src + (h.src_base32 << 5), p.dest + h.dest_base * 2, p.winfo + (h.dest_base >> 5), Es,
                                     p.values, h.n_dest, h.n_complete, n_dw, p.overwrite);
    mbar_arrive(bar_free);
  }
}

// int8 markers -> one bit per dof (32 dofs per word, one warp ballot per word)
__global__ void k_pack_marker_bits(int64_t n, const int8_t* __restrict__ markers, uint32_t* __restrict__ bits)
{
  const int64_t nw = (n + 31) / 32 * 32;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nw; t += (int64_t)gridDim.x * blockDim.x)
  {
    const unsigned b = __ballot_sync(0xffffffffu, t < n && markers[t] != 0);
    if ((t & 31) == 0)
      bits[t >> 5] = b;
  }
}

template <class E, bool SYM, int CB, int THREADS>
int launch_chunked_cb(const bfx_asm* P, const AsmArgs& a, int values_mode, cudaStream_t st)
{
  constexpr int N = E::ND * E::BS, NS = staged_per_cell(N, SYM);
  static_assert(CB > 0, "element too large for the chunked path");
  const bfx_chunks* c = P->chunks;
  if (c->cb != CB || c->n2 != NS)
    return fail(BFX_ERR_INVALID, "chunk plan (cb=%d, staged=%d) does not match the kernel (cb=%d, staged=%d)", c->cb,
                c->n2, CB, NS);
  if (a.n == 0)
    return BFX_OK;
  ChunkArgs ch;
  ch.hdr = c->hdr;
  ch.winfo = c->winfo;
  ch.dest_addr = c->dest_addr;
  ch.src = c->src;
  ch.perm = c->perm;
  ch.xdm = c->xdm;
  ch.dm0 = c->dm0;
  ch.dm1 = c->dm1 ? c->dm1 : c->dm0;
  ch.wv_ids = c->wv_ids, ch.wv_cnt = c->wv_cnt, ch.wv_loc = c->wv_loc;
  ch.wd_ids = c->wd_ids, ch.wd_cnt = c->wd_cnt, ch.wd_loc = c->wd_loc;
  ch.colour = c->colour;
  ch.overwrite = values_mode == BFX_VALUES_OVERWRITE;
  ch.bits0 = ch.bits1 = nullptr;
  ch.same_bc = a.bc0 && a.bc0 == a.bc1 && a.dofmap0 == a.dofmap1;
  if (a.bc0)
  {
    k_pack_marker_bits<<<grid_for(c->n_dofs0, 256, 8), 256, 0, st>>>(c->n_dofs0, a.bc0, c->bits0);
    ch.bits0 = c->bits0;
  }
  if (a.bc1 && !ch.same_bc)
  {
    k_pack_marker_bits<<<grid_for(c->n_dofs1, 256, 8), 256, 0, st>>>(c->n_dofs1, a.bc1, c->bits1);
    ch.bits1 = c->bits1;
  }
  ch.wr_addr = c->wr_addr;
  ch.wr_src = c->wr_src;
  ch.tables_complete = 0;
  ch.chunk_list = nullptr, ch.skip = nullptr;
  unsigned classic_grid = (unsigned)c->nchunks;
  if (c->launch_part != 0 && !c->slim)
  {
    if (!c->part_flag)
      return fail(BFX_ERR_INVALID, "chunk plan has no partition (bfx_asm_chunk_partition)");
    if (c->launch_part == 1)
      ch.chunk_list = c->part_list, classic_grid = (unsigned)c->n_part1;
    else
      ch.skip = c->part_flag;
    if (classic_grid == 0)
      return BFX_OK;
  }
  if constexpr (SYM && E::NX <= 4 && E::ND <= 4 && E::BS == 1 && E::WSIZE == 0 && THREADS == CB)
  {
    // BFX_CHUNK_KERNEL_LEAN: the instruction-lean kernel (linear staging, complete warp tables, 32-bit addresses)
    if (c->kernel_variant == BFX_CHUNK_KERNEL_LEAN && c->addr_bytes == 4 && !c->colour && c->tables_complete
        && P->ncells < 0xffffffffLL - CB && (!a.bc0 || c->wd_ids))
    {
      const bool lean_two = c->wr_addr != nullptr;
      const size_t smem = ChunkSmem<NS, CB, 2>::TOTAL;
      LeanArgs lp;
      lp.hdr = c->hdr, lp.winfo = c->winfo, lp.dest = static_cast<const uint32_t*>(c->dest_addr), lp.src = c->src;
      lp.wv_ids = c->wv_ids, lp.wv_loc = reinterpret_cast<const uint32_t*>(c->wv_loc);
      lp.wd_ids = c->wd_ids, lp.wd_loc = reinterpret_cast<const uint32_t*>(c->wd_loc);
      lp.bits = ch.bits0;
      lp.x = a.x;
      lp.values = a.values;
      for (int k = 0; k < 4; ++k)
        lp.constants[k] = a.constants[k];
      lp.n = (uint32_t)a.n;
      lp.overwrite = ch.overwrite;
      lp.chunk_begin = 0;
      lp.valid_in_hdr = c->part_rows >= 0 ? 1 : 0;
      unsigned lean_grid = (unsigned)c->nchunks;
      if (c->launch_part != 0 && c->part_rows < 0)
        return fail(BFX_ERR_INVALID, "chunk plan has no partition (bfx_asm_chunk_partition)");
      if (c->launch_part == 1)
        lean_grid = (unsigned)c->n_part1;
      else if (c->launch_part == 2)
        lp.chunk_begin = (uint32_t)c->n_part1, lean_grid = (unsigned)(c->nchunks - c->n_part1);
      if (lean_grid == 0)
        return BFX_OK;
      const int dbg = c->lean_dbg;
      if (c->launch_part != 0 && (c->wr_addr != nullptr || dbg != 0))
        return fail(BFX_ERR_UNSUPPORTED, "partial launches are implemented for the default lean kernel");
      if (lean_two)
      {
        using LT = LeanTwoSmem<NS, CB>;
        if (c->len_sorted)
        {
          BFX_CUDA(cudaFuncSetAttribute(k_matrix_lean_two<E, CB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT::TOTAL));
          k_matrix_lean_two<E, CB, 2><<<(unsigned)c->nchunks, CB, LT::TOTAL, st>>>(lp, c->wr_addr, c->wr_src);
        }
        else
        {
          BFX_CUDA(cudaFuncSetAttribute(k_matrix_lean_two<E, CB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT::TOTAL));
          k_matrix_lean_two<E, CB, 1><<<(unsigned)c->nchunks, CB, LT::TOTAL, st>>>(lp, c->wr_addr, c->wr_src);
        }
      }
      else if (dbg == 10) // persistent two-stage pipeline
      {
        using L2 = Lean2Smem<NS, CB>;
        BFX_CUDA(cudaFuncSetAttribute(k_matrix_lean2<E, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L2::TOTAL));
        int per_sm = 0;
        BFX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_matrix_lean2<E, CB>, CB, L2::TOTAL));
        const int64_t grid = std::min<int64_t>(c->nchunks, (int64_t)sm_count() * std::max(per_sm, 1));
        k_matrix_lean2<E, CB><<<(unsigned)grid, CB, L2::TOTAL, st>>>(lp, (uint32_t)c->nchunks);
      }
      else if (dbg == 20 && CB == 256) // 5 resident CTAs per SM
      {
        if constexpr (CB == 256)
        {
          BFX_CUDA(cudaFuncSetAttribute(k_matrix_lean<E, CB, 0, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          k_matrix_lean<E, CB, 0, 5><<<(unsigned)c->nchunks, CB, smem, st>>>(lp);
        }
      }
      else if (c->len_sorted)
      {
        BFX_CUDA(cudaFuncSetAttribute(k_matrix_lean<E, CB, 0, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_matrix_lean<E, CB, 0, 0, true><<<lean_grid, CB, smem, st>>>(lp);
      }
      else if (dbg == 0)
      {
        BFX_CUDA(cudaFuncSetAttribute(k_matrix_lean<E, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_matrix_lean<E, CB><<<lean_grid, CB, smem, st>>>(lp);
      }
      else if constexpr (CB == 256 && E::NX == 4)
      {
        auto go = [&](auto d)
        {
          constexpr int D = decltype(d)::value;
          cudaFuncSetAttribute(k_matrix_lean<E, CB, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          k_matrix_lean<E, CB, D><<<(unsigned)c->nchunks, CB, smem, st>>>(lp);
        };
        if (dbg == 1)
          go(std::integral_constant<int, 1>());
        else if (dbg == 2)
          go(std::integral_constant<int, 2>());
        else
          go(std::integral_constant<int, 3>());
      }
      BFX_CHECK_LAUNCH();
      return BFX_OK;
    }
  }
  if (c->launch_part != 0 && (c->wr_addr || c->kernel_variant == BFX_CHUNK_KERNEL_OCC5 || (c->kernel_variant >= 11 && c->kernel_variant <= 14)))
    return fail(BFX_ERR_UNSUPPORTED, "partial launches are implemented for the default / DIET / lean kernels");
  if (c->slim)
    return fail(BFX_ERR_INVALID, "chunk plan reduced for the lean kernel: this call (element, addresses or markers) needs the full plan");
  if (c->len_sorted)
    return fail(BFX_ERR_UNSUPPORTED, "chunk plan ordered by list length (BFX_CHUNKS_LEN_SORT) needs the lean kernel");
  if constexpr (SYM && NS <= 16 && CB == 256)
  {
    if (c->wr_addr) // two-stage plan (32-bit addresses)
    {
      const size_t smem2 = ChunkSmem<NS, CB, 2, true>::TOTAL;
      BFX_CUDA(cudaFuncSetAttribute(k_matrix_chunked<E, SYM, CB, THREADS, uint32_t, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      k_matrix_chunked<E, SYM, CB, THREADS, uint32_t, true><<<(unsigned)c->nchunks, THREADS, smem2, st>>>(a, ch);
      BFX_CHECK_LAUNCH();
      return BFX_OK;
    }
  }
  if constexpr (SYM && two_stage_lists_global(NS) && CB == chunk_cb(NS))
  {
    if (c->wr_addr) // two-stage plan of a large element: DIET list walk, write-back lists streamed from global memory
    {
      const size_t smem2 = ChunkSmem<NS, CB, 2, true>::TOTAL;
      BFX_CUDA(cudaFuncSetAttribute(k_matrix_chunked<E, SYM, CB, THREADS, uint32_t, true, 0, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      k_matrix_chunked<E, SYM, CB, THREADS, uint32_t, true, 0, true><<<(unsigned)c->nchunks, THREADS, smem2, st>>>(a, ch);
      BFX_CHECK_LAUNCH();
      return BFX_OK;
    }
  }
  if (c->wr_addr)
    return fail(BFX_ERR_INVALID, "two-stage chunk plan without a two-stage kernel");
  const size_t smem = ChunkSmem<NS, CB, SYM ? 2 : 1>::TOTAL;
  if constexpr (SYM && NS <= 16 && CB == 256)
  {
    // BFX_CHUNK_KERNEL_OCC5: the variant compiled for 5 resident CTAs per SM (48 registers)
    if (c->kernel_variant == BFX_CHUNK_KERNEL_OCC5 && c->addr_bytes == 4)
    {
      BFX_CUDA(cudaFuncSetAttribute(k_matrix_chunked<E, SYM, CB, THREADS, uint32_t, false, 5>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_matrix_chunked<E, SYM, CB, THREADS, uint32_t, false, 5><<<(unsigned)c->nchunks, THREADS, smem, st>>>(a, ch);
      BFX_CHECK_LAUNCH();
      return BFX_OK;
    }
  }
  if constexpr (SYM && CB == chunk_cb(NS) && THREADS == chunk_threads(NS))
  {
    // profiling-only variants (10 + DBG): see k_matrix_chunked
    if (c->kernel_variant >= 11 && c->kernel_variant <= 14 && c->addr_bytes == 4)
    {
      auto go = [&](auto dbg)
      {
        constexpr int D = decltype(dbg)::value;
        cudaFuncSetAttribute(k_matrix_chunked<E, SYM, CB, THREADS, uint32_t, false, 0, true, D>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_matrix_chunked<E, SYM, CB, THREADS, uint32_t, false, 0, true, D><<<(unsigned)c->nchunks, THREADS, smem, st>>>(a, ch);
      };
      if (c->kernel_variant == 11)
        go(std::integral_constant<int, 1>());
      else if (c->kernel_variant == 12)
        go(std::integral_constant<int, 2>());
      else if (c->kernel_variant == 13)
        go(std::integral_constant<int, 3>());
      else
        go(std::integral_constant<int, 4>());
      BFX_CHECK_LAUNCH();
      return BFX_OK;
    }
  }
  if constexpr (SYM && (CB == chunk_cb(NS) || two_stage_lists_global(NS)))
  {
    // BFX_CHUNK_KERNEL_DIET: the variant whose phase 2 is chunk_walk (the default of the large elements)
    if ((c->kernel_variant == BFX_CHUNK_KERNEL_DIET || c->kernel_variant == BFX_CHUNK_KERNEL_WIDE
         || (two_stage_lists_global(NS) && c->kernel_variant == BFX_CHUNK_KERNEL_DEFAULT))
        && c->addr_bytes == 4)
    {
      BFX_CUDA(cudaFuncSetAttribute(k_matrix_chunked<E, SYM, CB, THREADS, uint32_t, false, 0, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_matrix_chunked<E, SYM, CB, THREADS, uint32_t, false, 0, true><<<classic_grid, THREADS, smem, st>>>(a, ch);
      BFX_CHECK_LAUNCH();
      return BFX_OK;
    }
  }
  if (c->addr_bytes == 4)
  {
    BFX_CUDA(cudaFuncSetAttribute(k_matrix_chunked<E, SYM, CB, THREADS, uint32_t>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_matrix_chunked<E, SYM, CB, THREADS, uint32_t><<<classic_grid, THREADS, smem, st>>>(a, ch);
  }
  else
  {
    BFX_CUDA(cudaFuncSetAttribute(k_matrix_chunked<E, SYM, CB, THREADS, uint64_t>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_matrix_chunked<E, SYM, CB, THREADS, uint64_t><<<classic_grid, THREADS, smem, st>>>(a, ch);
  }
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

// the plan fixes the chunk size (bfx_asm_build_chunks: the element's default or BFX_CHUNKS_CB(cells))
template <class E, bool SYM, int CB>
int launch_chunked_try(const bfx_asm* P, const AsmArgs& a, int values_mode, cudaStream_t st, int& status)
{
  constexpr int NS = staged_per_cell(E::ND * E::BS, SYM);
  if constexpr (chunk_cb_supported(NS, CB))
  {
    if (P->chunks->cb == CB)
    {
      status = launch_chunked_cb<E, SYM, CB, chunk_threads_for(NS, CB)>(P, a, values_mode, st);
      return 1;
    }
  }
  return 0;
}

template <class E, bool SYM>
int launch_chunked_es(const bfx_asm* P, const AsmArgs& a, int values_mode, cudaStream_t st)
{
  constexpr int NS = staged_per_cell(E::ND * E::BS, SYM);
  int status = BFX_OK;
  if (launch_chunked_try<E, SYM, 64>(P, a, values_mode, st, status) || launch_chunked_try<E, SYM, 96>(P, a, values_mode, st, status)
      || launch_chunked_try<E, SYM, 128>(P, a, values_mode, st, status)
      || launch_chunked_try<E, SYM, 192>(P, a, values_mode, st, status)
      || launch_chunked_try<E, SYM, 384>(P, a, values_mode, st, status))
    return status;
  if constexpr (SYM && NS > 36)
  {
    // BFX_CHUNK_KERNEL_WIDE: four threads per cell (twice the warps per SM at the same shared memory) + chunk_walk
    if (P->chunks->kernel_variant == BFX_CHUNK_KERNEL_WIDE)
      return launch_chunked_cb<E, SYM, chunk_cb(NS), 4 * chunk_cb(NS)>(P, a, values_mode, st);
  }
  return launch_chunked_cb<E, SYM, chunk_cb(NS), chunk_threads(NS)>(P, a, values_mode, st);
}

// The symmetric plan needs a symmetric element matrix (all chunked kernels are) AND symmetric bc zeroing:
// the same dofmap and the same markers for rows and columns.  BFX_ERR_UNSUPPORTED lets the caller fall back.
template <class E>
int launch_chunked_e(const bfx_asm* P, const AsmArgs& a, int values_mode, cudaStream_t st)
{
  if (!P->chunks->sym)
    return launch_chunked_es<E, false>(P, a, values_mode, st);
  if (a.dofmap0 != a.dofmap1 || a.bc0 != a.bc1)
    return fail(BFX_ERR_UNSUPPORTED, "symmetric chunk plan needs one dofmap and one marker array for rows and columns");
  return launch_chunked_es<E, true>(P, a, values_mode, st);
}

// Morton order of the centroids of the plan's cells: perm[slot] = entity
int morton_perm(const bfx_asm* P, const double* x_dev, int32_t** perm_out, cudaStream_t st)
{
  unsigned long long* bb = nullptr;
  uint64_t *k0 = nullptr, *k1 = nullptr;
  int32_t* i0 = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  const int64_t n = P->ncells;
  int e;
  if ((e = dev_alloc(&bb, 6)) || (e = dev_alloc(&k0, (size_t)n)) || (e = dev_alloc(&k1, (size_t)n))
      || (e = dev_alloc(&i0, (size_t)n)) || (e = dev_alloc(perm_out, (size_t)n)))
  {
    cudaFree(bb), cudaFree(k0), cudaFree(k1), cudaFree(i0), cudaFree(*perm_out);
    *perm_out = nullptr;
    return e;
  }
  const unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0, 0, 0};
  BFX_CUDA(cudaMemcpyAsync(bb, init, sizeof(init), cudaMemcpyHostToDevice, st));
  k_centroid_bbox<<<grid_for(n, 256, 8), 256, 0, st>>>(n, P->cells, P->x_dofmap, P->nx, x_dev, bb);
  k_morton_keys<<<grid_for(n, 256, 8), 256, 0, st>>>(n, P->cells, P->x_dofmap, P->nx, x_dev, bb, k0, i0);
  BFX_CHECK_LAUNCH();
  auto release = [&]()
  {
    cudaFree(tmp);
    cudaFree(bb);
    cudaFree(k0);
    cudaFree(k1);
    cudaFree(i0);
  };
  cudaError_t ce = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, i0, *perm_out, n, 0, 63, st);
  if (ce == cudaSuccess)
    ce = cudaMalloc(&tmp, tmp_bytes);
  if (ce == cudaSuccess)
    ce = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, i0, *perm_out, n, 0, 63, st);
  if (ce == cudaSuccess)
    ce = cudaStreamSynchronize(st);
  release();
  if (ce != cudaSuccess)
  {
    (void)cudaGetLastError();
    cudaFree(*perm_out);
    *perm_out = nullptr;
    return fail(BFX_ERR_CUDA, "Morton ordering of the cell list failed: %s", cudaGetErrorString(ce));
  }
  return BFX_OK;
}

// ---- vector plans: per group of 32 cell slots, which (cell, local dof) pairs land on each distinct dof ----
__global__ void __launch_bounds__(128)
    k_group_lists(int64_t ngroups, int nd, const uint8_t* __restrict__ wd_cnt, const uint8_t* __restrict__ wd_loc,
                  int64_t n, uint8_t* __restrict__ glist, uint8_t* __restrict__ goff)
{
  __shared__ int s_cnt[4][32], s_off[4][33], s_cur[4][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int locw = 4 * ((nd + 3) / 4);
  for (int64_t g = (int64_t)blockIdx.x * 4 + wib; g < ngroups; g += (int64_t)gridDim.x * 4)
  {
    const int64_t slot = g * 32 + lane;
    const int dcnt = wd_cnt[g];
    s_cnt[wib][lane] = 0;
    s_cur[wib][lane] = 0;
    __syncwarp();
    if (dcnt && slot < n)
      for (int i = 0; i < nd; ++i)
        atomicAdd(&s_cnt[wib][wd_loc[slot * locw + i]], 1);
    __syncwarp();
    int incl = s_cnt[wib][lane];
    for (int o = 1; o < 32; o <<= 1)
    {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o)
        incl += v;
    }
    s_off[wib][lane + 1] = incl;
    if (lane == 0)
      s_off[wib][0] = 0;
    __syncwarp();
    // cells in ascending lane order inside every dof: a fixed summation order
    if (dcnt)
      for (int c = 0; c < 32; ++c)
      {
        if (lane == c && slot < n)
          for (int i = 0; i < nd; ++i)
          {
            const int l = wd_loc[slot * locw + i];
            glist[g * 32 * nd + s_off[wib][l] + s_cur[wib][l]++] = (uint8_t)(i * 32 + lane);
          }
        __syncwarp();
      }
    goff[g * 36 + lane] = (uint8_t)s_off[wib][lane];
    if (lane == 0)
      goff[g * 36 + 32] = (uint8_t)min(s_off[wib][32], 255);
    __syncwarp();
  }
}

// Grouped vector assembly (fem::impl::assemble_cells, fem/assemble_vector_impl.h:72-116) for P1-sized elements:
// one warp per group of 32 cells.  Coordinates and coefficient dofs come through the warp tables (one load per
// distinct node / dof, shuffles), the element vectors are staged in shared memory, and one lane per DISTINCT dof
// of the group sums its contributions and issues ONE RED (4 per cell in the cell-parallel kernel).
template <class E>
__global__ void __launch_bounds__(256) k_vector_grouped(const AsmArgs a, const ChunkArgs ch, const uint8_t* __restrict__ glist,
                                                        const uint8_t* __restrict__ goff, int64_t ngroups)
{
  constexpr int NX = E::NX, ND = E::ND;
  static_assert(E::BS == 1 && ND <= 4 && NX <= 4, "grouped vector kernel: P1-sized scalar elements");
  constexpr int LOCWV = (NX + 3) / 4, LOCWD = (ND + 3) / 4;
  __shared__ double s_be[8][ND][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int64_t g = (int64_t)blockIdx.x * 8 + wib; g < ngroups; g += (int64_t)gridDim.x * 8)
  {
    const int64_t slot = g * 32 + lane;
    const bool active = slot < a.n;
    const int vcnt = ch.wv_cnt[g], dcnt = ch.wd_cnt[g];
    const int32_t vtx_id = ch.wv_ids[g * 32 + lane];
    const int64_t dof_id = ch.wd_ids[g * 32 + lane];
    uint32_t locv[LOCWV], locd[LOCWD];
#pragma unroll
    for (int k = 0; k < LOCWV; ++k)
      locv[k] = __ldg(reinterpret_cast<const uint32_t*>(ch.wv_loc) + slot * LOCWV + k);
#pragma unroll
    for (int k = 0; k < LOCWD; ++k)
      locd[k] = __ldg(reinterpret_cast<const uint32_t*>(ch.wd_loc) + slot * LOCWD + k);
    const int my_lo = goff[g * 36 + lane], my_hi = goff[g * 36 + lane + 1];
    // the group's list (32 * ND bytes) travels in registers: one word per lane, bytes fetched with shuffles
    const uint32_t lword = lane < 8 * ND ? __ldg(reinterpret_cast<const uint32_t*>(glist + g * 32 * ND) + lane) : 0u;
    int64_t e = slot;
    int32_t cell = (int32_t)slot;
    if (active && (!vcnt || !dcnt || E::WSIZE > 0))
    {
      e = ch.perm ? ch.perm[slot] : slot;
      cell = a.cells ? a.cells[e] : (int32_t)e;
    }
    // coordinates
    double xc[NX][3];
    if (vcnt)
    {
      const double* pp = a.x + 3 * (int64_t)vtx_id;
      const double px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
#pragma unroll
      for (int v = 0; v < NX; ++v)
      {
        const int l = (int)((locv[v >> 2] >> (8 * (v & 3))) & 31u);
        xc[v][0] = __shfl_sync(0xffffffffu, px, l);
        xc[v][1] = __shfl_sync(0xffffffffu, py, l);
        xc[v][2] = __shfl_sync(0xffffffffu, pz, l);
      }
    }
    else if (active)
    {
      int32_t xd[NX];
      load_ints<NX>(ch.xdm ? ch.xdm + slot * NX : a.x_dofmap + (int64_t)cell * NX, xd);
      gather_coords<NX>(a.x, xd, xc);
    }
    // coefficient: through the dof table when it lives on the plan's dofmap, else the general gather
    double w[E::WSIZE > 0 ? E::WSIZE : 1];
    if constexpr (E::WSIZE > 0)
    {
      const bool same_map = !a.coef.packed && a.coef.f[0].dm == a.dofmap0 && E::WND == ND && E::WBS == 1;
      if (dcnt && same_map)
      {
        const double fv = __ldg(a.coef.f[0].v + dof_id);
#pragma unroll
        for (int i = 0; i < ND; ++i)
          w[i] = __shfl_sync(0xffffffffu, fv, (locd[i >> 2] >> (8 * (i & 3))) & 31u);
      }
      else if (active)
        load_w<E>(a, e, cell, w);
    }
    double out[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i)
      out[i] = 0.0;
    if (active)
    {
      typename E::Geo geo;
      E::prepare(geo, xc, w, a.constants, 0);
      E::vec(geo, out);
    }
    if (dcnt)
    {
#pragma unroll
      for (int i = 0; i < ND; ++i)
        s_be[wib][i][lane] = out[i];
      __syncwarp();
      {
        const double* be = &s_be[wib][0][0];
        const int mine = lane < dcnt ? my_hi - my_lo : 0;
        int longest = mine;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, o));
        double sum = 0.0;
        for (int k = 0; k < longest; ++k) // warp-uniform trip count: every lane takes part in the shuffles
        {
          const int t = k < mine ? my_lo + k : 0;
          const uint32_t wv = __shfl_sync(0xffffffffu, lword, t >> 2);
          if (k < mine)
            sum += be[(wv >> (8 * (t & 3))) & 0xffu];
        }
        if (mine > 0)
          red_add(a.b + dof_id, sum);
      }
      __syncwarp();
    }
    else if (active)
    {
      int32_t d0[ND];
      load_ints<ND>(ch.dm0 ? ch.dm0 + slot * ND : a.dofmap0 + (int64_t)cell * ND, d0);
#pragma unroll
      for (int i = 0; i < ND; ++i)
        red_add(a.b + d0[i], out[i]);
    }
  }
}

// Chunk-aggregated VECTOR assembly (fem::impl::assemble_cells, fem/assemble_vector_impl.h:72-116): the matrix
// machinery on a vector - element vectors staged in shared memory per chunk of CB cells, one sum per distinct
// (dof, component) of the chunk through the plan's source lists, a plain load/add/store for the dofs whose cells all
// lie in the chunk and ONE RED per chunk-boundary dof (the cell-parallel kernel issues one RED per (cell, local dof)).
// The plan is the chunk plan of an identity "matrix" (row = dof, one column): bfx_asm_build_chunks_vector.
template <class E, int CB, typename AddrT>
__global__ void __launch_bounds__(CB, (E::ND * E::BS > 16) ? 1 : chunk_min_ctas(CB)) k_vector_chunked(const AsmArgs a, const ChunkArgs ch)
{
  constexpr int NX = E::NX, ND = E::ND, N = ND * E::BS, LD = CB + 1;
  using L = ChunkSmem<N, CB, 1>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Es = reinterpret_cast<double*>(smem_raw);
  uint16_t* s_src = reinterpret_cast<uint16_t*>(smem_raw + L::SRC_OFF);
  AddrT* s_dest = reinterpret_cast<AddrT*>(smem_raw + L::DEST_OFF);
  uint32_t* s_winfo = reinterpret_cast<uint32_t*>(smem_raw + L::WINFO_OFF);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L::BAR_OFF);

  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t q = blockIdx.x;
  const int64_t slot = q * CB + tid;
  const int64_t gw = slot >> 5;
  // ---- first-level loads
  const int vcnt = (ch.tables_complete && NX <= 4) ? 32 : (ch.wv_cnt ? ch.wv_cnt[gw] : 0);
  const int32_t vtx = vcnt ? __ldg(ch.wv_ids + slot) : 0;
  const uint32_t locv = (vcnt && NX <= 4) ? __ldg(reinterpret_cast<const uint32_t*>(ch.wv_loc) + slot) : 0u;
  // coefficient on the plan's own dofmap (P1-sized scalar elements): through the dof table, like the coordinates
  constexpr bool DOF_TAB = E::WSIZE > 0 && E::WND == ND && E::WBS == 1 && E::BS == 1 && ND <= 4;
  int dcnt = 0;
  int32_t dof = 0;
  uint32_t locd = 0;
  if constexpr (DOF_TAB)
  {
    if (ch.wd_cnt && !a.coef.packed && a.coef.f[0].dm == a.dofmap0)
    {
      dcnt = ch.tables_complete ? 32 : ch.wd_cnt[gw];
      if (dcnt)
      {
        dof = __ldg(ch.wd_ids + slot);
        locd = __ldg(reinterpret_cast<const uint32_t*>(ch.wd_loc) + slot);
      }
    }
  }
  const ChunkHdr h = ch.hdr[q];
  const int n_dw = (h.n_dest + 31) >> 5;
  const uint32_t src_bytes = (uint32_t)h.n_src32 * 64u, dest_bytes = (uint32_t)n_dw * 32u * (uint32_t)sizeof(AddrT);
  const bool fits = h.n_src32 <= L::SRC_GROUPS && dest_bytes <= (uint32_t)L::DEST_BYTES && n_dw <= L::WINFO;
  const uint16_t* g_src = ch.src + (h.src_base32 << 5);
  const AddrT* g_dest = static_cast<const AddrT*>(ch.dest_addr) + h.dest_base;
  const uint32_t* g_winfo = ch.winfo + (h.dest_base >> 5);
  if (tid == 0)
  {
    mbar_init(bar, 1);
    Es[N * LD] = 0.0;
    if (fits && n_dw > 0)
    {
      const uint32_t winfo_bytes = ((uint32_t)n_dw * 4u + 15u) & ~15u;
      mbar_expect_tx(bar, src_bytes + dest_bytes + winfo_bytes);
      bulk_g2s(s_src, g_src, src_bytes, bar);
      bulk_g2s(s_dest, g_dest, dest_bytes, bar);
      bulk_g2s(s_winfo, g_winfo, winfo_bytes, bar);
    }
  }
  const bool active = slot < a.n;
  int64_t e = slot;
  int32_t cell = (int32_t)slot;
  const bool need_cell = !(vcnt && NX <= 4) || (E::WSIZE > 0 && !dcnt); // direct gathers index the caller's arrays
  if (active && need_cell && (ch.perm || a.cells))
  {
    e = ch.perm ? ch.perm[slot] : slot;
    cell = a.cells ? a.cells[e] : (int32_t)e;
  }
  // ---- coordinates
  double xc[NX][3];
  if (vcnt && NX <= 4) // warp-uniform
  {
    const double* pp = a.x + 3 * (int64_t)vtx;
    const double px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
#pragma unroll
    for (int v = 0; v < NX; ++v)
    {
      const int l = (int)(locv >> (8 * (v & 3)));
      xc[v][0] = __shfl_sync(0xffffffffu, px, l);
      xc[v][1] = __shfl_sync(0xffffffffu, py, l);
      xc[v][2] = __shfl_sync(0xffffffffu, pz, l);
    }
  }
  else if (active)
  {
    int32_t xd[NX];
    load_ints<NX>(ch.xdm ? ch.xdm + slot * NX : a.x_dofmap + (int64_t)cell * NX, xd);
    gather_coords<NX>(a.x, xd, xc);
  }
  // ---- coefficient
  double w[E::WSIZE > 0 ? E::WSIZE : 1];
  if constexpr (E::WSIZE > 0)
  {
    bool done = false;
    if constexpr (DOF_TAB)
    {
      if (dcnt) // warp-uniform
      {
        const double fv = __ldg(a.coef.f[0].v + dof);
#pragma unroll
        for (int i = 0; i < ND; ++i)
          w[i] = __shfl_sync(0xffffffffu, fv, (int)(locd >> (8 * i)));
        done = true;
      }
    }
    if (!done && active)
      load_w<E>(a, e, cell, w);
  }
  if (active)
  {
    typename E::Geo g;
    E::prepare(g, xc, w, a.constants, 0);
    double out[N];
    E::vec(g, out);
    double* es = Es + tid;
#pragma unroll
    for (int k = 0; k < N; ++k)
      es[k * LD] = out[k];
  }
  __syncthreads();
  if (fits && n_dw > 0)
    mbar_wait(bar, 0);
  if (fits)
    chunk_walk<false, CB, AddrT>(s_src, s_dest, s_winfo, Es, a.b, h.n_dest, h.n_complete, n_dw, 0);
  else
    chunk_walk<false, CB, AddrT>(g_src, g_dest, g_winfo, Es, a.b, h.n_dest, h.n_complete, n_dw, 0);
  (void)lane;
}

// Table kernel (round 2): P1-sized scalar elements.  One thread per cell, no shared memory, no barrier: the node this
// lane fetches for its group of 32 cells and the coefficient value of the dof it fetches travel to the cells by warp
// shuffles (one coordinate gather per LANE instead of four per cell: 824 M -> ~250 M load sectors at C2), the element
// vector leaves by one RED per (cell, dof).  Grid-stride over the groups at full occupancy: the three dependent load
// levels are hidden by the other warps instead of by a CTA's phases.
template <class E>
__global__ void __launch_bounds__(256) k_vector_tables(const AsmArgs a, const ChunkArgs ch, const int64_t nslots_pad)
{
  constexpr int NX = E::NX, ND = E::ND;
  static_assert(NX <= 4 && ND <= 4 && E::BS == 1, "table kernel: P1-sized scalar elements");
  constexpr bool HAS_W = E::WSIZE > 0;
  for (int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; (slot & ~(int64_t)31) < nslots_pad;
       slot += (int64_t)gridDim.x * blockDim.x)
  {
    const int32_t vtx = __ldg(ch.wv_ids + slot);
    const uint32_t locv = __ldg(reinterpret_cast<const uint32_t*>(ch.wv_loc) + slot);
    const int32_t dof = __ldg(ch.wd_ids + slot);
    const uint32_t locd = __ldg(reinterpret_cast<const uint32_t*>(ch.wd_loc) + slot);
    const double* pp = a.x + 3 * (int64_t)vtx;
    const double px = __ldg(pp), py = __ldg(pp + 1), pz = __ldg(pp + 2);
    double fv = 0.0;
    if constexpr (HAS_W)
      fv = __ldg(a.coef.f[0].v + dof);
    double xc[NX][3];
#pragma unroll
    for (int v = 0; v < NX; ++v)
    {
      const int l = (int)(locv >> (8 * v));
      xc[v][0] = __shfl_sync(0xffffffffu, px, l);
      xc[v][1] = __shfl_sync(0xffffffffu, py, l);
      xc[v][2] = __shfl_sync(0xffffffffu, pz, l);
    }
    double w[HAS_W ? E::WSIZE : 1];
    int32_t d[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i)
    {
      const int l = (int)(locd >> (8 * i));
      if constexpr (HAS_W)
        w[i] = __shfl_sync(0xffffffffu, fv, l);
      d[i] = __shfl_sync(0xffffffffu, dof, l);
    }
    if (slot < a.n)
    {
      typename E::Geo g;
      E::prepare(g, xc, w, a.constants, 0);
      double out[ND];
      E::vec(g, out);
#pragma unroll
      for (int i = 0; i < ND; ++i)
        red_add(a.b + d[i], out[i]);
    }
  }
}

template <class E>
int launch_vector_chunked_e(const bfx_asm* P, const AsmArgs& a, cudaStream_t st)
{
  constexpr int N = E::ND * E::BS;
  constexpr int CB = N <= 10 ? 384 : 256;
  const bfx_chunks* c = P->chunks;
  if (a.n == 0)
    return BFX_OK;
  if (!c->vector_plan || c->cb != CB || c->n2 != N || c->colour || c->sym)
    return fail(BFX_ERR_INVALID, "vector chunk plan (cb=%d, staged=%d) does not match the kernel (cb=%d, staged=%d)", c->cb,
                c->n2, CB, N);
  ChunkArgs ch;
  memset(&ch, 0, sizeof(ch));
  ch.hdr = c->hdr, ch.winfo = c->winfo, ch.dest_addr = c->dest_addr, ch.src = c->src;
  ch.perm = c->perm;
  ch.xdm = c->xdm, ch.dm0 = c->dm0;
  ch.wv_ids = c->wv_ids, ch.wv_cnt = c->wv_cnt, ch.wv_loc = c->wv_loc;
  ch.wd_ids = c->wd_ids, ch.wd_cnt = c->wd_cnt, ch.wd_loc = c->wd_loc;
  ch.tables_complete = c->tables_complete && (P->nd0 > 4 || c->wd_ids);
  if constexpr (E::NX <= 4 && E::ND <= 4 && E::BS == 1 && (E::WSIZE == 0 || (E::WND == E::ND && E::WBS == 1 && E::WSIZE == E::ND)))
  {
    // table kernel: complete tables, the coefficient (if any) lives on the plan's own dofmap and is gathered, not packed
    const bool coef_ok = E::WSIZE == 0 || (!a.coef.packed && a.coef.f[0].dm == a.dofmap0);
    if (ch.tables_complete && c->wv_ids && c->wd_ids && coef_ok && !getenv("BFX_VECTOR_NO_TABLES"))
    {
      const int64_t nslots_pad = c->nchunks * (int64_t)c->cb;
      const unsigned grid = (unsigned)std::min<int64_t>((nslots_pad + 255) / 256, (int64_t)sm_count() * 8);
      k_vector_tables<E><<<grid, 256, 0, st>>>(a, ch, nslots_pad);
      BFX_CHECK_LAUNCH();
      return BFX_OK;
    }
  }
  if (c->slim)
    return fail(BFX_ERR_UNSUPPORTED, "vector chunk plan reduced to its warp tables: this call (packed or foreign coefficient) needs the RED kernel");
  const size_t smem = ChunkSmem<N, CB, 1>::TOTAL;
  if (c->addr_bytes == 4)
  {
    BFX_CUDA(cudaFuncSetAttribute(k_vector_chunked<E, CB, uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_vector_chunked<E, CB, uint32_t><<<(unsigned)c->nchunks, CB, smem, st>>>(a, ch);
  }
  else
  {
    BFX_CUDA(cudaFuncSetAttribute(k_vector_chunked<E, CB, uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_vector_chunked<E, CB, uint64_t><<<(unsigned)c->nchunks, CB, smem, st>>>(a, ch);
  }
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}

template <class E>
int launch_vector_grouped_e(const bfx_asm* P, const AsmArgs& a, cudaStream_t st)
{
  const bfx_chunks* c = P->chunks;
  if (a.n == 0)
    return BFX_OK;
  ChunkArgs ch;
  memset(&ch, 0, sizeof(ch));
  ch.perm = c->perm;
  ch.xdm = c->xdm;
  ch.dm0 = c->dm0;
  ch.wv_ids = c->wv_ids, ch.wv_cnt = c->wv_cnt, ch.wv_loc = c->wv_loc;
  ch.wd_ids = c->wd_ids, ch.wd_cnt = c->wd_cnt, ch.wd_loc = c->wd_loc;
  const int64_t ngroups = c->nchunks;
  k_vector_grouped<E><<<grid_for(ngroups, 8, 0), 256, 0, st>>>(a, ch, c->glist, c->goff, ngroups);
  BFX_CHECK_LAUNCH();
  return BFX_OK;
}
} // namespace

namespace bfx
{
int chunked_supported(int kernel_id)
{
  switch (kernel_id)
  {
  case BFX_K_LAPLACE_P1_TRI_A:
  case BFX_K_MASS_COEFF_P1_TRI_A:
  case BFX_K_POISSON_P1_TET_A:
  case BFX_K_POISSON_P2_TET_A: return 1;
  default: return 0;
  }
}

int launch_chunked(const bfx_asm* P, int kernel_id, const AsmArgs& a, int values_mode, cudaStream_t st)
{
  if (!P->chunks)
    return fail(BFX_ERR_INVALID, "BFX_ASM_CHUNKED needs bfx_asm_build_chunks() on the plan first");
  switch (kernel_id)
  {
  case BFX_K_LAPLACE_P1_TRI_A: return launch_chunked_e<el::LaplaceP1Tri>(P, a, values_mode, st);
  case BFX_K_MASS_COEFF_P1_TRI_A: return launch_chunked_e<el::MassCoeffP1Tri>(P, a, values_mode, st);
  case BFX_K_POISSON_P1_TET_A: return launch_chunked_e<el::PoissonP1Tet>(P, a, values_mode, st);
  case BFX_K_POISSON_P2_TET_A: return launch_chunked_e<el::PoissonP2Tet>(P, a, values_mode, st);
  default: return fail(BFX_ERR_UNSUPPORTED, "kernel id %d has no chunk-aggregated variant", kernel_id);
  }
}

int grouped_vector_supported(int kernel_id)
{
  switch (kernel_id)
  {
  case BFX_K_SOURCE_P1_TRI_L:
  case BFX_K_LOAD_COEFF_P1_TRI_L:
  case BFX_K_LOAD_P1_TET_L:
  case BFX_K_ACTION_POISSON_P1_TET_L: return 1;
  default: return 0;
  }
}

int chunked_vector_cells(int kernel_id)
{
  // cells per chunk of the chunk-aggregated vector kernel of a linear-form kernel, 0 = none
  switch (kernel_id)
  {
  case BFX_K_SOURCE_P1_TRI_L:
  case BFX_K_LOAD_COEFF_P1_TRI_L:
  case BFX_K_LOAD_P1_TET_L:
  case BFX_K_LOAD_P2_TET_L:
  case BFX_K_ACTION_POISSON_P1_TET_L:
  case BFX_K_ACTION_POISSON_P2_TET_L: return 384;
  case BFX_K_LOAD_Q1_HEX_L: return 256;
  default: return 0;
  }
}

int launch_vector_grouped(const bfx_asm* P, int kernel_id, const AsmArgs& a, cudaStream_t st)
{
  if (P->chunks && P->chunks->vector_plan)
  {
    switch (kernel_id)
    {
    case BFX_K_SOURCE_P1_TRI_L: return launch_vector_chunked_e<el::SourceP1Tri>(P, a, st);
    case BFX_K_LOAD_COEFF_P1_TRI_L: return launch_vector_chunked_e<el::LoadCoeffP1Tri>(P, a, st);
    case BFX_K_LOAD_P1_TET_L: return launch_vector_chunked_e<el::LoadP1Tet>(P, a, st);
    case BFX_K_LOAD_P2_TET_L: return launch_vector_chunked_e<el::LoadP2Tet>(P, a, st);
    case BFX_K_LOAD_Q1_HEX_L: return launch_vector_chunked_e<el::LoadQ1Hex>(P, a, st);
    case BFX_K_ACTION_POISSON_P1_TET_L: return launch_vector_chunked_e<el::ActionOf<el::PoissonP1Tet>>(P, a, st);
    case BFX_K_ACTION_POISSON_P2_TET_L: return launch_vector_chunked_e<el::ActionOf<el::PoissonP2Tet>>(P, a, st);
    default: return fail(BFX_ERR_UNSUPPORTED, "kernel id %d has no chunk-aggregated vector variant", kernel_id);
    }
  }
  if (!P->chunks || !P->chunks->glist)
    return fail(BFX_ERR_INVALID, "grouped vector assembly needs bfx_asm_build_groups() on the plan first");
  switch (kernel_id)
  {
  case BFX_K_SOURCE_P1_TRI_L: return launch_vector_grouped_e<el::SourceP1Tri>(P, a, st);
  case BFX_K_LOAD_COEFF_P1_TRI_L: return launch_vector_grouped_e<el::LoadCoeffP1Tri>(P, a, st);
  case BFX_K_LOAD_P1_TET_L: return launch_vector_grouped_e<el::LoadP1Tet>(P, a, st);
  case BFX_K_ACTION_POISSON_P1_TET_L: return launch_vector_grouped_e<el::ActionOf<el::PoissonP1Tet>>(P, a, st);
  default: return fail(BFX_ERR_UNSUPPORTED, "kernel id %d has no grouped vector variant", kernel_id);
  }
}

void free_chunks(bfx_chunks* c)
{
  if (!c)
    return;
  cudaFree(c->part_list);
  cudaFree(c->part_flag);
  cudaFree(c->glist);
  cudaFree(c->goff);
  cudaFree(c->hdr);
  cudaFree(c->winfo);
  cudaFree(c->dest_addr);
  cudaFree(c->wr_addr);
  cudaFree(c->wr_src);
  cudaFree(c->src);
  cudaFree(c->perm);
  cudaFree(c->xdm);
  cudaFree(c->dm0);
  cudaFree(c->dm1);
  cudaFree(c->bits0);
  cudaFree(c->bits1);
  cudaFree(c->colour);
  cudaFree(c->wv_ids);
  cudaFree(c->wv_cnt);
  cudaFree(c->wv_loc);
  cudaFree(c->wd_ids);
  cudaFree(c->wd_cnt);
  cudaFree(c->wd_loc);
  delete c;
}
} // namespace bfx

extern "C"
{
int bfx_asm_build_chunks(bfx_asm_t* P, const double* x_dev, int flags, bfx_stream_t stream)
{
  BFX_REQUIRE(P && P->csr, "bfx_asm_build_chunks: plan has no matrix");
  if (P->ncells == 0) // an empty integration domain (fem/assemble_matrix_impl.h:127): nothing to aggregate
    return fail(BFX_ERR_UNSUPPORTED, "chunk plan of an empty cell list");
  BFX_REQUIRE(P->pos, "bfx_asm_build_chunks: plan has no position map");
  cudaStream_t st = S(stream);
  const bfx_csr* csr = P->csr;
  const bool sym = (flags & BFX_CHUNKS_SYMMETRIC) != 0;
  if (sym && !(csr->bs0 == 1 && csr->bs1 == 1 && P->nd0 == P->nd1 && (!P->dofmap1 || P->dofmap1 == P->dofmap0)))
    return fail(BFX_ERR_UNSUPPORTED, "symmetric chunk plan needs block size 1 and one dofmap for rows and columns");
  const int n2 = sym ? staged_per_cell(P->nd0, true) : P->nd0 * csr->bs0 * P->nd1 * csr->bs1; // staged per cell
  const int cb_req = ((flags >> 8) & 0xff) * 32; // BFX_CHUNKS_CB(cells); 0 = the element's default
  int cb = (cb_req && (chunk_cb_supported(n2, cb_req) || (flags & BFX_CHUNKS_VECTOR))) ? cb_req : chunk_cb(n2); // not instantiated: default
  if (cb == 0 || P->ncells == 0)
    return fail(BFX_ERR_UNSUPPORTED, "chunk plan: element matrices of %d staged scalars are not supported", n2);
  int items = (n2 * cb + PLAN_THREADS - 1) / PLAN_THREADS;
  free_chunks(P->chunks);
  P->chunks = nullptr;
  bfx_chunks* c = new bfx_chunks();
  c->cb = cb;
  c->n2 = n2;
  c->sym = sym;
  c->nchunks = (P->ncells + cb - 1) / cb;
  int e = BFX_OK;
  auto bail = [&](int status)
  {
    free_chunks(c);
    return status;
  };

  // ---- locality ordering of the cell list
  if (x_dev && (e = morton_perm(P, x_dev, &c->perm, st)))
    return bail(e);

  // ---- index arrays in chunk order (phase 1 streams them)
  if (c->perm || P->cells)
  {
    const int64_t n = P->ncells;
    if ((e = dev_alloc(&c->xdm, (size_t)n * P->nx)) || (e = dev_alloc(&c->dm0, (size_t)n * P->nd0)))
      return bail(e);
    k_permute_rows<<<grid_for(n * P->nx, 256, 16), 256, 0, st>>>(n, c->perm, P->cells, P->x_dofmap, P->nx, c->xdm);
    k_permute_rows<<<grid_for(n * P->nd0, 256, 16), 256, 0, st>>>(n, c->perm, P->cells, P->dofmap0, P->nd0, c->dm0);
    if (P->dofmap1 && P->dofmap1 != P->dofmap0)
    {
      if ((e = dev_alloc(&c->dm1, (size_t)n * P->nd1)))
        return bail(e);
      k_permute_rows<<<grid_for(n * P->nd1, 256, 16), 256, 0, st>>>(n, c->perm, P->cells, P->dofmap1, P->nd1, c->dm1);
    }
    BFX_CHECK_LAUNCH();
  }

  // ---- warp tables of geometry nodes and dofs (groups of 32 consecutive cell slots)
  {
    // (padded to a common multiple of the chunk sizes: the plan may fall back to the element's default chunk size below)
    const int64_t nslots_pad = (P->ncells + 767) / 768 * 768, nw = nslots_pad / 32;
    const int32_t* xrows = c->xdm ? c->xdm : P->x_dofmap;
    const int32_t* drows = c->dm0 ? c->dm0 : P->dofmap0;
    const int lv = 4 * ((P->nx + 3) / 4), ld = 4 * ((P->nd0 + 3) / 4);
    const unsigned grid = grid_for(nw, 4, 16);
    if (P->nx <= 8)
    {
      if ((e = dev_alloc(&c->wv_ids, (size_t)nw * 32)) || (e = dev_alloc(&c->wv_cnt, (size_t)nw))
          || (e = dev_alloc(&c->wv_loc, (size_t)nslots_pad * lv)))
        return bail(e);
      k_warp_tables<8><<<grid, 128, 0, st>>>(nslots_pad, P->ncells, P->nx, xrows, c->wv_ids, c->wv_cnt, c->wv_loc);
    }
    if (P->nd0 <= 4 && csr->bs0 == 1) // more dofs per cell never fit 32 distinct per group
    {
      if ((e = dev_alloc(&c->wd_ids, (size_t)nw * 32)) || (e = dev_alloc(&c->wd_cnt, (size_t)nw))
          || (e = dev_alloc(&c->wd_loc, (size_t)nslots_pad * ld)))
        return bail(e);
      k_warp_tables<8><<<grid, 128, 0, st>>>(nslots_pad, P->ncells, P->nd0, drows, c->wd_ids, c->wd_cnt, c->wd_loc);
    }
    BFX_CHECK_LAUNCH();
    // groups that hold cells but have no table (more than 32 distinct ids): the lean kernel needs none
    {
      unsigned long long* d_bad = nullptr;
      unsigned long long h_bad = 0;
      if ((e = dev_alloc(&d_bad, 1)))
        return bail(e);
      BFX_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), st));
      const int64_t nw_active = (P->ncells + 31) / 32;
      if (c->wv_cnt)
        k_count_zero_bytes<<<grid_for(nw_active, 256, 8), 256, 0, st>>>(nw_active, c->wv_cnt, d_bad);
      if (c->wd_cnt)
        k_count_zero_bytes<<<grid_for(nw_active, 256, 8), 256, 0, st>>>(nw_active, c->wd_cnt, d_bad);
      BFX_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(h_bad), cudaMemcpyDeviceToHost, st));
      BFX_CUDA(cudaStreamSynchronize(st));
      cudaFree(d_bad);
      c->tables_complete = c->wv_cnt && h_bad == 0;
    }
  }

  // the lean kernel's plan options need complete warp tables, symmetric pairs and 32-bit value indices: without
  // them the plan is built for the classic kernel (bank-coloured staging, destinations sorted by completeness)
  if (flags & BFX_CHUNKS_VECTOR)
    flags &= ~BFX_CHUNKS_LEN_SORT; // (the vector kernel runs on the padded linear layout with any tables)
  else if (!(c->tables_complete && sym && (uint64_t)csr->nnz * csr->bs0 * csr->bs1 < 0xffffffffull && P->nx <= 4 && P->nd0 <= 4
             && c->wd_ids))
  {
    if (flags & BFX_CHUNKS_LEAN_ONLY)
      return bail(fail(BFX_ERR_UNSUPPORTED, "chunk plan: the lean options do not apply to this mesh (incomplete warp tables)"));
    // ... and with the element's default chunk size when the caller asked for the lean kernel's 384 cells: whole-cube
    // bricks only pay on meshes the Morton order cuts cleanly, which is what complete tables indicate (classic kernel
    // at C2: 3.28 ms with 256 cells, 3.63 ms with 384)
    if ((flags & BFX_CHUNKS_CB_SOFT) && cb != chunk_cb(n2) && chunk_cb(n2) > 0 && 768 % chunk_cb(n2) == 0)
    {
      cb = chunk_cb(n2);
      items = (n2 * cb + PLAN_THREADS - 1) / PLAN_THREADS;
      c->cb = cb;
      c->nchunks = (P->ncells + cb - 1) / cb;
    }
    flags &= ~(BFX_CHUNKS_LINEAR_STAGING | BFX_CHUNKS_BANK_ORDER | BFX_CHUNKS_LEN_SORT);
  }
  c->vector_plan = (flags & BFX_CHUNKS_VECTOR) != 0;
  // ---- scratch of the bit-packed Dirichlet markers (sized by the largest dof the cells reference)
  {
    int32_t* d_max = nullptr;
    int32_t h_max[2] = {0, 0};
    if ((e = dev_alloc(&d_max, 2)))
      return bail(e);
    BFX_CUDA(cudaMemsetAsync(d_max, 0, 2 * sizeof(int32_t), st));
    const int32_t* dm1p = P->dofmap1 ? P->dofmap1 : P->dofmap0;
    k_max_dof<<<grid_for(P->ncells, 256, 8), 256, 0, st>>>(P->ncells, P->cells, P->dofmap0, P->nd0, d_max);
    k_max_dof<<<grid_for(P->ncells, 256, 8), 256, 0, st>>>(P->ncells, P->cells, dm1p, P->nd1, d_max + 1);
    BFX_CUDA(cudaMemcpyAsync(h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_max);
    c->n_dofs0 = ((int64_t)h_max[0] + 1) * csr->bs0;
    c->n_dofs1 = ((int64_t)h_max[1] + 1) * csr->bs1;
    if ((e = dev_alloc(&c->bits0, (size_t)(c->n_dofs0 + 31) / 32 + 1))
        || (e = dev_alloc(&c->bits1, (size_t)(c->n_dofs1 + 31) / 32 + 1)))
      return bail(e);
  }

  // ---- contributions per block entry
  int32_t* total = nullptr;
  if ((e = dev_alloc(&total, (size_t)csr->nnz)))
    return bail(e);
  BFX_CUDA(cudaMemsetAsync(total, 0, sizeof(int32_t) * (size_t)csr->nnz, st));
  if (flags & BFX_CHUNKS_SHARED_MATRIX)
    k_count_contrib_all<<<grid_for(P->ncells_all * P->nd0, 256, 16), 256, 0, st>>>(
        P->ncells_all, P->dofmap0, P->nd0, P->dofmap1 ? P->dofmap1 : P->dofmap0, P->nd1, csr->row_ptr, csr->cols, total);
  else
    k_count_contrib<<<grid_for(P->ncells * P->nd0, 256, 16), 256, 0, st>>>(P->ncells, P->cells, P->dofmap0, P->nd0,
                                                                           P->nd1, csr->row_ptr, P->pos, P->pos_stride,
                                                                           P->pos_bytes, total);
  BFX_CHECK_LAUNCH();

  // ---- pass A: sizes; scans; pass B: write
  const uint64_t nvals = (uint64_t)csr->nnz * csr->bs0 * csr->bs1;
  int addr_bits = 1;
  while (addr_bits < 47 && ((1ull << addr_bits) - 1) < nvals)
    ++addr_bits;
  c->addr_bytes = nvals < 0xffffffffull ? 4 : 8;
  // two-stage write-back: symmetric P1-sized plans of 256-cell chunks with 32-bit addresses whose chunks all keep
  // their destinations within the kernel's shared-memory budget (checked by pass A)
  const bool lean_plan = (flags & BFX_CHUNKS_LINEAR_STAGING) != 0;
  bool two = (flags & BFX_CHUNKS_TWO_STAGE) && sym && c->addr_bytes == 4
             && ((n2 <= 16 && (cb == 256 || lean_plan)) || (two_stage_lists_global(n2) && cb == chunk_cb(n2)));
  int64_t *ndw = nullptr, *nsrc = nullptr;
  if ((e = dev_alloc(&ndw, (size_t)c->nchunks + 1)) || (e = dev_alloc(&nsrc, (size_t)c->nchunks + 1)))
    return bail(e);
  BFX_CUDA(cudaMemsetAsync(ndw, 0, sizeof(int64_t) * (size_t)(c->nchunks + 1), st));
  BFX_CUDA(cudaMemsetAsync(nsrc, 0, sizeof(int64_t) * (size_t)(c->nchunks + 1), st));
  ChunkBuildArgs p;
  memset(&p, 0, sizeof(p));
  p.n = P->ncells;
  p.cb = cb;
  p.nd0 = P->nd0, p.nd1 = P->nd1, p.bs0 = csr->bs0, p.bs1 = csr->bs1, p.ns = n2;
  p.sym = sym;
  p.perm = c->perm;
  p.cells = P->cells;
  p.dofmap0 = P->dofmap0;
  p.row_ptr = csr->row_ptr;
  p.pos = P->pos;
  p.pos_stride = P->pos_stride;
  p.pos_bytes = P->pos_bytes;
  p.total = total;
  p.addr_bits = addr_bits;
  p.o_ndw = ndw;
  p.o_nsrc32 = nsrc;
  p.err = csr->err_flag;
  p.two = two;
  p.pad4 = (flags & BFX_CHUNKS_PAD4) ? 1 : 0;
  p.len_sort = ((flags & BFX_CHUNKS_LEN_SORT) && lean_plan && sym) ? 1 : 0;
  c->len_sorted = p.len_sort != 0;
  p.dcap = lean_plan ? lean_two_dcap(n2, cb) : two_stage_dcap(n2, cb);
  if ((e = run_plan_pass_items(items, false, p, c->nchunks, st)))
    return bail(e);
  {
    void* tmp = nullptr;
    size_t bytes = 0;
    BFX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, ndw, ndw, c->nchunks + 1, st));
    BFX_CUDA(cudaMalloc(&tmp, bytes));
    BFX_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, ndw, ndw, c->nchunks + 1, st));
    BFX_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, nsrc, nsrc, c->nchunks + 1, st));
    int64_t tot[2];
    BFX_CUDA(cudaMemcpyAsync(&tot[0], ndw + c->nchunks, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaMemcpyAsync(&tot[1], nsrc + c->nchunks, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    int herr = 0;
    BFX_CUDA(cudaMemcpyAsync(&herr, csr->err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
    if (herr == 5) // some chunk has more destinations than the two-stage kernel holds: classic write-back
    {
      BFX_CUDA(cudaMemsetAsync(csr->err_flag, 0, sizeof(int), st));
      two = false;
      herr = 0;
    }
    if (herr)
    {
      BFX_CUDA(cudaMemsetAsync(csr->err_flag, 0, sizeof(int), st));
      cudaFree(total);
      cudaFree(ndw);
      cudaFree(nsrc);
      return bail(fail(BFX_ERR_UNSUPPORTED, "chunk plan: a CSR entry receives more than %d contributions from one chunk",
                       MAX_LIST));
    }
    c->n_dest_pad = tot[0] * 32;
    c->n_src32 = tot[1];
  }
  if ((e = dev_alloc(&c->hdr, (size_t)c->nchunks)) || (e = dev_alloc(&c->winfo, (size_t)(c->n_dest_pad / 32 + 8) * (c->len_sorted ? 2 : 1)))
      || (e = dev_alloc(&c->src, (size_t)c->n_src32 * 32 + 32)))
    return bail(e);
  if (two)
  {
    if ((e = dev_alloc(&c->wr_addr, (size_t)(2 * c->n_dest_pad + 64))) || (e = dev_alloc(&c->wr_src, (size_t)(2 * c->n_dest_pad + 64))))
      return bail(e);
  }
  else
  {
    void* da = nullptr;
    BFX_CUDA(cudaMalloc(&da, (size_t)(c->n_dest_pad + 32) * c->addr_bytes * (sym ? 2 : 1)));
    c->dest_addr = da;
  }
  p.two = two ? ((flags & BFX_CHUNKS_TWO_STAGE_SPLIT) ? 2 : 1) : 0;
  p.wr_addr = c->wr_addr;
  p.wr_src = c->wr_src;
  p.dest_base32 = ndw;
  p.src_base32 = nsrc;
  p.hdr = c->hdr;
  p.winfo = c->winfo;
  p.dest_addr = c->dest_addr;
  p.addr_bytes = c->addr_bytes;
  p.src = c->src;
  if ((e = run_plan_pass_items(items, true, p, c->nchunks, st)))
    return bail(e);
  if ((flags & BFX_CHUNKS_LINEAR_STAGING) && (flags & BFX_CHUNKS_BANK_ORDER))
  {
    unsigned long long* d_conf = nullptr;
    if ((e = dev_alloc(&d_conf, 1)))
      return bail(e);
    BFX_CUDA(cudaMemsetAsync(d_conf, 0, sizeof(unsigned long long), st));
    k_chunk_bank_order<<<grid_for(c->nchunks, 8, 16), 256, 0, st>>>(c->nchunks, c->hdr, c->winfo, c->src, n2 * (cb + 1),
                                                                       c->len_sorted ? 2 : 1, d_conf);
    BFX_CHECK_LAUNCH();
    unsigned long long h_conf = 0;
    BFX_CUDA(cudaMemcpyAsync(&h_conf, d_conf, sizeof(h_conf), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_conf);
    c->n_bank_conflicts = (int64_t)h_conf;
    c->bank_ordered = true;
  }
  // ---- bank-conflict-free staging layout (element matrices of at most 16 staged entries)
  if (n2 <= COLOUR_MAX_NS && !(flags & BFX_CHUNKS_LINEAR_STAGING))
  {
    const int colw = 4 * ((n2 + 3) / 4);
    unsigned long long* d_conf = nullptr;
    int* d_over = nullptr;
    if ((e = dev_alloc(&c->colour, (size_t)c->nchunks * cb * colw)) || (e = dev_alloc(&d_conf, 1))
        || (e = dev_alloc(&d_over, 1)))
      return bail(e);
    BFX_CUDA(cudaMemsetAsync(d_conf, 0, sizeof(unsigned long long), st));
    BFX_CUDA(cudaMemsetAsync(d_over, 0, sizeof(int), st));
    const int nst = n2 * cb, max_rg = 2 * src_group_cap(nst);
    const size_t per_warp = ((size_t)nst * 2 + (size_t)max_rg * 2 + (size_t)nst + 15) / 16 * 16;
    const size_t smem = per_warp * 8;
    const unsigned grid = grid_for((c->nchunks + 7) / 8, 1, 8);
    bool launched = false;
    auto colour_cb = [&](auto cbc)
    {
      constexpr int CBC = decltype(cbc)::value;
      if (cb != CBC || launched)
        return cudaSuccess;
      launched = true;
      cudaError_t ce = cudaFuncSetAttribute(k_chunk_colour<CBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (ce != cudaSuccess)
        return ce;
      k_chunk_colour<CBC><<<grid, 256, smem, st>>>(c->nchunks, n2, c->hdr, c->src, c->colour, colw, d_conf, d_over);
      return cudaSuccess;
    };
    BFX_CUDA(colour_cb(std::integral_constant<int, 96>()));
    BFX_CUDA(colour_cb(std::integral_constant<int, 128>()));
    BFX_CUDA(colour_cb(std::integral_constant<int, 192>()));
    BFX_CUDA(colour_cb(std::integral_constant<int, 256>()));
    BFX_CUDA(colour_cb(std::integral_constant<int, 384>()));
    if (!launched)
      return bail(fail(BFX_ERR_INVALID, "chunk plan: no colouring kernel for %d cells per chunk", cb));
    BFX_CHECK_LAUNCH();
    unsigned long long h_conf = 0;
    int h_over = 0;
    BFX_CUDA(cudaMemcpyAsync(&h_conf, d_conf, sizeof(h_conf), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaMemcpyAsync(&h_over, d_over, sizeof(h_over), cudaMemcpyDeviceToHost, st));
    BFX_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_conf);
    cudaFree(d_over);
    c->n_bank_conflicts = (int64_t)h_conf;
    if (h_over)
    {
      // some chunk's lists are far longer than its entries: keep the linear layout (rewrite the lists)
      cudaFree(c->colour);
      c->colour = nullptr;
      c->n_bank_conflicts = -1;
      if ((e = run_plan_pass_items(items, true, p, c->nchunks, st)))
        return bail(e);
    }
  }
  BFX_CUDA(cudaStreamSynchronize(st));
  cudaFree(total);
  cudaFree(ndw);
  cudaFree(nsrc);
  P->chunks = c;
  return BFX_OK;
}

int bfx_asm_build_chunks_vector(bfx_asm_t* P, const double* x_dev, int kernel_id, bfx_stream_t stream)
{
  BFX_REQUIRE(P && !P->csr, "bfx_asm_build_chunks_vector: needs the plan of a linear form");
  const int cb = chunked_vector_cells(kernel_id);
  bfx_kernel_info_t info;
  if (cb == 0 || bfx_kernel_info(kernel_id, &info) != BFX_OK)
    return fail(BFX_ERR_UNSUPPORTED, "kernel id %d has no chunk-aggregated vector variant", kernel_id);
  if (P->ncells == 0)
    return fail(BFX_ERR_UNSUPPORTED, "chunk plan of an empty cell list");
  cudaStream_t st = S(stream);
  // the chunk plan of an identity "matrix": row = dof, one column, position 0 - the value index of (cell, i, a) is
  // bs * dofmap(cell, i) + a, which is exactly where assemble_cells adds be[bs * i + a]
  const int64_t n = (int64_t)P->n_rows_all;
  bfx_csr fake;
  fake.n_rows_all = fake.n_rows_owned = P->n_rows_all;
  fake.bs0 = info.bs;
  fake.bs1 = 1;
  fake.nnz = fake.nnz_owned = n;
  char* zeros = nullptr;
  int e = BFX_OK;
  if ((e = dev_alloc(&fake.row_ptr, (size_t)n + 1)) || (e = dev_alloc(&fake.err_flag, 1)) || (e = dev_alloc(&zeros, 64)))
  {
    cudaFree(fake.row_ptr);
    cudaFree(fake.err_flag);
    return e;
  }
  k_iota64<<<grid_for(n + 1, 256, 16), 256, 0, st>>>(n + 1, fake.row_ptr);
  BFX_CUDA(cudaMemsetAsync(fake.err_flag, 0, sizeof(int), st));
  BFX_CUDA(cudaMemsetAsync(zeros, 0, 64, st));
  const int nd1_saved = P->nd1;
  P->csr = &fake;
  P->pos = zeros;
  P->pos_stride = 0;
  P->pos_bytes = 1;
  P->nd1 = 1;
  e = bfx_asm_build_chunks(P, x_dev, BFX_CHUNKS_VECTOR | BFX_CHUNKS_LINEAR_STAGING | BFX_CHUNKS_BANK_ORDER | BFX_CHUNKS_CB(cb), stream);
  P->csr = nullptr;
  P->pos = nullptr;
  P->nd1 = nd1_saved;
  cudaStreamSynchronize(st);
  cudaFree(fake.row_ptr);
  cudaFree(fake.err_flag);
  cudaFree(zeros);
  if (e == BFX_OK && info.nx <= 4 && info.nd <= 4 && info.bs == 1 && !getenv("BFX_VECTOR_NO_TABLES"))
  {
    // P1-sized kernels run the TABLE kernel (k_vector_tables): it needs the warp tables only.  Without complete tables
    // (meshes the Morton order does not cut into whole groups) the chunk-aggregated kernel is slower than the RED
    // kernel for these elements: tell the caller to keep that one.
    bfx_chunks* c = P->chunks;
    if (!(c->tables_complete && c->wv_ids && c->wd_ids))
    {
      free_chunks(P->chunks);
      P->chunks = nullptr;
      return fail(BFX_ERR_UNSUPPORTED, "vector chunk plan: incomplete warp tables for a P1-sized kernel");
    }
    cudaFree(c->src), c->src = nullptr;
    cudaFree(c->dest_addr), c->dest_addr = nullptr;
    cudaFree(c->winfo), c->winfo = nullptr;
    cudaFree(c->perm), c->perm = nullptr;
    cudaFree(c->xdm), c->xdm = nullptr;
    cudaFree(c->dm0), c->dm0 = nullptr;
    cudaFree(c->wv_cnt), c->wv_cnt = nullptr;
    cudaFree(c->wd_cnt), c->wd_cnt = nullptr;
    c->slim = true;
  }
  return e;
}

int bfx_asm_build_groups(bfx_asm_t* P, const double* x_dev, bfx_stream_t stream)
{
  BFX_REQUIRE(P, "bfx_asm_build_groups: null plan");
  if (P->nx > 4 || P->nd0 > 4 || P->ncells == 0)
    return fail(BFX_ERR_UNSUPPORTED, "grouped vector plan: implemented for cells of at most 4 nodes / 4 dofs");
  cudaStream_t st = S(stream);
  free_chunks(P->chunks);
  P->chunks = nullptr;
  bfx_chunks* c = new bfx_chunks();
  c->cb = 32;
  c->nchunks = (P->ncells + 31) / 32;
  int e = BFX_OK;
  auto bail = [&](int status)
  {
    free_chunks(c);
    return status;
  };
  if (x_dev && (e = morton_perm(P, x_dev, &c->perm, st)))
    return bail(e);
  const int64_t n = P->ncells, nslots_pad = c->nchunks * 32;
  if (c->perm || P->cells)
  {
    if ((e = dev_alloc(&c->xdm, (size_t)n * P->nx)) || (e = dev_alloc(&c->dm0, (size_t)n * P->nd0)))
      return bail(e);
    k_permute_rows<<<grid_for(n * P->nx, 256, 16), 256, 0, st>>>(n, c->perm, P->cells, P->x_dofmap, P->nx, c->xdm);
    k_permute_rows<<<grid_for(n * P->nd0, 256, 16), 256, 0, st>>>(n, c->perm, P->cells, P->dofmap0, P->nd0, c->dm0);
  }
  const int32_t* xrows = c->xdm ? c->xdm : P->x_dofmap;
  const int32_t* drows = c->dm0 ? c->dm0 : P->dofmap0;
  const int lv = 4 * ((P->nx + 3) / 4), ld = 4 * ((P->nd0 + 3) / 4);
  c->gstride = 32 * P->nd0;
  if ((e = dev_alloc(&c->wv_ids, (size_t)nslots_pad)) || (e = dev_alloc(&c->wv_cnt, (size_t)c->nchunks))
      || (e = dev_alloc(&c->wv_loc, (size_t)nslots_pad * lv)) || (e = dev_alloc(&c->wd_ids, (size_t)nslots_pad))
      || (e = dev_alloc(&c->wd_cnt, (size_t)c->nchunks)) || (e = dev_alloc(&c->wd_loc, (size_t)nslots_pad * ld))
      || (e = dev_alloc(&c->glist, (size_t)c->nchunks * c->gstride)) || (e = dev_alloc(&c->goff, (size_t)c->nchunks * 36)))
    return bail(e);
  const unsigned grid = grid_for(c->nchunks, 4, 16);
  k_warp_tables<8><<<grid, 128, 0, st>>>(nslots_pad, n, P->nx, xrows, c->wv_ids, c->wv_cnt, c->wv_loc);
  k_warp_tables<8><<<grid, 128, 0, st>>>(nslots_pad, n, P->nd0, drows, c->wd_ids, c->wd_cnt, c->wd_loc);
  k_group_lists<<<grid, 128, 0, st>>>(c->nchunks, P->nd0, c->wd_cnt, c->wd_loc, n, c->glist, c->goff);
  BFX_CHECK_LAUNCH();
  BFX_CUDA(cudaStreamSynchronize(st));
  P->chunks = c;
  return BFX_OK;
}

int bfx_asm_chunk_stats(const bfx_asm_t* P, int64_t* nchunks, int64_t* n_dest, int64_t* n_src_entries,
                        int64_t* plan_bytes)
{
  // (plan_bytes is followed by the diagnostic bfx_asm_chunk_bank_conflicts)
  BFX_REQUIRE(P && P->chunks, "bfx_asm_chunk_stats: no chunk plan");
  const bfx_chunks* c = P->chunks;
  if (nchunks)
    *nchunks = c->nchunks;
  if (n_dest)
    *n_dest = c->n_dest_pad;
  if (n_src_entries)
    *n_src_entries = c->n_src32 * 32;
  if (plan_bytes)
  {
    // device bytes the plan holds now: headers, list headers, destination addresses, source lists, the locality
    // permutation and chunk-ordered dofmaps (while kept), the warp tables (ids + positions per cell slot)
    const int64_t slots = c->nchunks * c->cb;
    const int64_t lv = 4 * ((P->nx + 3) / 4), ld = 4 * ((P->nd0 + 3) / 4);
    *plan_bytes = c->nchunks * (int64_t)sizeof(ChunkHdr) + c->n_dest_pad / 32 * 4 * (c->len_sorted ? 2 : 1)
                  + (c->wr_addr ? c->n_dest_pad * 12 : c->n_dest_pad * c->addr_bytes * (c->sym ? 2 : 1))
                  + c->n_src32 * 64 + (c->perm ? P->ncells * 4 : 0)
                  + (c->xdm ? P->ncells * 4 * (int64_t)P->nx : 0) + (c->dm0 ? P->ncells * 4 * (int64_t)P->nd0 : 0)
                  + (c->dm1 ? P->ncells * 4 * (int64_t)P->nd1 : 0) + (c->wv_ids ? slots * (4 + lv) : 0)
                  + (c->wd_ids ? slots * (4 + ld) : 0) + (c->colour ? slots * 4 * ((c->n2 + 3) / 4) : 0);
  }
  return BFX_OK;
}

int bfx_asm_chunk_partition(bfx_asm_t* P, int32_t n_owned_rows, int64_t* n_first)
{
  BFX_REQUIRE(P && P->chunks && n_first, "bfx_asm_chunk_partition: no chunk plan");
  bfx_chunks* c = P->chunks;
  if (c->part_rows == n_owned_rows)
  {
    *n_first = c->n_part1;
    return BFX_OK;
  }
  if (!c->slim)
  {
    // plans of the classic kernels (P2 ...): the chunks stay where they are; part 1 runs through a chunk list, part 2
    // skips the flagged chunks (one more load per CTA: nothing against a 4 ms launch of 128-cell chunks)
    if (c->vector_plan || c->wr_addr || !c->dm0)
      return fail(BFX_ERR_UNSUPPORTED, "chunk partition: needs the chunk-ordered dofmap of a matrix plan");
    cudaFree(c->part_list), c->part_list = nullptr;
    cudaFree(c->part_flag), c->part_flag = nullptr;
    unsigned long long* d_count = nullptr;
    int e0;
    if ((e0 = dev_alloc(&c->part_list, (size_t)c->nchunks + 1)) || (e0 = dev_alloc(&c->part_flag, (size_t)c->nchunks + 1))
        || (e0 = dev_alloc(&d_count, 1)))
      return e0;
    BFX_CUDA(cudaMemset(d_count, 0, sizeof(unsigned long long)));
    k_chunk_partition_dofmap<<<grid_for(c->nchunks * 32, 256, 16), 256>>>(c->nchunks, c->cb, c->dm0, P->nd0, P->ncells,
                                                                       n_owned_rows, c->part_flag, c->part_list, d_count);
    BFX_CHECK_LAUNCH();
    unsigned long long hc = 0;
    BFX_CUDA(cudaMemcpy(&hc, d_count, sizeof(hc), cudaMemcpyDeviceToHost));
    cudaFree(d_count);
    // (the list order decides nothing - every chunk of part 1 is launched - but a sorted list keeps the launch order,
    // and with it the order of the REDs, close to that of the whole plan)
    if (hc > 1)
    {
      void* tmp = nullptr;
      size_t bytes = 0;
      uint32_t* sorted = nullptr;
      if (dev_alloc(&sorted, (size_t)hc) == BFX_OK
          && cub::DeviceRadixSort::SortKeys(nullptr, bytes, c->part_list, sorted, (int)hc) == cudaSuccess
          && cudaMalloc(&tmp, bytes) == cudaSuccess
          && cub::DeviceRadixSort::SortKeys(tmp, bytes, c->part_list, sorted, (int)hc) == cudaSuccess
          && cudaDeviceSynchronize() == cudaSuccess)
      {
        cudaFree(c->part_list);
        c->part_list = sorted;
        sorted = nullptr;
      }
      (void)cudaGetLastError();
      cudaFree(tmp);
      cudaFree(sorted);
    }
    c->n_part1 = (int64_t)hc;
    c->part_rows = n_owned_rows;
    *n_first = c->n_part1;
    return BFX_OK;
  }
  if (!(c->tables_complete && c->wv_ids && c->wd_ids && P->nx <= 4 && P->nd0 <= 4 && !c->vector_plan))
    return fail(BFX_ERR_UNSUPPORTED, "chunk partition needs a plan reduced to the lean kernel's arrays");
  // flags, their exclusive scan, the new index of every chunk
  const int64_t nch = c->nchunks;
  int32_t *flag = nullptr, *before = nullptr, *to = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  ChunkHdr* hdr_new = nullptr;
  uint32_t* tn[4] = {nullptr, nullptr, nullptr, nullptr};
  const int64_t nslots_alloc = (P->ncells + 767) / 768 * 768; // (as allocated by bfx_asm_build_chunks)
  auto release = [&]()
  {
    cudaFree(flag), cudaFree(before), cudaFree(to), cudaFree(tmp);
  };
  int e;
  if ((e = dev_alloc(&flag, (size_t)nch + 1)) || (e = dev_alloc(&before, (size_t)nch + 1)) || (e = dev_alloc(&to, (size_t)nch))
      || (e = dev_alloc(&hdr_new, (size_t)nch)))
  {
    release();
    cudaFree(hdr_new);
    return e;
  }
  for (int k = 0; k < 4; ++k)
    if ((e = dev_alloc(&tn[k], (size_t)nslots_alloc)))
    {
      release();
      cudaFree(hdr_new);
      for (auto* t : tn)
        cudaFree(t);
      return e;
    }
  BFX_CUDA(cudaMemset(flag, 0, sizeof(int32_t) * ((size_t)nch + 1)));
  k_chunk_partition<<<grid_for(nch * 32, 256, 16), 256>>>(nch, c->cb, c->wd_ids, P->ncells, n_owned_rows, flag);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flag, before, nch + 1);
  cudaError_t ce = cudaMalloc(&tmp, tmp_bytes);
  if (ce == cudaSuccess)
    ce = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag, before, nch + 1);
  int32_t h_first = 0;
  if (ce == cudaSuccess)
    ce = cudaMemcpy(&h_first, before + nch, sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess)
  {
    (void)cudaGetLastError();
    release();
    cudaFree(hdr_new);
    for (auto* t : tn)
      cudaFree(t);
    return fail(BFX_ERR_CUDA, "bfx_asm_chunk_partition: %s", cudaGetErrorString(ce));
  }
  k_chunk_new_index<<<grid_for(nch, 256, 8), 256>>>(nch, flag, before, h_first, to);
  // tables keep their padding beyond the last chunk
  uint32_t* told[4] = {reinterpret_cast<uint32_t*>(c->wv_ids), reinterpret_cast<uint32_t*>(c->wv_loc),
                       reinterpret_cast<uint32_t*>(c->wd_ids), reinterpret_cast<uint32_t*>(c->wd_loc)};
  for (int k = 0; k < 4; ++k)
    cudaMemcpy(tn[k], told[k], sizeof(uint32_t) * (size_t)nslots_alloc, cudaMemcpyDeviceToDevice);
  k_chunk_move<<<grid_for(nch * c->cb, 256, 16), 256>>>(nch, c->cb, to, c->hdr, hdr_new, told[0], tn[0], told[1], tn[1],
                                                       told[2], tn[2], told[3], tn[3], P->ncells, c->part_rows >= 0 ? 1 : 0);
  ce = cudaDeviceSynchronize();
  release();
  if (ce != cudaSuccess)
  {
    (void)cudaGetLastError();
    cudaFree(hdr_new);
    for (auto* t : tn)
      cudaFree(t);
    return fail(BFX_ERR_CUDA, "bfx_asm_chunk_partition: %s", cudaGetErrorString(ce));
  }
  cudaFree(c->hdr);
  for (auto* t : told)
    cudaFree(t);
  c->hdr = hdr_new;
  c->wv_ids = reinterpret_cast<int32_t*>(tn[0]), c->wv_loc = reinterpret_cast<uint8_t*>(tn[1]);
  c->wd_ids = reinterpret_cast<int32_t*>(tn[2]), c->wd_loc = reinterpret_cast<uint8_t*>(tn[3]);
  c->n_part1 = h_first;
  c->part_rows = n_owned_rows;
  *n_first = c->n_part1;
  return BFX_OK;
}

int bfx_asm_chunk_set_kernel(bfx_asm_t* P, int variant)
{
  BFX_REQUIRE(P && P->chunks, "bfx_asm_chunk_set_kernel: no chunk plan");
  BFX_REQUIRE(variant == BFX_CHUNK_KERNEL_DEFAULT || variant == BFX_CHUNK_KERNEL_OCC5 || variant == BFX_CHUNK_KERNEL_DIET
                  || variant == BFX_CHUNK_KERNEL_LEAN || variant == BFX_CHUNK_KERNEL_WIDE || (variant >= 11 && variant <= 14),
              "bfx_asm_chunk_set_kernel: unknown variant %d", variant);
  bfx_chunks* c = P->chunks;
  if (c->slim && variant != BFX_CHUNK_KERNEL_LEAN)
    return fail(BFX_ERR_INVALID, "bfx_asm_chunk_set_kernel: the plan was reduced to what the lean kernel reads");
  c->kernel_variant = variant;
  // A plan that will run the lean kernel only keeps what that kernel reads: the Morton permutation and the chunk-ordered
  // copies of the geometry dofmap / dofmaps (36 bytes per P1 cell) served the plan construction and the kernels that
  // gather directly; with complete warp tables nothing at run time touches them.
  if (variant == BFX_CHUNK_KERNEL_LEAN && c->sym && c->addr_bytes == 4 && !c->colour && c->tables_complete && c->wv_ids
      && c->wd_ids && P->nx <= 4 && P->nd0 <= 4 && !c->vector_plan && !getenv("BFX_CHUNKS_KEEP_ALL"))
  {
    cudaFree(c->perm), c->perm = nullptr;
    cudaFree(c->xdm), c->xdm = nullptr;
    cudaFree(c->dm0), c->dm0 = nullptr;
    cudaFree(c->dm1), c->dm1 = nullptr;
    cudaFree(c->wv_cnt), c->wv_cnt = nullptr;
    cudaFree(c->wd_cnt), c->wd_cnt = nullptr;
    c->slim = true;
  }
  if (const char* d = getenv("BFX_LEAN_DBG")) // profiling only (wrong results): 1 phase 1, 2 phase 2, 3 phase 2 without updates
    P->chunks->lean_dbg = atoi(d);
  return BFX_OK;
}

int bfx_asm_chunk_two_stage(const bfx_asm_t* P, int* two_stage)
{
  BFX_REQUIRE(P && P->chunks && two_stage, "bfx_asm_chunk_two_stage: no chunk plan");
  *two_stage = P->chunks->wr_addr ? 1 : 0;
  return BFX_OK;
}

int bfx_asm_chunk_bank_conflicts(const bfx_asm_t* P, int64_t* n_conflicts)
{
  BFX_REQUIRE(P && P->chunks && n_conflicts, "bfx_asm_chunk_bank_conflicts: no chunk plan");
  *n_conflicts = (P->chunks->colour || P->chunks->bank_ordered) ? P->chunks->n_bank_conflicts : -1;
  return BFX_OK;
}
}
