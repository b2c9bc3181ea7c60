// Ghost exchange on device: common::Scatterer (common/Scatterer.h:46-538), la::Vector
// scatter_fwd / scatter_rev (la/Vector.h:219-379) and MatrixCSR::scatter_rev
// (la/MatrixCSR.h:399-468).  MPI_Ineighbor_alltoallv is replaced by one grouped
// ncclSend/ncclRecv per neighbour on a dedicated communication stream, so the exchange overlaps
// whatever the caller queues on its own stream between begin() and end().
//
// NCCL is loaded with dlopen so that libbfx.so has no link-time dependency on it: inside a torch
// process "libnccl.so.2" resolves to the copy torch already loaded.
#include "csr.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <vector>

using namespace bfx;

namespace
{
struct NcclApi
{
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl()
{
  static NcclApi api;
  static bool tried = false;
  if (!tried)
  {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names)
    {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle)
        break;
    }
    if (api.handle)
    {
#define LOAD(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym))
      LOAD(GetUniqueId, "ncclGetUniqueId");
      LOAD(CommInitRank, "ncclCommInitRank");
      LOAD(CommDestroy, "ncclCommDestroy");
      LOAD(Send, "ncclSend");
      LOAD(Recv, "ncclRecv");
      LOAD(GroupStart, "ncclGroupStart");
      LOAD(GroupEnd, "ncclGroupEnd");
      LOAD(AllReduce, "ncclAllReduce");
      LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart
               && api.GroupEnd && api.AllReduce && api.GetErrorString;
    }
  }
  return api;
}

#define BFX_NCCL(call)                                                                                               \
  do                                                                                                                 \
  {                                                                                                                  \
    ncclResult_t _r = (call);                                                                                        \
    if (_r != ncclSuccess)                                                                                           \
      return ::bfx::fail(BFX_ERR_NCCL, "%s failed: %s", #call, nccl().GetErrorString(_r));                           \
  } while (0)
} // namespace

struct bfx_comm
{
  ncclComm_t comm = nullptr;
  int rank = 0, size = 1;
};

struct bfx_scatter
{
  bfx_comm* comm = nullptr;
  int64_t n_local = 0, n_remote = 0;
  int32_t *local_inds = nullptr, *remote_inds = nullptr; // device
  double *buf_local = nullptr, *buf_remote = nullptr;    // device
  std::vector<int32_t> sizes_local, displs_local, dest, sizes_remote, displs_remote, src;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_packed = nullptr, ev_done = nullptr;
};

struct bfx_csr_scatter
{
  const bfx_csr* csr = nullptr;
  bfx_comm* comm = nullptr;
  int bs2 = 1;
  int64_t n_send_blocks = 0, n_recv_blocks = 0;
  int64_t* pack_src = nullptr;   // device: block position in values of each packed block
  int64_t* unpack_pos = nullptr; // device
  double *send_buf = nullptr, *recv_buf = nullptr;
  std::vector<int64_t> send_disp, recv_disp; // scalars
  std::vector<int32_t> src, dest;
  int64_t ghost_begin = 0, ghost_end = 0; // scalar range of the ghost rows in values
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_packed = nullptr, ev_done = nullptr;
};

namespace
{
__global__ void k_pack(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ in,
                       double* __restrict__ out)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[idx[i]];
}

// op 0: out[idx[i]] = in[i]; op 1: out[idx[i]] += in[i] (an owned slot may be hit by several ranks,
// la/Vector.h:96-114, hence the fp64 RED)
__global__ void k_unpack(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ in,
                         double* __restrict__ out, int op)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
  {
    if (op)
      red_add(out + idx[i], in[i]);
    else
      out[idx[i]] = in[i];
  }
}

__global__ void k_pack_blocks(int64_t n_blocks, int bs2, const int64_t* __restrict__ src_pos,
                              const double* __restrict__ values, double* __restrict__ out)
{
  const int64_t total = n_blocks * bs2;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t k = t / bs2;
    out[t] = values[src_pos[k] * bs2 + (t - k * bs2)];
  }
}

__global__ void k_unpack_blocks_add(int64_t n_blocks, int bs2, const int64_t* __restrict__ pos,
                                    const double* __restrict__ in, double* __restrict__ values)
{
  const int64_t total = n_blocks * bs2;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
  {
    const int64_t k = t / bs2;
    red_add(values + pos[k] * bs2 + (t - k * bs2), in[t]);
  }
}

int exchange(bfx_comm* comm, cudaStream_t cs, const double* sendbuf, const int64_t* sdisp, const int32_t* speers,
             int ns, double* recvbuf, const int64_t* rdisp, const int32_t* rpeers, int nr)
{
  if (ns == 0 && nr == 0)
    return BFX_OK;
  if (!comm || !comm->comm)
    return fail(BFX_ERR_NCCL, "ghost exchange with neighbours requires a communicator");
  NcclApi& n = nccl();
  BFX_NCCL(n.GroupStart());
  for (int i = 0; i < nr; ++i)
    if (rdisp[i + 1] > rdisp[i])
      BFX_NCCL(n.Recv(recvbuf + rdisp[i], (size_t)(rdisp[i + 1] - rdisp[i]), ncclDouble, rpeers[i], comm->comm, cs));
  for (int i = 0; i < ns; ++i)
    if (sdisp[i + 1] > sdisp[i])
      BFX_NCCL(n.Send(sendbuf + sdisp[i], (size_t)(sdisp[i + 1] - sdisp[i]), ncclDouble, speers[i], comm->comm, cs));
  BFX_NCCL(n.GroupEnd());
  return BFX_OK;
}

std::vector<int64_t> to_i64(const std::vector<int32_t>& v) { return std::vector<int64_t>(v.begin(), v.end()); }
} // namespace

extern "C"
{
int bfx_comm_unique_id(char id_out[128])
{
  NcclApi& n = nccl();
  if (!n.ok)
    return fail(BFX_ERR_NCCL, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  BFX_NCCL(n.GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return BFX_OK;
}

int bfx_comm_create(bfx_comm_t** out, const char id_in[128], int rank, int size)
{
  NcclApi& n = nccl();
  if (!n.ok)
    return fail(BFX_ERR_NCCL, "libnccl.so.2 could not be loaded");
  BFX_REQUIRE(out && size >= 1 && rank >= 0 && rank < size, "bfx_comm_create: bad rank/size");
  ncclUniqueId id;
  memcpy(id.internal, id_in, 128);
  bfx_comm* c = new bfx_comm();
  c->rank = rank;
  c->size = size;
  BFX_NCCL(n.CommInitRank(&c->comm, size, id, rank));
  *out = c;
  return BFX_OK;
}

int bfx_comm_destroy(bfx_comm_t* c)
{
  if (!c)
    return BFX_OK;
  if (c->comm)
    nccl().CommDestroy(c->comm);
  delete c;
  return BFX_OK;
}

int bfx_comm_rank(const bfx_comm_t* c, int* rank, int* size)
{
  BFX_REQUIRE(c, "null comm");
  if (rank)
    *rank = c->rank;
  if (size)
    *size = c->size;
  return BFX_OK;
}

int bfx_comm_allreduce(bfx_comm_t* c, double* buf, int64_t n, int op, bfx_stream_t stream)
{
  if (!c || c->size == 1)
    return BFX_OK;
  BFX_NCCL(nccl().AllReduce(buf, buf, (size_t)n, ncclDouble, op == 1 ? ncclMax : ncclSum, c->comm, S(stream)));
  return BFX_OK;
}

int bfx_scatter_create(bfx_scatter_t** out, bfx_comm_t* comm, const int32_t* local_inds, int64_t n_local,
                       const int32_t* remote_inds, int64_t n_remote, const int32_t* sizes_local,
                       const int32_t* displs_local, const int32_t* dest, int n_dest, const int32_t* sizes_remote,
                       const int32_t* displs_remote, const int32_t* src, int n_src)
{
  BFX_REQUIRE(out && n_local >= 0 && n_remote >= 0 && n_dest >= 0 && n_src >= 0, "bfx_scatter_create: bad arguments");
  // (a rank may have neighbours on one side only - the owner of a shared plane ghosts nothing: the arrays of the empty
  // side may be NULL)
  if (n_dest + n_src > 0)
    BFX_REQUIRE(comm && (n_dest == 0 || (sizes_local && displs_local && dest)) && (n_src == 0 || (sizes_remote && displs_remote && src)),
                "bfx_scatter_create: neighbours given without a communicator / sizes");
  bfx_scatter* s = new bfx_scatter();
  s->comm = comm;
  s->n_local = n_local;
  s->n_remote = n_remote;
  int e;
  if ((e = upload(&s->local_inds, local_inds, (size_t)n_local)) || (e = upload(&s->remote_inds, remote_inds, (size_t)n_remote))
      || (e = dev_alloc(&s->buf_local, (size_t)n_local)) || (e = dev_alloc(&s->buf_remote, (size_t)n_remote)))
    return e;
  s->sizes_local.assign(sizes_local, sizes_local + n_dest);
  s->displs_local.assign(displs_local, displs_local + n_dest + (n_dest || displs_local ? 1 : 0));
  s->dest.assign(dest, dest + n_dest);
  s->sizes_remote.assign(sizes_remote, sizes_remote + n_src);
  s->displs_remote.assign(displs_remote, displs_remote + n_src + (n_src || displs_remote ? 1 : 0));
  s->src.assign(src, src + n_src);
  if (s->displs_local.empty())
    s->displs_local.push_back(0);
  if (s->displs_remote.empty())
    s->displs_remote.push_back(0);
  BFX_CUDA(cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking));
  BFX_CUDA(cudaEventCreateWithFlags(&s->ev_packed, cudaEventDisableTiming));
  BFX_CUDA(cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming));
  BFX_CUDA(cudaDeviceSynchronize());
  *out = s;
  return BFX_OK;
}

int bfx_scatter_destroy(bfx_scatter_t* s)
{
  if (!s)
    return BFX_OK;
  cudaFree(s->local_inds);
  cudaFree(s->remote_inds);
  cudaFree(s->buf_local);
  cudaFree(s->buf_remote);
  if (s->comm_stream)
    cudaStreamDestroy(s->comm_stream);
  if (s->ev_packed)
    cudaEventDestroy(s->ev_packed);
  if (s->ev_done)
    cudaEventDestroy(s->ev_done);
  delete s;
  return BFX_OK;
}

int bfx_scatter_fwd_begin(bfx_scatter_t* s, const double* x, bfx_stream_t stream)
{
  BFX_REQUIRE(s, "bfx_scatter_fwd_begin: null plan");
  if (s->dest.empty() && s->src.empty())
    return BFX_OK; // Scatterer.h:256-257: nothing to exchange (x may be an empty array)
  BFX_REQUIRE(x, "bfx_scatter_fwd_begin: null array");
  if (s->dest.empty() && s->src.empty())
    return BFX_OK; // Scatterer.h:256-257
  cudaStream_t st = S(stream);
  if (s->n_local > 0)
  {
    k_pack<<<grid_for(s->n_local, 256, 8), 256, 0, st>>>(s->n_local, s->local_inds, x, s->buf_local);
    BFX_CHECK_LAUNCH();
  }
  BFX_CUDA(cudaEventRecord(s->ev_packed, st));
  BFX_CUDA(cudaStreamWaitEvent(s->comm_stream, s->ev_packed, 0));
  const std::vector<int64_t> sd = to_i64(s->displs_local), rd = to_i64(s->displs_remote);
  int e = exchange(s->comm, s->comm_stream, s->buf_local, sd.data(), s->dest.data(), (int)s->dest.size(), s->buf_remote,
                   rd.data(), s->src.data(), (int)s->src.size());
  if (e)
    return e;
  BFX_CUDA(cudaEventRecord(s->ev_done, s->comm_stream));
  return BFX_OK;
}

int bfx_scatter_fwd_end(bfx_scatter_t* s, double* x, int64_t n_owned, bfx_stream_t stream)
{
  BFX_REQUIRE(s, "bfx_scatter_fwd_end: null plan");
  if (s->dest.empty() && s->src.empty())
    return BFX_OK; // Scatterer.h:256-257: nothing to exchange (x may be an empty array)
  BFX_REQUIRE(x, "bfx_scatter_fwd_end: null array");
  if (s->dest.empty() && s->src.empty())
    return BFX_OK;
  cudaStream_t st = S(stream);
  BFX_CUDA(cudaStreamWaitEvent(st, s->ev_done, 0));
  if (s->n_remote > 0)
  {
    k_unpack<<<grid_for(s->n_remote, 256, 8), 256, 0, st>>>(s->n_remote, s->remote_inds, s->buf_remote, x + n_owned, 0);
    BFX_CHECK_LAUNCH();
  }
  return BFX_OK;
}

int bfx_scatter_rev_begin(bfx_scatter_t* s, const double* x, int64_t n_owned, bfx_stream_t stream)
{
  BFX_REQUIRE(s, "bfx_scatter_rev_begin: null plan");
  if (s->dest.empty() && s->src.empty())
    return BFX_OK; // Scatterer.h:256-257: nothing to exchange (x may be an empty array)
  BFX_REQUIRE(x, "bfx_scatter_rev_begin: null array");
  if (s->dest.empty() && s->src.empty())
    return BFX_OK;
  cudaStream_t st = S(stream);
  if (s->n_remote > 0)
  {
    k_pack<<<grid_for(s->n_remote, 256, 8), 256, 0, st>>>(s->n_remote, s->remote_inds, x + n_owned, s->buf_remote);
    BFX_CHECK_LAUNCH();
  }
  BFX_CUDA(cudaEventRecord(s->ev_packed, st));
  BFX_CUDA(cudaStreamWaitEvent(s->comm_stream, s->ev_packed, 0));
  const std::vector<int64_t> sd = to_i64(s->displs_remote), rd = to_i64(s->displs_local);
  int e = exchange(s->comm, s->comm_stream, s->buf_remote, sd.data(), s->src.data(), (int)s->src.size(), s->buf_local,
                   rd.data(), s->dest.data(), (int)s->dest.size());
  if (e)
    return e;
  BFX_CUDA(cudaEventRecord(s->ev_done, s->comm_stream));
  return BFX_OK;
}

int bfx_scatter_rev_end(bfx_scatter_t* s, double* x, int op, bfx_stream_t stream)
{
  BFX_REQUIRE(s, "bfx_scatter_rev_end: null plan");
  if (s->dest.empty() && s->src.empty())
    return BFX_OK; // Scatterer.h:256-257: nothing to exchange (x may be an empty array)
  BFX_REQUIRE(x, "bfx_scatter_rev_end: null array");
  if (s->dest.empty() && s->src.empty())
    return BFX_OK;
  cudaStream_t st = S(stream);
  BFX_CUDA(cudaStreamWaitEvent(st, s->ev_done, 0));
  if (s->n_local > 0)
  {
    k_unpack<<<grid_for(s->n_local, 256, 8), 256, 0, st>>>(s->n_local, s->local_inds, s->buf_local, x, op);
    BFX_CHECK_LAUNCH();
  }
  return BFX_OK;
}

int bfx_csr_scatter_create(bfx_csr_scatter_t** out, const bfx_csr_t* csr, bfx_comm_t* comm,
                           const int32_t* ghost_row_to_rank, int32_t n_ghost_rows, const int64_t* val_send_disp,
                           const int32_t* src, int n_src, const int64_t* val_recv_disp, const int32_t* dest, int n_dest,
                           const int64_t* unpack_pos)
{
  BFX_REQUIRE(out && csr && n_ghost_rows == csr->n_rows_all - csr->n_rows_owned,
              "bfx_csr_scatter_create: ghost row count does not match the matrix");
  bfx_csr_scatter* p = new bfx_csr_scatter();
  p->csr = csr;
  p->comm = comm;
  p->bs2 = csr->bs0 * csr->bs1;
  p->src.assign(src, src + n_src);
  p->dest.assign(dest, dest + n_dest);
  p->send_disp.assign(val_send_disp, val_send_disp + n_src + 1);
  p->recv_disp.assign(val_recv_disp, val_recv_disp + n_dest + 1);
  p->n_send_blocks = p->send_disp.back() / p->bs2;
  p->n_recv_blocks = p->recv_disp.back() / p->bs2;
  p->ghost_begin = csr->nnz_owned * p->bs2;
  p->ghost_end = csr->nnz * p->bs2;
  // pack order: per neighbour, ghost rows in order (la/MatrixCSR.h:406-420)
  std::vector<int64_t> rp((size_t)n_ghost_rows + 1);
  BFX_CUDA(cudaMemcpy(rp.data(), csr->row_ptr + csr->n_rows_owned, sizeof(int64_t) * rp.size(), cudaMemcpyDeviceToHost));
  std::vector<int64_t> pack((size_t)p->n_send_blocks);
  std::vector<int64_t> insert(n_src);
  for (int i = 0; i < n_src; ++i)
    insert[i] = p->send_disp[i] / p->bs2;
  for (int32_t g = 0; g < n_ghost_rows; ++g)
  {
    const int r = ghost_row_to_rank[g];
    BFX_REQUIRE(r >= 0 && r < n_src, "bfx_csr_scatter_create: ghost_row_to_rank out of range");
    for (int64_t k = rp[g]; k < rp[g + 1]; ++k)
      pack[insert[r]++] = k;
  }
  for (int i = 0; i < n_src; ++i)
    BFX_REQUIRE(insert[i] == p->send_disp[i + 1] / p->bs2, "bfx_csr_scatter_create: val_send_disp inconsistent");
  int e;
  if ((e = upload(&p->pack_src, pack.data(), pack.size()))
      || (e = upload(&p->unpack_pos, unpack_pos, (size_t)p->n_recv_blocks))
      || (e = dev_alloc(&p->send_buf, (size_t)p->send_disp.back()))
      || (e = dev_alloc(&p->recv_buf, (size_t)p->recv_disp.back())))
    return e;
  BFX_CUDA(cudaStreamCreateWithFlags(&p->comm_stream, cudaStreamNonBlocking));
  BFX_CUDA(cudaEventCreateWithFlags(&p->ev_packed, cudaEventDisableTiming));
  BFX_CUDA(cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
  BFX_CUDA(cudaDeviceSynchronize());
  *out = p;
  return BFX_OK;
}

int bfx_csr_scatter_destroy(bfx_csr_scatter_t* p)
{
  if (!p)
    return BFX_OK;
  cudaFree(p->pack_src);
  cudaFree(p->unpack_pos);
  cudaFree(p->send_buf);
  cudaFree(p->recv_buf);
  if (p->comm_stream)
    cudaStreamDestroy(p->comm_stream);
  if (p->ev_packed)
    cudaEventDestroy(p->ev_packed);
  if (p->ev_done)
    cudaEventDestroy(p->ev_done);
  delete p;
  return BFX_OK;
}

int bfx_csr_scatter_rev_begin(bfx_csr_scatter_t* p, const double* values, bfx_stream_t stream)
{
  BFX_REQUIRE(p && values, "bfx_csr_scatter_rev_begin: null argument");
  if (p->src.empty() && p->dest.empty())
    return BFX_OK;
  cudaStream_t st = S(stream);
  if (p->n_send_blocks > 0)
  {
    k_pack_blocks<<<grid_for(p->n_send_blocks * p->bs2, 256, 8), 256, 0, st>>>(p->n_send_blocks, p->bs2, p->pack_src,
                                                                               values, p->send_buf);
    BFX_CHECK_LAUNCH();
  }
  BFX_CUDA(cudaEventRecord(p->ev_packed, st));
  BFX_CUDA(cudaStreamWaitEvent(p->comm_stream, p->ev_packed, 0));
  int e = exchange(p->comm, p->comm_stream, p->send_buf, p->send_disp.data(), p->src.data(), (int)p->src.size(),
                   p->recv_buf, p->recv_disp.data(), p->dest.data(), (int)p->dest.size());
  if (e)
    return e;
  BFX_CUDA(cudaEventRecord(p->ev_done, p->comm_stream));
  return BFX_OK;
}

int bfx_csr_scatter_rev_end(bfx_csr_scatter_t* p, double* values, bfx_stream_t stream)
{
  BFX_REQUIRE(p && values, "bfx_csr_scatter_rev_end: null argument");
  cudaStream_t st = S(stream);
  if (!(p->src.empty() && p->dest.empty()))
  {
    BFX_CUDA(cudaStreamWaitEvent(st, p->ev_done, 0));
    if (p->n_recv_blocks > 0)
    {
      k_unpack_blocks_add<<<grid_for(p->n_recv_blocks * p->bs2, 256, 8), 256, 0, st>>>(p->n_recv_blocks, p->bs2,
                                                                                       p->unpack_pos, p->recv_buf, values);
      BFX_CHECK_LAUNCH();
    }
  }
  // Set ghost row data to zero (la/MatrixCSR.h:465-467)
  if (p->ghost_end > p->ghost_begin)
    BFX_CUDA(cudaMemsetAsync(values + p->ghost_begin, 0, sizeof(double) * (size_t)(p->ghost_end - p->ghost_begin), st));
  return BFX_OK;
}
}
