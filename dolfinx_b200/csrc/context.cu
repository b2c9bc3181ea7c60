// Status, memory plumbing and host helpers of libbfx.so.
#include "common.cuh"
#include <vector>

namespace bfx
{
char* error_buffer()
{
  static thread_local char buf[512] = "";
  return buf;
}

int fail(int status, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return status;
}

int sm_count()
{
  static int n = 0;
  if (n == 0)
  {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess
        || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}
} // namespace bfx

using namespace bfx;

// fem/dofmapbuilder.cpp:446-459 — sequential first-touch scan (fixtures only)
template <typename I>
static int first_touch(const I* dofmap, int64_t n, I ndofs, I* new_index)
{
  for (I i = 0; i < ndofs; ++i)
    new_index[i] = -1;
  I counter = 0;
  for (int64_t k = 0; k < n; ++k)
  {
    I d = dofmap[k];
    if (d < 0 || d >= ndofs)
      return fail(BFX_ERR_INVALID, "first_touch: dof %lld out of range", (long long)d);
    if (new_index[d] == -1)
      new_index[d] = counter++;
  }
  for (I i = 0; i < ndofs; ++i)
    if (new_index[i] == -1)
      new_index[i] = counter++;
  return BFX_OK;
}

extern "C"
{
const char* bfx_last_error(void) { return error_buffer(); }

const char* bfx_status_string(int status)
{
  switch (status)
  {
  case BFX_OK: return "ok";
  case BFX_ERR_CUDA: return "CUDA error";
  case BFX_ERR_INVALID: return "invalid argument";
  case BFX_ERR_NOT_IN_SPARSITY: return "Entry not in sparsity";
  case BFX_ERR_UNSUPPORTED: return "unsupported";
  case BFX_ERR_NCCL: return "NCCL error";
  case BFX_ERR_NO_DEVICE: return "no CUDA device (libbfx has no CPU fallback)";
  default: return "unknown status";
  }
}

int bfx_version(void) { return BFX_VERSION; }

int bfx_device_count(int* count)
{
  *count = 0;
  BFX_CUDA(cudaGetDeviceCount(count));
  return BFX_OK;
}

int bfx_set_device(int device)
{
  BFX_CUDA(cudaSetDevice(device));
  return BFX_OK;
}

int bfx_malloc(void** dev_ptr, size_t bytes)
{
  *dev_ptr = nullptr;
  if (bytes)
    BFX_CUDA(cudaMalloc(dev_ptr, bytes));
  return BFX_OK;
}

int bfx_free(void* dev_ptr)
{
  if (dev_ptr)
    BFX_CUDA(cudaFree(dev_ptr));
  return BFX_OK;
}

int bfx_memcpy(void* dst, const void* src, size_t bytes, bfx_stream_t stream)
{
  if (bytes)
    BFX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, S(stream)));
  return BFX_OK;
}

int bfx_memset(void* dev_ptr, int value, size_t bytes, bfx_stream_t stream)
{
  if (bytes)
    BFX_CUDA(cudaMemsetAsync(dev_ptr, value, bytes, S(stream)));
  return BFX_OK;
}

int bfx_stream_sync(bfx_stream_t stream)
{
  BFX_CUDA(cudaStreamSynchronize(S(stream)));
  return BFX_OK;
}

int bfx_host_alloc(void** host_ptr, size_t bytes)
{
  *host_ptr = nullptr;
  if (bytes)
    BFX_CUDA(cudaMallocHost(host_ptr, bytes));
  return BFX_OK;
}

int bfx_host_free(void* host_ptr)
{
  if (host_ptr)
    BFX_CUDA(cudaFreeHost(host_ptr));
  return BFX_OK;
}

int bfx_host_first_touch_i32(const int32_t* dofmap, int64_t n, int32_t ndofs, int32_t* new_index)
{
  return first_touch<int32_t>(dofmap, n, ndofs, new_index);
}
int bfx_host_first_touch_i64(const int64_t* dofmap, int64_t n, int64_t ndofs, int64_t* new_index)
{
  return first_touch<int64_t>(dofmap, n, ndofs, new_index);
}
}
