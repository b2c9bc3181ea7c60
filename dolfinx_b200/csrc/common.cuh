// Shared host/device helpers for libbfx.so (sm_100a only).
#pragma once
#include "../../include/bfx.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

namespace bfx
{
// Thread-local last-error text; returned by bfx_last_error().
char* error_buffer();
int fail(int status, const char* fmt, ...);

#define BFX_CUDA(call)                                                                                               \
  do                                                                                                                 \
  {                                                                                                                  \
    cudaError_t _e = (call);                                                                                         \
    if (_e != cudaSuccess)                                                                                           \
      (void)cudaGetLastError(); /* (a failed allocation must not surface again at the next launch check) */          \
    if (_e != cudaSuccess)                                                                                           \
      return ::bfx::fail(_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver ? BFX_ERR_NO_DEVICE           \
                                                                                       : BFX_ERR_CUDA,               \
                         "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__);                \
  } while (0)

#define BFX_CHECK_LAUNCH() BFX_CUDA(cudaGetLastError())

#define BFX_REQUIRE(cond, ...)                                                                                       \
  do                                                                                                                 \
  {                                                                                                                  \
    if (!(cond))                                                                                                     \
      return ::bfx::fail(BFX_ERR_INVALID, __VA_ARGS__);                                                              \
  } while (0)

inline cudaStream_t S(bfx_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (148 on B200); cached.
int sm_count();

// Device allocation helper that records failures.
template <typename T>
inline int dev_alloc(T** p, size_t n)
{
  *p = nullptr;
  if (n == 0)
    return BFX_OK;
  BFX_CUDA(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  return BFX_OK;
}

template <typename T>
inline int upload(T** p, const T* src_any, size_t n, cudaStream_t st = 0)
{
  int e = dev_alloc(p, n);
  if (e)
    return e;
  if (n)
    BFX_CUDA(cudaMemcpyAsync(*p, src_any, n * sizeof(T), cudaMemcpyDefault, st));
  return BFX_OK;
}

inline unsigned grid_for(int64_t n, int block, int max_waves = 0)
{
  int64_t g = (n + block - 1) / block;
  if (g < 1)
    g = 1;
  if (max_waves > 0)
  {
    int64_t cap = (int64_t)sm_count() * max_waves;
    if (g > cap)
      g = cap;
  }
  if (g > 2147483647LL)
    g = 2147483647LL;
  return (unsigned)g;
}

#ifdef __CUDACC__
// 128-bit streaming loads that do not pollute L1 (index / value streams read exactly once)
__device__ __forceinline__ int4 ldg_stream(const int4* p)
{
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ldg_stream(const double2* p)
{
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void red_add(double* addr, double v)
{
  // fp64 reduction at L2, no return value (SASS: RED.E.ADD.F64)
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
} // namespace bfx
