// Microbenchmarks behind DESIGN.md's choice of the scatter-add mechanism (B200, sm_100a):
//   red     : fp64 RED.E.ADD.F64 to pseudo-random addresses, one per lane
//   bulk    : cp.reduce.async.bulk .add.f64 (SASS UBLKRED) of S-byte segments from shared memory to
//             pseudo-random 16-byte aligned addresses, one op per warp-leader
//   ldsrand : 64-bit LDS at pseudo-random shared-memory addresses (bank-conflict cost)
// usage: redbench  (prints one line per case)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z)
{
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// window: addresses of one CTA stay inside a window of `window` doubles (locality knob)
__global__ void k_red(double* g, uint64_t n, int per_thread, uint64_t window)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t base = window >= n ? 0 : (mix(blockIdx.x) % (n - window));
  for (int k = 0; k < per_thread; ++k)
  {
    const uint64_t a = base + mix(tid * 1315423911ull + k) % window;
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(g + a), "d"(1.0) : "memory");
  }
}

// coalescing: the 32 lanes of a warp form 32 / run groups of `run` consecutive doubles, every group at a
// pseudo-random run-aligned address of the window; op 0 = RED, 1 = plain store
__global__ void k_runs(double* g, uint64_t n, int per_thread, uint64_t window, int run, int op)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t base = window >= n ? 0 : (mix(blockIdx.x) % (n - window));
  const uint64_t grp = tid / run, in = tid % run;
  for (int k = 0; k < per_thread; ++k)
  {
    const uint64_t a = base + (mix(grp * 1315423911ull + k) % (window / run)) * run + in;
    if (op == 0)
      asm volatile("red.global.add.f64 [%0], %1;" ::"l"(g + a), "d"(1.0) : "memory");
    else
      asm volatile("st.global.f64 [%0], %1;" ::"l"(g + a), "d"(1.0) : "memory");
  }
}

__global__ void k_bulk(double* g, uint64_t n, int per_warp, int seg_doubles, uint64_t window)
{
  extern __shared__ __align__(128) double s[];
  for (int i = threadIdx.x; i < seg_doubles * (blockDim.x / 32); i += blockDim.x)
    s[i] = 1.0;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x / 32;
  const uint64_t base = window >= n ? 0 : (mix(blockIdx.x) % (n - window));
  if ((threadIdx.x & 31) == 0)
  {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s + warp * seg_doubles);
    const uint64_t gw = (uint64_t)blockIdx.x * (blockDim.x / 32) + warp;
    for (int k = 0; k < per_warp; ++k)
    {
      uint64_t a = base + mix(gw * 2654435761ull + k) % (window - seg_doubles);
      a &= ~1ull; // 16-byte alignment
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(g + a), "r"(sa),
                   "r"(seg_doubles * 8)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if ((k & 7) == 7)
        asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void k_lds(double* out, int iters, int nslots, int mode)
{
  extern __shared__ __align__(128) double s[];
  for (int i = threadIdx.x; i < nslots; i += blockDim.x)
    s[i] = 1.0;
  __syncthreads();
  double acc = 0.0;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t r = (uint32_t)mix(tid);
  for (int k = 0; k < iters; ++k)
  {
    int a;
    r = r * 1664525u + 1013904223u;
    if (mode == 0)
      a = (int)((r >> 10) & (nslots - 1)); // random (nslots is a power of two)
    else
      a = (threadIdx.x + k * 33) & (nslots - 1); // conflict-free
    acc += s[a];
  }
  if (acc == -1.0)
    out[0] = acc;
}

int main()
{
  const uint64_t n = 1ull << 28; // 2 GiB of doubles
  double* g;
  CK(cudaMalloc(&g, n * 8));
  CK(cudaMemset(g, 0, n * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  const int grid = 148 * 8;
  for (uint64_t window : {n, (uint64_t)1 << 16, (uint64_t)1 << 12})
  {
    const int per_thread = 256;
    k_red<<<grid, 256>>>(g, n, 8, window);
    cudaEventRecord(e0);
    k_red<<<grid, 256>>>(g, n, per_thread, window);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * 256 * per_thread;
    printf("red     window %10llu doubles: %.3f ms, %.2f G red/s, %.3f cyc/lane/SM @1.965GHz\n",
           (unsigned long long)window, ms, ops / ms * 1e-6, ms * 1e-3 * 1.965e9 * 148 / ops);
  }
  for (int op : {0, 1})
    for (uint64_t window : {n, (uint64_t)1 << 16})
      for (int run : {1, 2, 4, 8, 16, 32})
      {
        const int per_thread = 128;
        k_runs<<<grid, 256>>>(g, n, 8, window, run, op);
        cudaEventRecord(e0);
        k_runs<<<grid, 256>>>(g, n, per_thread, window, run, op);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)grid * 256 * per_thread;
        printf("%s window %10llu run %2d: %.3f ms, %.2f G lanes/s, %.3f cyc/lane/SM\n", op == 0 ? "red-runs  " : "store-runs",
               (unsigned long long)window, run, ms, ops / ms * 1e-6, ms * 1e-3 * 1.965e9 * 148 / ops);
      }
  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (uint64_t window : {n, (uint64_t)1 << 16})
    for (int seg : {2, 4, 8, 16, 32, 64, 128, 512})
    {
      const int per_warp = 512;
      const int smem = seg * 8 * 8;
      k_bulk<<<grid, 256, smem>>>(g, n, 8, seg, window);
      cudaEventRecord(e0);
      k_bulk<<<grid, 256, smem>>>(g, n, per_warp, seg, window);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      const double ops = (double)grid * 8 * per_warp;
      printf("bulk    window %10llu seg %4d B: %.3f ms, %.2f G ops/s, %.1f GB/s payload, %.2f cyc/op/SM, %.3f cyc/double/SM\n",
             (unsigned long long)window, seg * 8, ms, ops / ms * 1e-6, ops * seg * 8 / ms * 1e-6,
             ms * 1e-3 * 1.965e9 * 148 / ops, ms * 1e-3 * 1.965e9 * 148 / (ops * seg));
    }
  CK(cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  for (int mode : {0, 1})
  {
    const int iters = 4096;
    k_lds<<<148, 1024, 128 * 1024>>>(g, 16, 16384, mode);
    cudaEventRecord(e0);
    k_lds<<<148, 1024, 128 * 1024>>>(g, iters, 16384, mode);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 148.0 * 1024 * iters;
    printf("lds64   %s: %.3f ms, %.3f cyc/lane/SM\n", mode == 0 ? "random" : "conflict-free", ms,
           ms * 1e-3 * 1.965e9 * 148 / ops);
  }
  return 0;
}
