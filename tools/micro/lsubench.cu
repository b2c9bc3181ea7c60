// Microbenchmark: SM-cycles per warp instruction of the LSU-side operations the assembly kernels are made of
// (B200, 32 warps per SM, 148 x 4 CTAs of 256 threads).  nvcc -O3 -arch=sm_100a lsubench.cu -o lsubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, const int* perm, long long* cyc)
{
  __shared__ double sm[4096];
  __shared__ unsigned short s16[4096];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 4096; i += 256) { sm[i] = i; s16[i] = (unsigned short)((i * 37) & 4095); }
  __syncthreads();
  int idx = perm[tid];            // random slot 0..4095
  double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  double v = tid;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it += 8)
  {
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      if (MODE == 0) { // SHFL.IDX of a double (2 SHFL)
        v = __shfl_sync(0xffffffffu, v, (idx + u) & 31);
      } else if (MODE == 1) { // LDS.64 conflict free
        acc0 += sm[((tid + 32 * u) & 4095)];
      } else if (MODE == 2) { // LDS.64 random
        acc0 += sm[(idx + 97 * u) & 4095];
      } else if (MODE == 3) { // LDS.U16 conflict free + dependent LDS.64 (list walk)
        unsigned short a = s16[(tid + 32 * u + it) & 4095];
        acc0 += sm[a];
      } else if (MODE == 4) { // STS.64 conflict free
        sm[(tid + 256 * u) & 4095] = v + u;
      } else if (MODE == 5) { // STS.64 random
        sm[(idx + 97 * u) & 4095] = v + u;
      } else if (MODE == 6) { // LDS.32 random (one word)
        acc0 += (double)reinterpret_cast<int*>(sm)[(idx * 2 + 194 * u) & 8191];
      } else if (MODE == 7) { // DFMA only
        acc0 = fma(acc0, v, acc1); acc1 = fma(acc1, v, acc2); acc2 = fma(acc2, v, acc3); acc3 = fma(acc3, v, acc0);
      } else if (MODE == 8) { // LDS.128 random (16-byte aligned)
        double2 t = reinterpret_cast<double2*>(sm)[(idx + 97 * u) & 2047];
        acc0 += t.x; acc1 += t.y;
      }
    }
    idx = (idx * 5 + 1) & 4095;
  }
  long long t1 = clock64();
  out[blockIdx.x * 256 + tid] = acc0 + acc1 + acc2 + acc3 + v + sm[tid];
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
  int h_perm[256];
  unsigned s = 12345;
  for (int i = 0; i < 256; ++i) { s = s * 1664525u + 1013904223u; h_perm[i] = (s >> 8) & 4095; }
  int* perm; double* out; long long* cyc;
  const int grid = 148 * 4;
  cudaMalloc(&perm, sizeof(h_perm)); cudaMalloc(&out, grid * 256 * 8); cudaMalloc(&cyc, grid * 8);
  cudaMemcpy(perm, h_perm, sizeof(h_perm), cudaMemcpyHostToDevice);
  const char* names[] = {"SHFL.IDX f64 (2 SHFL)", "LDS.64 conflict-free", "LDS.64 random", "LDS.U16 + dependent LDS.64", "STS.64 conflict-free",
                         "STS.64 random", "LDS.32 random", "4 DFMA", "LDS.128 random"};
  long long h[148 * 4];
#define RUN(M) { k<M><<<grid, 256>>>(out, perm, cyc); k<M><<<grid, 256>>>(out, perm, cyc); cudaDeviceSynchronize(); \
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); double a = 0; for (int i = 0; i < grid; ++i) a += h[i]; a /= grid; \
    /* 32 warps per SM each ITERS ops: SM-cycles per warp-op = cycles / (ITERS * 32 warps) */ \
    printf("%-32s %8.0f cycles/CTA  %.3f SM-cycles per warp-op (32 warps/SM)\n", names[M], a, a / (ITERS * 32.0)); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
