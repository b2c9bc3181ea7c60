// Device helpers shared by the assembly kernels (assemble.cu, chunked.cu).
#pragma once
#include "assemble.cuh"

namespace bfx
{
// ---------------------------------------------------------------------------------------------
// index / coordinate / coefficient loads
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t find_col(const int32_t* __restrict__ cols, int64_t b, int64_t e, int32_t c)
{
  int64_t lo = b, hi = e;
  while (lo < hi)
  {
    const int64_t mid = (lo + hi) >> 1;
    if (cols[mid] < c)
      lo = mid + 1;
    else
      hi = mid;
  }
  return (lo < e && cols[lo] == c) ? lo : -1;
}

template <int N>
__device__ __forceinline__ void load_ints(const int32_t* __restrict__ p, int32_t (&out)[N])
{
  if constexpr (N % 4 == 0)
  {
#pragma unroll
    for (int k = 0; k < N / 4; ++k)
    {
      const int4 v = ldg_stream(reinterpret_cast<const int4*>(p) + k);
      out[4 * k] = v.x, out[4 * k + 1] = v.y, out[4 * k + 2] = v.z, out[4 * k + 3] = v.w;
    }
  }
  else if constexpr (N % 2 == 0)
  {
#pragma unroll
    for (int k = 0; k < N / 2; ++k)
    {
      const int2 v = __ldg(reinterpret_cast<const int2*>(p) + k);
      out[2 * k] = v.x, out[2 * k + 1] = v.y;
    }
  }
  else
  {
#pragma unroll
    for (int k = 0; k < N; ++k)
      out[k] = __ldg(p + k);
  }
}

// coordinate_dofs gather: x is (N,3) row-major (assemble_matrix_impl.h:146-148)
template <int NX>
__device__ __forceinline__ void gather_coords(const double* __restrict__ x, const int32_t (&xd)[NX],
                                              double (&xc)[NX][3])
{
#pragma unroll
  for (int i = 0; i < NX; ++i)
  {
    const double* p = x + 3 * (int64_t)xd[i];
    xc[i][0] = __ldg(p);
    xc[i][1] = __ldg(p + 1);
    xc[i][2] = __ldg(p + 2);
  }
}

// Coefficients: packed reference layout, or the fused gather of the element's coefficients (NCOEF of them, one unless
// the element says otherwise, each WND nodes x WBS components: pack_impl, fem/pack.h:77-103), w = [coef 0 | coef 1 ..]
template <class E, class = void>
struct ncoef_of
{
  static constexpr int value = 1;
};
template <class E>
struct ncoef_of<E, decltype((void)E::NCOEF)>
{
  static constexpr int value = E::NCOEF;
};

template <class E>
__device__ __forceinline__ void load_w(const AsmArgs& a, int64_t e, int32_t cell, double* w)
{
  if constexpr (E::WSIZE > 0)
  {
    if (a.coef.packed)
    {
#pragma unroll
      for (int k = 0; k < E::WSIZE; ++k)
        w[k] = __ldg(a.coef.packed + e * a.coef.cstride + a.coef.f[0].off + k);
    }
    else
    {
      constexpr int NC = ncoef_of<E>::value;
      static_assert(NC <= 4 && NC * E::WND * E::WBS == E::WSIZE, "fused gather: NCOEF coefficients of WND x WBS scalars");
#pragma unroll
      for (int c = 0; c < NC; ++c)
      {
        const int32_t* dm = a.coef.f[c].dm + (int64_t)cell * E::WND;
        const double* v = a.coef.f[c].v;
#pragma unroll
        for (int i = 0; i < E::WND; ++i)
        {
          const int64_t d = __ldg(dm + i);
#pragma unroll
          for (int k = 0; k < E::WBS; ++k)
            w[(c * E::WND + i) * E::WBS + k] = __ldg(v + E::WBS * d + k);
        }
      }
    }
  }
}

template <typename PosT, int COUNT>
struct PosRegs
{
  static constexpr int BYTES = COUNT * (int)sizeof(PosT);
  static constexpr int STRIDE = (BYTES + 15) / 16 * 16;
  uint32_t wds[STRIDE / 4];
  __device__ __forceinline__ void load(const void* base, int64_t e)
  {
    const uint4* p = reinterpret_cast<const uint4*>(static_cast<const char*>(base) + e * STRIDE);
#pragma unroll
    for (int k = 0; k < STRIDE / 16; ++k)
    {
      const int4 v = ldg_stream(reinterpret_cast<const int4*>(p + k));
      wds[4 * k] = v.x, wds[4 * k + 1] = v.y, wds[4 * k + 2] = v.z, wds[4 * k + 3] = v.w;
    }
  }
  __device__ __forceinline__ uint32_t get(int t) const
  {
    if constexpr (sizeof(PosT) == 1)
      return (wds[t >> 2] >> ((t & 3) * 8)) & 0xffu;
    else
      return (wds[t >> 1] >> ((t & 1) * 16)) & 0xffffu;
  }
};

} // namespace bfx
