// 128-bit global memory access helpers and table record addressing shared by msm.cu and srs.cu.
#pragma once
#include "g1.cuh"

namespace gm {

template <class T>
__device__ __forceinline__ T load_ro(const T* p) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiple");
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(T) / 16); k++) d[k] = __ldg(s + k);
  return r;
}
template <class T>
__device__ __forceinline__ T load_rw(const T* p) {
  T r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(T) / 16); k++) d[k] = s[k];
  return r;
}
template <class T>
__device__ __forceinline__ void store_rw(T* p, const T& v) {
  uint4* d = reinterpret_cast<uint4*>(p);
  const uint4* s = reinterpret_cast<const uint4*>(&v);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(T) / 16); k++) d[k] = s[k];
}

// Record idx of a base table whose records are rec_q 16-byte quads apart: 6 = packed 96-byte points, 8 = points
// padded to 128 bytes so that a gathered record never straddles two 128-byte lines (DRAM traffic of the gathers
// is counted in whole lines: profiles/r01_summary.md).
__device__ __forceinline__ const Affine* rec_at(const Affine* base, size_t idx, int rec_q) {
  return reinterpret_cast<const Affine*>(reinterpret_cast<const uint4*>(base) + idx * (size_t)rec_q);
}

static constexpr int GEN_RUN = 16;        // consecutive multiples per thread in the point generators (srs.cu, srs_setup.cu)
static constexpr int RED_THREADS = 128;    // CTA size of the XYZZ tree reductions

// CTA-wide sum of one XYZZ per thread (shared-memory tree); result valid in thread 0.
__device__ __forceinline__ XYZZ block_sum_xyzz(XYZZ v, XYZZ* sh) {
  store_rw(sh + threadIdx.x, v);
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      XYZZ a = load_rw(sh + threadIdx.x);
      XYZZ b = load_rw(sh + threadIdx.x + s);
      xyzz_add(a, b);
      store_rw(sh + threadIdx.x, a);
    }
    __syncthreads();
  }
  XYZZ r = load_rw(sh);
  __syncthreads();
  return r;
}

}  // namespace gm
