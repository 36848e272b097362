// CommitterKey::new on the device (/root/reference/src/kzg/time.rs:49-72): fixed-base MSM for powers_of_g.  Its own
// translation unit: these three kernels take minutes to compile and would otherwise serialise the build.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "fr.cuh"
#include "g1.cuh"
#include "msm.cuh"
#include "msm_mem.cuh"

namespace gm {

#define LAUNCH(ctx, kernel, grid, block, shmem, ...)                       \
  do {                                                                     \
    kernel<<<grid, block, shmem, (ctx)->stream>>>(__VA_ARGS__);            \
    (ctx)->launches++;                                                     \
  } while (0)

// ---- CommitterKey::new (/root/reference/src/kzg/time.rs:49-72): powers_of_g[i] = tau^i * g, the fixed-base MSM of
//      ark-ec (FixedBase::get_window_table / FixedBase::msm + normalize_batch) re-shaped for the device:
//        k_fb_bases    base_j = 2^(16 j) g for the 16 windows of 16 bits (one thread, 240 doublings)
//        k_fb_table    T_j[d] = d * base_j, d < 2^16, as affine points (runs of consecutive multiples, one shared
//                      inversion per run) - 16 x 65536 x 96 B = 100 MB, L2 resident
//        k_fb_msm      out_i = sum_j T_j[digit_j(tau^i)]: at most 16 mixed additions per point, then back to affine
//                      with one shared inversion per run of FB_RUN points
static constexpr int FB_BITS = 16;
static constexpr int FB_WINDOWS = 16;      // 16 x 16 = 256 >= 255 scalar bits
static constexpr int FB_RUN = 8;

__global__ void k_fb_bases(Affine g, Affine* __restrict__ bases) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ p = xyzz_from_affine(g);
  for (int j = 0; j < FB_WINDOWS; j++) {
    Affine a;
    if (p.is_identity()) { a.x = Fq::zero(); a.y = Fq::zero(); }
    else {
      const Fq izzz = fp_inv(p.zzz);
      const Fq iz = p.zz * izzz;
      a.x = p.x * iz.sqr();
      a.y = p.y * izzz;
    }
    store_rw(bases + j, a);
    for (int d = 0; d < FB_BITS; d++) xyzz_dbl(p);
  }
}

// T_j[d] = d * base_j; grid.y = window j; each thread owns GEN_RUN consecutive multiples
__global__ void __launch_bounds__(64)
k_fb_table(const Affine* __restrict__ bases, Affine* __restrict__ table) {
  const int j = blockIdx.y;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t d0 = t * GEN_RUN;
  if (d0 >= (1u << FB_BITS)) return;
  const Affine b = load_ro(bases + j);
  Affine* out = table + ((size_t)j << FB_BITS);
  XYZZ p = XYZZ::identity();
  for (int bit = FB_BITS - 1; bit >= 0; bit--) {
    xyzz_dbl(p);
    if ((d0 >> bit) & 1u) xyzz_madd(p, b);
  }
  Fq xs[GEN_RUN], ys[GEN_RUN], zzs[GEN_RUN], zzzs[GEN_RUN], pref[GEN_RUN];
  Fq run = Fq::one();
#pragma unroll 1   // compile time: the unrolled run is 16 inlined mixed additions (minutes of ptxas for setup-only code)
  for (int r = 0; r < GEN_RUN; r++) {
    xs[r] = p.x; ys[r] = p.y; zzs[r] = p.zz; zzzs[r] = p.zzz;
    pref[r] = run;
    if (!p.is_identity()) run = run * p.zzz;
    xyzz_madd(p, b);
  }
  Fq inv = fp_inv(run);
#pragma unroll 1
  for (int r = GEN_RUN - 1; r >= 0; r--) {
    Affine a;
    if (zzs[r].is_zero()) { a.x = Fq::zero(); a.y = Fq::zero(); }
    else {
      const Fq izzz = inv * pref[r];
      inv = inv * zzzs[r];
      const Fq iz = zzs[r] * izzz;
      a.x = xs[r] * iz.sqr();
      a.y = ys[r] * izzz;
    }
    store_rw(out + d0 + r, a);
  }
}

__global__ void __launch_bounds__(64)
k_fb_msm(const Affine* __restrict__ table, const uint32_t* __restrict__ scalars, size_t n, Affine* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t i0 = t * FB_RUN;
  if (i0 >= n) return;
  const int cnt = (int)min((size_t)FB_RUN, n - i0);
  Fq xs[FB_RUN], ys[FB_RUN], zzs[FB_RUN], zzzs[FB_RUN], pref[FB_RUN];
  Fq run = Fq::one();
#pragma unroll 1
  for (int r = 0; r < cnt; r++) {
    Fr s;
    const uint4* sp = reinterpret_cast<const uint4*>(scalars + (i0 + r) * 8);
    const uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
    s.v[0] = lo.x; s.v[1] = lo.y; s.v[2] = lo.z; s.v[3] = lo.w;
    s.v[4] = hi.x; s.v[5] = hi.y; s.v[6] = hi.z; s.v[7] = hi.w;
    s = s.from_mont();
    XYZZ acc = XYZZ::identity();
#pragma unroll 1
    for (int j = 0; j < FB_WINDOWS; j++) {
      const uint32_t d = (s.v[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
      if (d) xyzz_madd(acc, load_ro(table + ((size_t)j << FB_BITS) + d));
    }
    xs[r] = acc.x; ys[r] = acc.y; zzs[r] = acc.zz; zzzs[r] = acc.zzz;
    pref[r] = run;
    if (!acc.is_identity()) run = run * acc.zzz;
  }
  Fq inv = fp_inv(run);
#pragma unroll 1
  for (int r = cnt - 1; r >= 0; r--) {
    Affine a;
    if (zzs[r].is_zero()) { a.x = Fq::zero(); a.y = Fq::zero(); }
    else {
      const Fq izzz = inv * pref[r];
      inv = inv * zzzs[r];
      const Fq iz = zzs[r] * izzz;
      a.x = xs[r] * iz.sqr();
      a.y = ys[r] * izzz;
    }
    store_rw(out + i0 + r, a);
  }
}

// powers_of_g[i] = scalars[i] * g for i < n (scalars: device, Montgomery Fr); table scratch is allocated and freed here
int srs_fixed_base(gm_ctx* ctx, const Affine& g, const uint32_t* d_scalars, size_t n, Affine* d_out) {
  if (n == 0) return GM_OK;
  Affine* d_tab = nullptr;
  const size_t tab_pts = ((size_t)FB_WINDOWS << FB_BITS) + FB_WINDOWS;
  GM_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_tab), tab_pts * sizeof(Affine)));
  Affine* d_bases = d_tab + ((size_t)FB_WINDOWS << FB_BITS);
  LAUNCH(ctx, k_fb_bases, 1, 32, 0, g, d_bases);
  LAUNCH(ctx, k_fb_table, dim3(((1u << FB_BITS) / GEN_RUN + 63) / 64, FB_WINDOWS), 64, 0, d_bases, d_tab);
  const size_t threads = (n + FB_RUN - 1) / FB_RUN;
  LAUNCH(ctx, k_fb_msm, (unsigned)((threads + 63) / 64), 64, 0, d_tab, d_scalars, n, d_out);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_tab);
  if (e != cudaSuccess) { set_error("fixed-base MSM: %s", cudaGetErrorString(e)); return GM_ERR_CUDA; }
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

}  // namespace gm
