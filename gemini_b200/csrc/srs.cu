// SRS construction kernels for sm_100a: packing arkworks records, synthetic bases, CommitterKey::new (fixed-base MSM) and
// the precomputed 2^(c*w) tables.  Split from msm.cu so that the two translation units compile in parallel; these are
// key-setup kernels (untimed, like CommitterKey::new in the reference), the MSM pipeline itself lives in msm.cu.
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

// -------------------------------------------------------------------------------------------
// SRS import / synthetic SRS
// -------------------------------------------------------------------------------------------
// records of `stride` bytes (x | y | ... | flag at inf_offset) -> packed 96-byte points, identity = (0,0)
__global__ void k_pack_points(const uint8_t* __restrict__ raw, size_t n, uint32_t stride, int inf_offset, Affine* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* rec = raw + i * stride;
  Affine p;
  uint32_t* dst = reinterpret_cast<uint32_t*>(&p);
  const bool inf = inf_offset >= 0 && rec[inf_offset] != 0;
  if ((stride & 3u) == 0) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(rec);
#pragma unroll
    for (int k = 0; k < 24; k++) dst[k] = inf ? 0u : src[k];
  } else {
    for (int k = 0; k < 24; k++) {
      uint32_t v = rec[4 * k] | (rec[4 * k + 1] << 8) | (rec[4 * k + 2] << 16) | ((uint32_t)rec[4 * k + 3] << 24);
      dst[k] = inf ? 0u : v;
    }
  }
  store_rw(out + i, p);
}

__global__ void k_fill_points(Affine p, size_t n, Affine* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) store_rw(out + i, p);
}

// P_i = [first + i] G.  Each thread owns GEN_RUN consecutive points: one double-and-add for the
// first, one mixed add of G per further point, one shared inversion (Montgomery's trick) to
// return to affine coordinates.
__device__ __forceinline__ Affine generator_affine() {
  // BLS12-381 G1 generator, Montgomery form (standard generator)
  const uint32_t gx[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u,
                           0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
  const uint32_t gy[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u,
                           0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};
  Affine g;
#pragma unroll
  for (int k = 0; k < 12; k++) { g.x.v[k] = gx[k]; g.y.v[k] = gy[k]; }
  return g;
}

__global__ void __launch_bounds__(64)
k_generate_points(size_t n, uint64_t first, Affine* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t i0 = t * GEN_RUN;
  if (i0 >= n) return;
  const Affine g = generator_affine();
  const uint64_t k = first + i0;
  XYZZ p = XYZZ::identity();
  for (int bit = 63; bit >= 0; bit--) {
    xyzz_dbl(p);
    if ((k >> bit) & 1ull) xyzz_madd(p, g);
  }
  // run of points in XYZZ, prefix products of zzz for the batched inversion
  Fq xs[GEN_RUN], ys[GEN_RUN], zzs[GEN_RUN], zzzs[GEN_RUN], pref[GEN_RUN];
  Fq run = Fq::one();
  const int cntv = (int)min((size_t)GEN_RUN, n - i0);
  for (int r = 0; r < cntv; r++) {
    xs[r] = p.x; ys[r] = p.y; zzs[r] = p.zz; zzzs[r] = p.zzz;
    pref[r] = run;
    if (!p.is_identity()) run = run * p.zzz;
    xyzz_madd(p, g);
  }
  Fq inv = fp_inv(run);
  for (int r = cntv - 1; r >= 0; r--) {
    Affine a;
    if (zzs[r].is_zero()) { a.x = Fq::zero(); a.y = Fq::zero(); }
    else {
      Fq izzz = inv * pref[r];       // 1 / zzz_r
      inv = inv * zzzs[r];
      Fq iz = zzs[r] * izzz;         // 1 / z
      a.x = xs[r] * iz.sqr();
      a.y = ys[r] * izzz;
    }
    store_rw(out + i0 + r, a);
  }
}

// table[w][i] = 2^(c*w) * P_i for w < W (level 0 = the points themselves).  One thread per point:
// c doublings per level in XYZZ, then one shared inversion (Montgomery's trick) back to affine.
static constexpr int PRE_MAX_W = 26;
__global__ void __launch_bounds__(64)
k_precompute(const Affine* __restrict__ pts, size_t n, int c, int W, int rec_q, Affine* __restrict__ table) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Affine p = load_ro(pts + i);
  store_rw(const_cast<Affine*>(rec_at(table, i, rec_q)), p);
  if (p.is_identity()) {
    for (int w = 1; w < W; w++) store_rw(const_cast<Affine*>(rec_at(table, (size_t)w * n + i, rec_q)), p);
    return;
  }
  Fq xs[PRE_MAX_W], ys[PRE_MAX_W], zzs[PRE_MAX_W], zzzs[PRE_MAX_W], pref[PRE_MAX_W];
  XYZZ q = xyzz_from_affine(p);
  Fq run = Fq::one();
  for (int w = 1; w < W; w++) {
    for (int d = 0; d < c; d++) xyzz_dbl(q);
    xs[w] = q.x; ys[w] = q.y; zzs[w] = q.zz; zzzs[w] = q.zzz;
    pref[w] = run;
    run = run * q.zzz;
  }
  Fq inv = fp_inv(run);
  for (int w = W - 1; w >= 1; w--) {
    Fq izzz = inv * pref[w];
    inv = inv * zzzs[w];
    Fq iz = zzs[w] * izzz;
    Affine a;
    a.x = xs[w] * iz.sqr();
    a.y = ys[w] * izzz;
    store_rw(const_cast<Affine*>(rec_at(table, (size_t)w * n + i, rec_q)), a);
  }
}

// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
int msm_precompute(gm_ctx* ctx, const Affine* d_points, size_t n, int c, int W, int rec_q, Affine* d_table) {
  if (n == 0) return GM_OK;
  if (W > PRE_MAX_W) { set_error("precompute: too many windows (%d)", W); return GM_ERR_ARG; }
  LAUNCH(ctx, k_precompute, (unsigned)((n + 63) / 64), 64, 0, d_points, n, c, W, rec_q, d_table);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int srs_pack(gm_ctx* ctx, const uint8_t* d_raw, size_t n, size_t stride, long inf_offset, Affine* d_out) {
  if (n == 0) return GM_OK;
  LAUNCH(ctx, k_pack_points, (unsigned)((n + 255) / 256), 256, 0, d_raw, n, (uint32_t)stride, (int)inf_offset, d_out);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int srs_fill(gm_ctx* ctx, const Affine& p, size_t n, Affine* d_out) {
  if (n == 0) return GM_OK;
  LAUNCH(ctx, k_fill_points, (unsigned)((n + 255) / 256), 256, 0, p, n, d_out);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int srs_generate(gm_ctx* ctx, size_t n, uint64_t first, Affine* d_out) {
  if (n == 0) return GM_OK;
  const size_t threads = (n + GEN_RUN - 1) / GEN_RUN;
  LAUNCH(ctx, k_generate_points, (unsigned)((threads + 63) / 64), 64, 0, n, first, d_out);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}
}  // namespace gm
