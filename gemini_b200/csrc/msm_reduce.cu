// Bucket reduction of the MSM (rows 7-10 of the pipeline described in msm.cu) and the accumulator helpers: running sums over
// slices of buckets (/root/reference/src/kzg/msm/variable_base.rs:150-166), row / column weighting of the slice sums,
// window totals (variable_base.rs:168-175), final sum and normalisation.  A separate translation unit so that it compiles
// in parallel with the sort / accumulation kernels of msm.cu.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "g1.cuh"
#include "msm.cuh"
#include "msm_mem.cuh"

namespace gm {

#define LAUNCH(ctx, kernel, grid, block, shmem, ...)                       \
  do {                                                                     \
    kernel<<<grid, block, shmem, (ctx)->stream>>>(__VA_ARGS__);            \
    (ctx)->launches++;                                                     \
  } while (0)

// 7. running-sum reduction over slices of L consecutive buckets of one window
// (ncu r02: 255 registers, 8 warps / SM, 67 % sm throughput.  Capping the registers at 168 for 12 warps / SM spills
// 1.2 KB per thread and was measured slower: 1.90 vs 1.72 ms for the whole reduction at 2^20.)
__global__ void __launch_bounds__(128)
k_bucket_chunks(const XYZZ* __restrict__ buckets, const uint32_t* __restrict__ counts, uint32_t nb, int L, uint32_t nchunks,
                int W, XYZZ* __restrict__ chunk_s, XYZZ* __restrict__ chunk_w) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = blockIdx.y;
  if (t >= nchunks) return;
  const size_t g0 = (size_t)w * nb + (size_t)t * L;
  XYZZ running = XYZZ::identity(), sum = XYZZ::identity();
  for (int b = L - 1; b >= 0; b--) {
    if (counts[g0 + b]) {
      XYZZ v = load_rw(buckets + g0 + b);
      xyzz_add(running, v);
    }
    xyzz_add(sum, running);
  }
  store_rw(chunk_s + (size_t)w * nchunks + t, running);
  store_rw(chunk_w + (size_t)w * nchunks + t, sum);
}

// 8. The chunk sums S_t (t < T) still carry the weights t*L.  View t = hi*L2 + lo as an H2 x L2 matrix:
//      sum_t t*S_t = sum_lo lo * C_lo + L2 * sum_hi hi * R_hi,   C_lo / R_hi = plain column / row sums,
//    so the weighting needs one short double-and-add per ROW and per COLUMN (H2 + L2 of them) instead
//    of one per chunk.  Warp roles by index: [0,H2) rows of S, [H2,2*H2) rows of W (plain sums of
//    the locally weighted chunk sums), [2*H2, 2*H2+L2) columns of S.
//    One WARP per row / column (lane-strided loads, then a shuffle tree): 2 H2 + L2 warps fit the chip in one wave,
//    where one 128-thread CTA per row needed three (ncu r02: 0.41 ms at 2^20, 8.6 warps per issue waiting at barriers).
__device__ __forceinline__ XYZZ shfl_down_xyzz_r(const XYZZ& v, int delta) {
  XYZZ r;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* dst = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int k = 0; k < 48; k++) dst[k] = __shfl_down_sync(0xffffffffu, src[k], delta);
  return r;
}

__global__ void __launch_bounds__(RED_THREADS)
k_rowcol(const XYZZ* __restrict__ chunk_s, const XYZZ* __restrict__ chunk_w, uint32_t T, uint32_t H2, uint32_t L2,
         XYZZ* __restrict__ row_sum, XYZZ* __restrict__ wrow_sum, XYZZ* __restrict__ col_sum) {
  const int w = blockIdx.y;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t bx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // one warp per row / column
  if (bx >= 2 * H2 + L2) return;
  const XYZZ* src = (bx >= H2 && bx < 2 * H2) ? chunk_w : chunk_s;
  src += (size_t)w * T;
  XYZZ acc = XYZZ::identity();
  if (bx < 2 * H2) {
    const uint32_t row = bx < H2 ? bx : bx - H2;
    for (uint32_t lo = lane; lo < L2; lo += 32u) {
      XYZZ v = load_rw(src + (size_t)row * L2 + lo);
      xyzz_add(acc, v);
    }
  } else {
    const uint32_t col = bx - 2 * H2;
    for (uint32_t hi = lane; hi < H2; hi += 32u) {
      XYZZ v = load_rw(src + (size_t)hi * L2 + col);
      xyzz_add(acc, v);
    }
  }
#pragma unroll 1
  for (int d = 16; d > 0; d >>= 1) {
    XYZZ other = shfl_down_xyzz_r(acc, d);
    if (lane + d < 32) xyzz_add(acc, other);
  }
  if (lane == 0) {
    if (bx < H2) store_rw(row_sum + (size_t)w * H2 + bx, acc);
    else if (bx < 2 * H2) store_rw(wrow_sum + (size_t)w * H2 + (bx - H2), acc);
    else store_rw(col_sum + (size_t)w * L2 + (bx - 2 * H2), acc);
  }
}

__device__ __forceinline__ XYZZ small_scalar_mul(const XYZZ& p, uint32_t k) {
  XYZZ x = XYZZ::identity();
  if (k == 0 || p.is_identity()) return x;
  for (int bit = 31 - __clz(k); bit >= 0; bit--) {
    xyzz_dbl(x);
    if ((k >> bit) & 1u) xyzz_add(x, p);
  }
  return x;
}

// 9a. one thread per column / row: lo * C_lo  or  2^log2(L2) * hi * R_hi; CTA tree-sum -> partials.
//     CTAs [0, ctas_c) handle columns, the rest rows.
__global__ void __launch_bounds__(RED_THREADS)
k_weighted(const XYZZ* __restrict__ row_sum, const XYZZ* __restrict__ col_sum, uint32_t H2, uint32_t L2, int log_l2,
           uint32_t ctas_c, XYZZ* __restrict__ part) {
  extern __shared__ uint4 sh_raw[];
  XYZZ* sh = reinterpret_cast<XYZZ*>(sh_raw);
  const int w = blockIdx.y;
  XYZZ v = XYZZ::identity();
  if (blockIdx.x < ctas_c) {
    const uint32_t lo = blockIdx.x * blockDim.x + threadIdx.x;
    if (lo < L2) v = small_scalar_mul(load_rw(col_sum + (size_t)w * L2 + lo), lo);
  } else {
    const uint32_t hi = (blockIdx.x - ctas_c) * blockDim.x + threadIdx.x;
    if (hi < H2) {
      v = small_scalar_mul(load_rw(row_sum + (size_t)w * H2 + hi), hi);
      for (int d = 0; d < log_l2; d++) xyzz_dbl(v);
    }
  }
  XYZZ tot = block_sum_xyzz(v, sh);
  if (threadIdx.x == 0) store_rw(part + (size_t)w * gridDim.x + blockIdx.x, tot);
}

// 9b. window total = sum W + L * (sum of weighted partials), times 2^(c*w) when windows keep their own
//     bucket sets (variable_base.rs:168-175).  Lower half of the CTA tree-sums the weighted partials,
//     upper half the W row sums, concurrently.
__global__ void __launch_bounds__(2 * RED_THREADS)
k_window_total(const XYZZ* __restrict__ part, uint32_t nparts, const XYZZ* __restrict__ wrow_sum, uint32_t H2, int log_l,
               int c, int merged, XYZZ* __restrict__ win_sum) {
  extern __shared__ uint4 sh_raw[];
  XYZZ* sh = reinterpret_cast<XYZZ*>(sh_raw);
  const int w = blockIdx.x;
  const uint32_t half = blockDim.x >> 1;
  const uint32_t t = threadIdx.x % half;
  const bool upper = threadIdx.x >= half;
  const XYZZ* src = upper ? wrow_sum + (size_t)w * H2 : part + (size_t)w * nparts;
  const uint32_t cnt = upper ? H2 : nparts;
  XYZZ acc = XYZZ::identity();
  for (uint32_t k = t; k < cnt; k += half) {
    XYZZ v = load_rw(src + k);
    xyzz_add(acc, v);
  }
  store_rw(sh + threadIdx.x, acc);
  __syncthreads();
  for (uint32_t s = half >> 1; s > 0; s >>= 1) {
    if (t < s) {
      XYZZ a = load_rw(sh + threadIdx.x);
      XYZZ b2 = load_rw(sh + threadIdx.x + s);
      xyzz_add(a, b2);
      store_rw(sh + threadIdx.x, a);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    XYZZ tot = load_rw(sh);
    for (int d = 0; d < log_l; d++) xyzz_dbl(tot);
    XYZZ wsum = load_rw(sh + half);
    xyzz_add(tot, wsum);
    if (!merged)
      for (int d = 0; d < c * w; d++) xyzz_dbl(tot);
    store_rw(win_sum + w, tot);
  }
}

// 10. acc += sum of windows
__global__ void k_final(const XYZZ* __restrict__ win_sum, int W, XYZZ* __restrict__ acc) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ a = load_rw(acc);
  for (int w = 0; w < W; w++) {
    XYZZ b = load_rw(win_sum + w);
    xyzz_add(a, b);
  }
  store_rw(acc, a);
}

__global__ void k_normalize(const XYZZ* __restrict__ acc, Jacobian* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ a = load_rw(acc);
  Jacobian j = xyzz_to_jacobian_normalized(a);
  store_rw(out, j);
}

__global__ void k_add_jacobians(const Jacobian* __restrict__ in, uint32_t k, XYZZ* __restrict__ acc) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ a = load_rw(acc);
  for (uint32_t i = 0; i < k; i++) {
    Jacobian j = load_rw(in + i);
    XYZZ b = xyzz_from_jacobian(j);
    xyzz_add(a, b);
  }
  store_rw(acc, a);
}

__global__ void k_sum_xyzz(const uint8_t* __restrict__ in, uint32_t k, uint32_t stride, XYZZ* __restrict__ acc) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ a = XYZZ::identity();
  for (uint32_t i = 0; i < k; i++) {
    XYZZ b = load_rw(reinterpret_cast<const XYZZ*>(in + (size_t)i * stride));
    xyzz_add(a, b);
  }
  store_rw(acc, a);
}

// ---- constant scalar vectors -----------------------------------------------------------------------------------
// sum_i s P_i = s * sum_i P_i.  The reference's own example inputs are constant vectors (circuit.rs:349-365 dummy_r1cs: every
// witness entry is the same field element; examples/snark.rs:62-65), for which arkworks' Pippenger - and ours - puts all n
// points of a window into ONE bucket: n * windows additions where n + one scalar multiplication do.  gm_msm_g1_dev asks
// k_scalars_equal first (a strided sample, then - only if the sample is constant - every scalar).
// flag (pre-set to 1) is cleared when some scalar differs from scalars[0]; term i is read at i * stride * 8 words
__global__ void k_scalars_equal(const uint32_t* __restrict__ scalars, size_t n, size_t stride, size_t step, uint32_t* flag) {
  const uint4* p0 = reinterpret_cast<const uint4*>(scalars);
  const uint4 a0 = __ldg(p0), a1 = __ldg(p0 + 1);
  bool same = true;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * step; i < n; i += (size_t)gridDim.x * blockDim.x * step) {
    const uint4* p = reinterpret_cast<const uint4*>(scalars + i * stride * 8);
    const uint4 b0 = __ldg(p), b1 = __ldg(p + 1);
    same = same && a0.x == b0.x && a0.y == b0.y && a0.z == b0.z && a0.w == b0.w && a1.x == b1.x && a1.y == b1.y && a1.z == b1.z && a1.w == b1.w;
  }
  if (!same) *flag = 0u;
}
__global__ void k_acc_to_affine(const XYZZ* __restrict__ acc, Affine* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  const Jacobian j = xyzz_to_jacobian_normalized(load_rw(acc));
  Affine a;
  if (j.z.is_zero()) { a.x = Fq::zero(); a.y = Fq::zero(); }   // (0, 0) = identity
  else { a.x = j.x; a.y = j.y; }
  store_rw(out, a);
}

// host-synchronising: *out = every one of the n scalars (stride apart) equals the first one.  d_flag: one device word.
int msm_scalars_all_equal(gm_ctx* ctx, const uint32_t* d_scalars, size_t n, size_t stride, uint32_t* d_flag, uint32_t* pinned_flag, bool* out) {
  *out = false;
  if (n < 2) { *out = true; return GM_OK; }
  // 1. a sample of <= 4096 scalars spread over the vector: non-constant vectors (every real witness) stop here
  for (int pass = 0; pass < 2; pass++) {
    const size_t step = pass == 0 ? std::max<size_t>(1, n / 4096) : 1;
    if (pass == 1 && n / 4096 <= 1) break;             // the sample already was the whole vector
    const size_t items = (n + step - 1) / step;
    const unsigned grid = (unsigned)std::min<size_t>((items + 255) / 256, (size_t)ctx->sm_count * 8);
    *pinned_flag = 1u;
    GM_CUDA(cudaMemcpyAsync(d_flag, pinned_flag, 4, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_scalars_equal, grid, 256, 0, d_scalars, n, stride, step, d_flag);
    GM_CUDA(cudaMemcpyAsync(pinned_flag, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    GM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*pinned_flag == 0u) return GM_OK;
  }
  *out = true;
  return GM_OK;
}

int msm_acc_to_affine(gm_ctx* ctx, const XYZZ* d_acc, Affine* d_out) {
  LAUNCH(ctx, k_acc_to_affine, 1, 32, 0, d_acc, d_out);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

// Phase B: bucket reduction (running sums, weights, windows) and *d_acc += result.  `valid[gb] != 0` marks the
// buckets that hold a point (per-call counts, or the persistent live flags of a stream).
int msm_reduce(gm_ctx* ctx, const MsmPlan& P, const XYZZ* buckets, const uint32_t* valid, XYZZ* d_acc) {
  MsmScratch& S = ctx->msm;
  const bool merged = P.merged;
  const int Weff = merged ? 1 : P.W;
  // chunk index t = hi * L2 + lo, an H2 x L2 matrix (both powers of two)
  int log_t = 0;
  while ((1u << log_t) < P.nchunks) log_t++;
  const int log_l2 = (log_t + 1) / 2;
  const uint32_t L2 = 1u << log_l2, H2 = P.nchunks >> log_l2;
  int log_l = 0;
  while ((1 << log_l) < P.L) log_l++;
  const uint32_t ctas_c = (L2 + RED_THREADS - 1) / RED_THREADS, ctas_r = (H2 + RED_THREADS - 1) / RED_THREADS;
  const uint32_t nparts = ctas_c + ctas_r;
  // small: chunk_s | chunk_w | row_sum | wrow_sum | col_sum | part | win_sum
  const size_t off_cs = 0;
  const size_t off_cw = off_cs + (size_t)Weff * P.nchunks * sizeof(XYZZ);
  const size_t off_rs = off_cw + (size_t)Weff * P.nchunks * sizeof(XYZZ);
  const size_t off_wr = off_rs + (size_t)Weff * H2 * sizeof(XYZZ);
  const size_t off_col = off_wr + (size_t)Weff * H2 * sizeof(XYZZ);
  const size_t off_bp = off_col + (size_t)Weff * L2 * sizeof(XYZZ);
  const size_t off_ws = off_bp + (size_t)Weff * nparts * sizeof(XYZZ);
  const size_t small_bytes = off_ws + (size_t)Weff * sizeof(XYZZ);
  GM_TRY(S.small.reserve(small_bytes));
  uint8_t* sm = S.small.as<uint8_t>();
  XYZZ* chunk_s = reinterpret_cast<XYZZ*>(sm + off_cs);
  XYZZ* chunk_w = reinterpret_cast<XYZZ*>(sm + off_cw);
  XYZZ* row_sum = reinterpret_cast<XYZZ*>(sm + off_rs);
  XYZZ* wrow_sum = reinterpret_cast<XYZZ*>(sm + off_wr);
  XYZZ* col_sum = reinterpret_cast<XYZZ*>(sm + off_col);
  XYZZ* part = reinterpret_cast<XYZZ*>(sm + off_bp);
  XYZZ* win_sum = reinterpret_cast<XYZZ*>(sm + off_ws);
  const size_t red_sh = RED_THREADS * sizeof(XYZZ);
  LAUNCH(ctx, k_bucket_chunks, dim3((P.nchunks + 127) / 128, Weff), 128, 0, buckets, valid, P.nb, P.L, P.nchunks, Weff, chunk_s, chunk_w);
  LAUNCH(ctx, k_rowcol, dim3((2 * H2 + L2 + RED_THREADS / 32 - 1) / (RED_THREADS / 32), Weff), RED_THREADS, 0, chunk_s, chunk_w, P.nchunks, H2, L2, row_sum, wrow_sum, col_sum);
  LAUNCH(ctx, k_weighted, dim3(nparts, Weff), RED_THREADS, red_sh, row_sum, col_sum, H2, L2, log_l2, ctas_c, part);
  LAUNCH(ctx, k_window_total, Weff, 2 * RED_THREADS, 2 * red_sh, part, nparts, wrow_sum, H2, log_l, P.c, merged ? 1 : 0, win_sum);
  LAUNCH(ctx, k_final, 1, 32, 0, win_sum, Weff, d_acc);
  GM_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int msm_acc_reset(gm_ctx* ctx, XYZZ* d_acc) {
  GM_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(XYZZ), ctx->stream));
  return GM_OK;
}

int msm_acc_add_jacobians(gm_ctx* ctx, const Jacobian* d_in, size_t k, XYZZ* d_acc) {
  LAUNCH(ctx, k_add_jacobians, 1, 32, 0, d_in, (uint32_t)k, d_acc);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int msm_acc_set_sum_xyzz(gm_ctx* ctx, const void* d_in, size_t k, size_t stride_bytes, XYZZ* d_acc) {
  LAUNCH(ctx, k_sum_xyzz, 1, 32, 0, reinterpret_cast<const uint8_t*>(d_in), (uint32_t)k, (uint32_t)stride_bytes, d_acc);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int msm_acc_normalize(gm_ctx* ctx, const XYZZ* d_acc, Jacobian* d_out) {
  LAUNCH(ctx, k_normalize, 1, 32, 0, d_acc, d_out);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}
}  // namespace gm
