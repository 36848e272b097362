// C ABI of the device-resident Fr vector helpers (include/gemini_b200.h, "Fr vectors of the time prover").
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "fr.cuh"

using namespace gm;

static inline void fr_from_u64(Fr& dst, const uint64_t* src) { memcpy(dst.v, src, 32); }

extern "C" {

int gm_dev_memset(gm_ctx* ctx, void* dev, int byte, size_t bytes) {
  GM_ARG(ctx && (dev || bytes == 0), "NULL argument");
  GM_ENTER(ctx);
  GM_CUDA(cudaMemsetAsync(dev, byte, bytes, ctx->stream));
  return GM_OK;
}

int gm_dev_copy(gm_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes) {
  GM_ARG(ctx && ((dst_dev && src_dev) || bytes == 0), "NULL argument");
  GM_ENTER(ctx);
  GM_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return GM_OK;
}

int gm_fr_reverse_dev(gm_ctx* ctx, const void* in_dev, size_t n, void* out_dev) {
  GM_ARG(ctx && ((in_dev && out_dev) || n == 0), "NULL argument");
  GM_ENTER(ctx);
  return fr_reverse_dev(lane_of(ctx), ctx->sm_count, reinterpret_cast<const Fr*>(in_dev), n, reinterpret_cast<Fr*>(out_dev));
}

int gm_fr_powers_dev(gm_ctx* ctx, const uint64_t x[4], size_t n, void* out_dev) {
  GM_ARG(ctx && x && (out_dev || n == 0), "NULL argument");
  GM_ENTER(ctx);
  Fr xx;
  fr_from_u64(xx, x);
  return fr_powers_dev(ctx, xx, n, reinterpret_cast<Fr*>(out_dev));
}

int gm_fr_eval_dev(gm_ctx* ctx, const void* f_dev, size_t n, const uint64_t x[4], uint64_t out_even_odd[8]) {
  GM_ARG(ctx && x && out_even_odd && (f_dev || n == 0), "NULL argument");
  GM_ENTER(ctx);
  Fr xx;
  fr_from_u64(xx, x);
  const size_t ctas = sc_max_ctas(n, n);
  const bool fresh = ctx->fr_red.cap < ctas * 64 + 128;
  GM_TRY(ctx->fr_red.reserve(ctas * 64 + 128));
  uint8_t* base = ctx->fr_red.as<uint8_t>();
  unsigned int* ticket = reinterpret_cast<unsigned int*>(base);
  Fr* out = reinterpret_cast<Fr*>(base + 64);
  Fr* partials = reinterpret_cast<Fr*>(base + 128);
  if (fresh) GM_CUDA(cudaMemsetAsync(ticket, 0, 64, ctx->stream));
  GM_TRY(fr_eval_even_odd_dev(ctx, reinterpret_cast<const Fr*>(f_dev), n, xx, partials, ticket, out));
  GM_CUDA(cudaMemcpyAsync(ctx->pinned, out, 64, cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out_even_odd, ctx->pinned, 64);
  return GM_OK;
}

int gm_fr_tensor_dev(gm_ctx* ctx, const uint64_t* rho, size_t k, void* out_dev) {
  GM_ARG(ctx && out_dev && (rho || k == 0), "NULL argument");
  GM_ARG(k <= 32, "at most 32 tensor factors");
  GM_ENTER(ctx);
  Fr r[32];
  for (size_t j = 0; j < k; j++) fr_from_u64(r[j], rho + 4 * j);
  return fr_tensor_dev(ctx, r, (int)k, reinterpret_cast<Fr*>(out_dev));
}

int gm_fr_hadamard_dev(gm_ctx* ctx, const void* a_dev, const void* b_dev, size_t n, void* out_dev) {
  GM_ARG(ctx && ((a_dev && b_dev && out_dev) || n == 0), "NULL argument");
  GM_ENTER(ctx);
  return fr_hadamard_dev(ctx, reinterpret_cast<const Fr*>(a_dev), reinterpret_cast<const Fr*>(b_dev), n, reinterpret_cast<Fr*>(out_dev));
}

int gm_fr_axpy_dev(gm_ctx* ctx, void* acc_dev, const void* x_dev, size_t n, const uint64_t c[4]) {
  GM_ARG(ctx && c && ((acc_dev && x_dev) || n == 0), "NULL argument");
  GM_ENTER(ctx);
  Fr cc;
  fr_from_u64(cc, c);
  return fr_axpy_dev(ctx, reinterpret_cast<Fr*>(acc_dev), reinterpret_cast<const Fr*>(x_dev), n, cc);
}

int gm_fr_spmv_dev(gm_ctx* ctx, const void* rowptr_dev, const void* col_dev, const void* vals_dev, size_t nrows,
                   const void* x_dev, void* y_dev) {
  GM_ARG(ctx && ((rowptr_dev && y_dev) || nrows == 0), "NULL argument");
  GM_ENTER(ctx);
  return fr_spmv_dev(ctx, reinterpret_cast<const uint32_t*>(rowptr_dev), reinterpret_cast<const uint32_t*>(col_dev),
                     reinterpret_cast<const Fr*>(vals_dev), nrows, reinterpret_cast<const Fr*>(x_dev), reinterpret_cast<Fr*>(y_dev));
}

int gm_fr_div_linear_dev(gm_ctx* ctx, const void* f_dev, size_t n, const uint64_t a[4], void* q_dev, uint64_t out_rem[4]) {
  GM_ARG(ctx && a && out_rem && (f_dev || n == 0) && (q_dev || n <= 1), "NULL argument");
  GM_ENTER(ctx);
  Fr aa;
  fr_from_u64(aa, a);
  GM_TRY(ctx->fr_div.reserve((fr_div_scratch_elems(n) + 1) * 32));
  Fr* rem = ctx->fr_div.as<Fr>();
  GM_TRY(fr_div_linear_dev(ctx, reinterpret_cast<const Fr*>(f_dev), n, aa, reinterpret_cast<Fr*>(q_dev), rem, rem + 1));
  GM_CUDA(cudaMemcpyAsync(ctx->pinned, rem, 32, cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out_rem, ctx->pinned, 32);
  return GM_OK;
}

int gm_fr_fold_chain_dev(gm_ctx* ctx, const void* f_dev, size_t n, const uint64_t* challenges, size_t k, void* out_levels_dev) {
  GM_ARG(ctx && ((f_dev && out_levels_dev) || n == 0 || k == 0) && (challenges || k == 0), "NULL argument");
  GM_ENTER(ctx);
  const Fr* src = reinterpret_cast<const Fr*>(f_dev);
  Fr* dst = reinterpret_cast<Fr*>(out_levels_dev);
  size_t len = n;
  for (size_t j = 0; j < k && len; j++) {
    Fr rr;
    fr_from_u64(rr, challenges + 4 * j);
    GM_TRY(fr_fold_dev(ctx, src, len, rr, dst));
    src = dst;
    len = (len + 1) / 2;
    dst += len;
  }
  return GM_OK;
}

}  // extern "C"
