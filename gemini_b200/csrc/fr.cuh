// Internal interface of the Fr kernels (fr.cu) used by the C ABI layer (api.cu).
#pragma once
#include <algorithm>
#include "common.cuh"
#include "fp.cuh"

namespace gm {

// d_out[i] = d_f[2i] + r * d_f[2i+1]   (asynchronous on the lane's stream)
int fr_fold_dev(const Lane& ln, int sm_count, const Fr* d_f, size_t n, const Fr& r, Fr* d_out);
inline int fr_fold_dev(gm_ctx* ctx, const Fr* d_f, size_t n, const Fr& r, Fr* d_out) {
  return fr_fold_dev(lane_of(ctx), ctx->sm_count, d_f, n, r, d_out);
}
int fr_random_dev(gm_ctx* ctx, Fr* d_out, size_t n, uint64_t seed);
// d_out[i] = d_in[n - 1 - i] (in place when d_out == d_in)
int fr_reverse_dev(const Lane& ln, int sm_count, const Fr* d_in, size_t n, Fr* d_out);

// number of CTA partial slots a prover over vectors of these lengths can ever need
size_t sc_max_ctas(size_t nf, size_t ng);
// (a, b) of the current vectors -> d_out[0..1]
int sc_message_dev(const Lane& ln, const Fr* d_f, size_t nf, const Fr* d_g, size_t ng, const Fr& twist, bool use_twist,
                   Fr* d_partials, unsigned int* d_ticket, Fr* d_out);
// fold f by rf, g by rg into the out buffers and compute the message of the folded vectors
int sc_fold_message_dev(const Lane& ln, const Fr* d_f, size_t nf, const Fr* d_g, size_t ng, const Fr& rf, const Fr& rg,
                        Fr* d_f_out, Fr* d_g_out, const Fr& new_twist, bool use_twist, Fr* d_partials,
                        unsigned int* d_ticket, Fr* d_out);

// ---- vector helpers of the time prover (all asynchronous on ctx->stream) ----
int fr_powers_dev(gm_ctx* ctx, const Fr& x, size_t n, Fr* d_out);
int fr_eval_even_odd_dev(gm_ctx* ctx, const Fr* d_f, size_t n, const Fr& x, Fr* d_partials, unsigned int* d_ticket, Fr* d_out);
int fr_tensor_dev(gm_ctx* ctx, const Fr* rho, int k, Fr* d_out);
int fr_hadamard_dev(gm_ctx* ctx, const Fr* d_a, const Fr* d_b, size_t n, Fr* d_out);
int fr_axpy_dev(gm_ctx* ctx, Fr* d_acc, const Fr* d_x, size_t n, const Fr& c);
int fr_spmv_dev(gm_ctx* ctx, const uint32_t* d_rowptr, const uint32_t* d_col, const Fr* d_vals, size_t nrows, const Fr* d_x, Fr* d_y);
size_t fr_div_scratch_elems(size_t n);
int fr_div_linear_dev(gm_ctx* ctx, const Fr* d_f, size_t n, const Fr& a, Fr* d_q, Fr* d_rem, Fr* d_scratch);

}  // namespace gm
