// Internal interface of the Fr kernels (fr.cu) used by the C ABI layer (api.cu).
#pragma once
#include <algorithm>
#include "common.cuh"
#include "fp.cuh"

namespace gm {

// Mailbox of a sumcheck prover: pinned host memory that the device writes directly (UVA).  The last CTA of a round's
// kernel publishes the message (msg, then msg_seq); the host spins on msg_seq instead of copying and synchronising.
struct ScMailbox {
  Fr msg[2];                   // (a, b) of the last published round
  volatile uint32_t msg_seq;   // device -> host: messages published so far
  uint32_t pad0[15];
  Fr scratch[2];               // 64-byte D2H slot of the calls that do copy (final foldings)
  uint32_t pad1[16];           // 256 bytes: mailboxes of different provers never share a cache line
};
static_assert(sizeof(ScMailbox) == 256, "mailbox layout");

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
// (mb, seq): the message is also published in the prover's pinned mailbox (msg, then msg_seq = seq)
int sc_message_dev(const Lane& ln, const Fr* d_f, size_t nf, const Fr* d_g, size_t ng, const Fr& twist, bool use_twist,
                   Fr* d_partials, unsigned int* d_ticket, Fr* d_out, ScMailbox* mb, uint32_t seq);
// fold f by rf, g by rg into the out buffers and compute the message of the folded vectors
int sc_fold_message_dev(const Lane& ln, const Fr* d_f, size_t nf, const Fr* d_g, size_t ng, const Fr& rf, const Fr& rg,
                        Fr* d_f_out, Fr* d_g_out, const Fr& new_twist, bool use_twist, Fr* d_partials,
                        unsigned int* d_ticket, Fr* d_out, ScMailbox* mb, uint32_t seq);

// ---- vector helpers of the time prover (all asynchronous on ctx->stream) ----
int fr_powers_dev(gm_ctx* ctx, const Fr& x, size_t n, Fr* d_out);
int fr_eval_even_odd_dev(gm_ctx* ctx, const Fr* d_f, size_t n, const Fr& x, Fr* d_partials, unsigned int* d_ticket, Fr* d_out);
int fr_tensor_dev(gm_ctx* ctx, const Fr* rho, int k, Fr* d_out);
int fr_hadamard_dev(gm_ctx* ctx, const Fr* d_a, const Fr* d_b, size_t n, Fr* d_out);
int fr_axpy_dev(gm_ctx* ctx, Fr* d_acc, const Fr* d_x, size_t n, const Fr& c);
int fr_spmv_dev(gm_ctx* ctx, const uint32_t* d_rowptr, const uint32_t* d_col, const Fr* d_vals, size_t nrows, const Fr* d_x, Fr* d_y);
size_t fr_div_scratch_elems(size_t n);
int fr_div_linear_dev(gm_ctx* ctx, const Fr* d_f, size_t n, const Fr& a, Fr* d_q, Fr* d_rem, Fr* d_scratch);

// host side of the mailbox: spin until the device has published message `seq` (false: the kernel died or 10 s passed)
bool sc_wait_message(cudaStream_t stream, ScMailbox* mb, uint32_t seq);
static constexpr size_t SC_TAIL_MAX = (size_t)1 << 13;   // vectors of at most this many elements finish in the tail kernel

}  // namespace gm

struct gm_sumcheck {
  gm_ctx* ctx = nullptr;
  cudaStream_t stream = nullptr;        // private: distinct provers run concurrently (proof.rs:85 drives them from rayon)
  cudaEvent_t ev[2] = {nullptr, nullptr};  // per-call device time
  cudaEvent_t tm[2] = {nullptr, nullptr};  // gm_sumcheck_timer_start / _stop
  float last_ms = 0.f;
  int slot = -1;                        // pinned 64-byte message slot of the context (-1: own allocation)
  gm::Fr* f[2] = {nullptr, nullptr};
  gm::Fr* g[2] = {nullptr, nullptr};
  int cur = 0;
  size_t nf = 0, ng = 0;
  gm::Fr twist;
  size_t round = 0, tot_rounds = 0;
  int flavour = 0;
  gm::Fr* d_partials = nullptr;
  unsigned int* d_ticket = nullptr;
  gm::Fr* d_out = nullptr;   // 2 Fr
  gm::Fr* h_out = nullptr;   // pinned, 2 Fr (= mbox->scratch)
  gm::ScMailbox* mbox = nullptr;   // pinned, device-visible (UVA): round messages land here
  uint32_t seq = 0;                // messages published so far (mbox->msg_seq after the last round)
};

