// Internal interface of the MSM pipeline (msm.cu) used by the C ABI layer (api.cu).
#pragma once
#include "common.cuh"
#include "g1.cuh"

namespace gm {

struct MsmPlan {
  int c;             // window bits
  int W;             // number of windows = ceil(256 / c)
  uint32_t nb;       // buckets per window = 2^(c-1) (signed digits)
  int L;             // buckets per running-sum slice
  uint32_t nchunks;  // nb / L
  bool merged;       // all windows share one bucket set (precomputed 2^(c*w) multiples of the bases)
};
MsmPlan msm_plan(size_t n);
MsmPlan msm_plan_merged(size_t n, int c_forced);

// Where the bases of an MSM live: plain points, or a precomputed table[w][i] = 2^(c*w) P_i of n points
struct MsmBases {
  const Affine* points = nullptr;
  const Affine* table = nullptr;
  size_t n = 0;  // points per table level
  int c = 0, W = 0;
  int rec_q = 6;  // 16-byte quads per table record: 6 = packed 96 B, 8 = padded to one 128-byte line
};

// *d_acc (XYZZ, device) += sum_i scalars[i] * bases[i]; asynchronous on ctx->stream
int msm_accumulate(gm_ctx* ctx, const MsmBases& bases, size_t base_offset, const uint32_t* d_scalars, size_t n, bool bigint, XYZZ* d_acc);
// the same with the scalars in PINNED host memory: copied in `pieces` pieces, each converted to digits as it lands
// (GM_ERR_ARG: the input needs several passes - use the streamed path)
int msm_accumulate_pinned(gm_ctx* ctx, const MsmBases& bases, size_t base_offset, const uint64_t* h_scalars, size_t n, bool bigint, int pieces,
                          XYZZ* d_acc);
// streamed MSM: buckets (msm_plan_buckets(P) XYZZ) and live flags (same count of u32, zero-initialised) persist
// across pushes; every push must use the same plan and the same kind of bases
MsmPlan msm_stream_plan(const MsmBases& bases, size_t chunk_cap);
size_t msm_plan_buckets(const MsmPlan& P);
int msm_stream_push(gm_ctx* ctx, const MsmBases& bases, size_t base_offset, const uint32_t* d_scalars, size_t n, bool bigint,
                    const MsmPlan& P, XYZZ* d_buckets, uint32_t* d_live);
int msm_stream_reduce(gm_ctx* ctx, const MsmPlan& P, const XYZZ* d_buckets, const uint32_t* d_live, XYZZ* d_acc);
int msm_precompute(gm_ctx* ctx, const Affine* d_points, size_t n, int c, int W, int rec_q, Affine* d_table);
// bucket reduction (msm_reduce.cu): *d_acc += sum over the buckets flagged by valid[]; shared by the one-shot and streamed MSM
int msm_reduce(gm_ctx* ctx, const MsmPlan& P, const XYZZ* buckets, const uint32_t* valid, XYZZ* d_acc);
void msm_describe_plan(size_t n, bool with_table, int sm_count, int out[8]);
int msm_acc_reset(gm_ctx* ctx, XYZZ* d_acc);
int msm_acc_add_jacobians(gm_ctx* ctx, const Jacobian* d_in, size_t k, XYZZ* d_acc);
int msm_acc_normalize(gm_ctx* ctx, const XYZZ* d_acc, Jacobian* d_out);
// constant scalar vectors (msm_reduce.cu): host-synchronising test, and the accumulator as one affine base point
int msm_scalars_all_equal(gm_ctx* ctx, const uint32_t* d_scalars, size_t n, size_t stride, uint32_t* d_flag, uint32_t* pinned_flag, bool* out);
int msm_acc_to_affine(gm_ctx* ctx, const XYZZ* d_acc, Affine* d_out);
// *d_acc = sum of k XYZZ points laid out `stride_bytes` apart (the gathered per-rank partials)
int msm_acc_set_sum_xyzz(gm_ctx* ctx, const void* d_in, size_t k, size_t stride_bytes, XYZZ* d_acc);

// comm.cu: NCCL communicator of the context (multi-GPU jobs)
void comm_destroy(gm_ctx* ctx);
int comm_world(const gm_ctx* ctx);
int comm_rank(const gm_ctx* ctx);
int comm_allgather_dev(gm_ctx* ctx, const void* d_send, size_t bytes, void** out_d_all);

int srs_pack(gm_ctx* ctx, const uint8_t* d_raw, size_t n, size_t stride, long inf_offset, Affine* d_out);
int srs_fill(gm_ctx* ctx, const Affine& p, size_t n, Affine* d_out);
int srs_generate(gm_ctx* ctx, size_t n, uint64_t first, Affine* d_out);
int srs_fixed_base(gm_ctx* ctx, const Affine& g, const uint32_t* d_scalars, size_t n, Affine* d_out);

}  // namespace gm
