// Shared host-side plumbing of libgemini_b200: context, error reporting, scratch arena.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gemini_b200.h"

namespace gm {

void set_error(const char* fmt, ...);

#define GM_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::gm::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (_e == cudaErrorMemoryAllocation) ? GM_ERR_OOM : GM_ERR_CUDA;                  \
    }                                                                                       \
  } while (0)

#define GM_TRY(expr)        \
  do {                      \
    int _r = (expr);        \
    if (_r != GM_OK) return _r; \
  } while (0)

#define GM_ARG(cond, msg)                   \
  do {                                      \
    if (!(cond)) {                          \
      ::gm::set_error("bad argument: %s", msg); \
      return GM_ERR_ARG;                    \
    }                                       \
  } while (0)

// Grow-only device buffer (cudaMalloc is slow and synchronising; scratch is reused across calls).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return GM_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + (bytes >> 3);
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
      if (e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return GM_ERR_OOM;
      }
    }
    cap = want;
    return GM_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct MsmScratch {
  size_t scalar_stride = 1;   // scalars of the running call are `stride` Fr elements apart (cyclic sharding across GPUs); set per call
  DevBuf scalars;    // staged scalars (host-call path)
  DevBuf digits;     // W*n digit codes
  DevBuf sorted;     // W*n point references grouped by bucket
  DevBuf counts, starts, cursor, poff;  // per global bucket
  DevBuf buckets;    // XYZZ per global bucket
  DevBuf partials;   // XYZZ partial sums of split buckets
  DevBuf work;       // work items
  DevBuf split;      // split-bucket list
  DevBuf small;      // chunk sums, row / column sums, window sums
  DevBuf meta;       // counters and size histogram of the work list
  DevBuf scan_tmp;
  DevBuf bases_tmp;  // ad-hoc bases (hostbases / stream pushes with points)
  // affine levels (msm.cu 3b): x|y planes of the level outputs (ping-pong), prefix products, pair kinds, butterfly
  // products, warp totals
  DevBuf aff_a, aff_b, aff_prefix, aff_meta, aff_others, aff_totals;
  void release() {
    scalars.release(); digits.release(); sorted.release(); counts.release(); starts.release();
    cursor.release(); poff.release(); buckets.release(); partials.release(); work.release();
    split.release(); small.release(); meta.release(); scan_tmp.release(); bases_tmp.release();
    aff_a.release(); aff_b.release(); aff_prefix.release(); aff_meta.release(); aff_others.release(); aff_totals.release();
  }
};

}  // namespace gm

struct gm_comm;  // NCCL communicator of a multi-GPU job (comm.cu)

// Lifetime: the context is reference counted.  gm_init returns it with one reference (the caller's, dropped by
// gm_shutdown); every SRS / sumcheck / msm-stream handle holds another.  gm_shutdown marks the context closed and
// gives its scratch back, but the struct, its streams and events stay valid until the last handle is freed - so a
// handle may outlive its context (Rust Drop order, Python __del__ order) and its *_free is always safe.
// Threads: entry points that take a gm_ctx* (MSM, folds, Fr vector helpers, timers) share the context's stream,
// scratch arena, result slot and events and are serialised by `mu`.  A gm_sumcheck handle owns its stream, events,
// device buffers and pinned message slot: distinct provers run concurrently from different host threads.
struct gm_ctx {
  std::atomic<int> refs{1};
  std::atomic<bool> closed{false};
  std::recursive_mutex mu;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev[8] = {};
  cudaEvent_t ev_join = nullptr;   // orders a handle's stream after the work queued on `stream` (disable-timing)
  float last_ms[4] = {0, 0, 0, 0};
  std::atomic<uint64_t> launches{0};
  int sm_count = 148;
  gm::MsmScratch msm;
  void* pinned = nullptr;  // pinned staging block: first 4 KB call results, then 64-byte message slots of sumcheck handles
  size_t pinned_bytes = 0;
  void* bounce = nullptr;  // pinned bounce buffer of gm_msm_g1 for pageable host scalars (one chunk)
  size_t bounce_bytes = 0;
  std::vector<uint32_t> free_slots;  // free 64-byte pinned slots (guarded by mu)
  void* d_result = nullptr;  // device result slot (accumulator + normalised output)
  void* d_flush = nullptr;   // 256 MB scratch written by gm_l2_flush
  gm::DevBuf fr_red;         // reduction partials + ticket + result of the Fr vector helpers
  gm::DevBuf fr_div;         // level arrays of the synthetic-division scan
  gm_comm* comm = nullptr;   // set by gm_comm_init (multi-GPU jobs)
  gm_msm_stream* host_stream = nullptr;   // internal msm stream of gm_msm_g1 for large host inputs (chunked H2D / compute overlap)
};

struct gm_srs {
  gm_ctx* ctx = nullptr;
  void* d_points = nullptr;  // n * 96 B, Montgomery x|y, (0,0) = identity
  size_t n = 0;
  bool owned = true;
  // optional precomputed multiples (gm_srs_precompute): table[w][i] = 2^(c*w) * P_i for i < prefix.  Several
  // tables over nested prefixes, each with the window size that suits MSMs of about that length: a short
  // commitment against a long SRS (the fold levels of tensorcheck) must not pay the bucket reduction of c = 22.
  struct PreTable {
    void* d_table = nullptr;
    size_t prefix = 0;
    int c = 0, W = 0;
    int rec_q = 6;  // 16-byte quads per record (8 = padded to 128 B)
  };
  PreTable pre[5];
  int npre = 0;
};

namespace gm {
inline int set_device(const gm_ctx* ctx) {
  GM_CUDA(cudaSetDevice(ctx->device));
  return GM_OK;
}
void ctx_retain(gm_ctx* ctx);
void ctx_release(gm_ctx* ctx);   // destroys the context when the last reference goes

// Where a kernel is queued: the context's shared stream, or the private stream of a sumcheck handle.
struct Lane {
  cudaStream_t stream;
  std::atomic<uint64_t>* launches;
};
inline Lane lane_of(gm_ctx* ctx) { return Lane{ctx->stream, &ctx->launches}; }
}  // namespace gm

// First statement of every entry point that works on the context's shared stream / scratch: takes the context
// lock for the rest of the call, rejects a context that was shut down, selects its device.
#define GM_ENTER(ctx)                                                         \
  std::lock_guard<std::recursive_mutex> _gm_guard((ctx)->mu);                 \
  do {                                                                        \
    if ((ctx)->closed.load()) {                                               \
      ::gm::set_error("the context was shut down");                           \
      return GM_ERR_STATE;                                                    \
    }                                                                         \
    GM_CUDA(cudaSetDevice((ctx)->device));                                    \
  } while (0)
