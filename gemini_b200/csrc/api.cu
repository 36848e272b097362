// C ABI of libgemini_b200 (include/gemini_b200.h): argument checking, host<->device staging and
// handle lifetimes.  All arithmetic happens in the kernels of msm.cu / fr.cu; there is no CPU path.
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"
#include "fr.cuh"
#include "msm.cuh"

namespace gm {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace gm

namespace gm {
void ctx_retain(gm_ctx* ctx) { ctx->refs.fetch_add(1); }
void ctx_release(gm_ctx* ctx) {
  if (ctx->refs.fetch_sub(1) != 1) return;
  // last reference: the scratch went at gm_shutdown, what is left are the streams, events and the pinned block
  cudaSetDevice(ctx->device);
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->bounce) cudaFreeHost(ctx->bounce);
  delete ctx;
}
}  // namespace gm

using namespace gm;

struct ResultSlot {  // device-side result area of a context
  XYZZ acc;
  Jacobian out;
  Affine sum_point;      // constant-scalar path: sum of the bases as one affine point
  uint32_t one[8];       // the scalar 1 as a plain integer
  uint32_t flag[4];
};

struct gm_msm_stream {
  gm_ctx* ctx = nullptr;
  const gm_srs* srs = nullptr;
  size_t chunk_cap = 0;
  XYZZ* d_acc = nullptr;
  DevBuf scal[2], pts_raw[2], pts[2];
  // device-resident buckets: every chunk only sorts and accumulates into them; the bucket reduction runs once,
  // at finalize (or when the kind of bases changes)
  DevBuf buckets, live;
  MsmPlan plan{};
  bool plan_set = false, plan_srs = false, dirty = false;
  cudaEvent_t copied[2] = {}, consumed[2] = {};
  bool used[2] = {false, false};
  unsigned turn = 0;
  bool holds_ctx_ref = true;   // false for the context's own internal stream (gm_msm_g1 of large host inputs)
};

static size_t ceil_log2(size_t x) {  // ark_std::log2
  size_t r = 0;
  while (((size_t)1 << r) < x) r++;
  return r;
}

static inline void fr_from_u64(Fr& dst, const uint64_t* src) { memcpy(dst.v, src, 32); }

extern "C" {

const char* gm_last_error(void) { return g_err; }
int gm_abi_version(void) { return 1; }

int gm_msm_describe_plan(size_t n, int with_table, int sm_count, int out[8]) {
  if (!out || sm_count <= 0) return GM_ERR_ARG;
  msm_describe_plan(n, with_table != 0, sm_count, out);
  return GM_OK;
}

int gm_init(int device_id, gm_ctx** out_ctx) {
  GM_ARG(out_ctx != nullptr, "out_ctx is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("no CUDA device available (%s); libgemini_b200 has no CPU fallback", cudaGetErrorString(e));
    return GM_ERR_CUDA;
  }
  GM_ARG(device_id >= 0 && device_id < count, "device_id out of range");
  GM_CUDA(cudaSetDevice(device_id));
  gm_ctx* ctx = new (std::nothrow) gm_ctx();
  if (!ctx) return GM_ERR_OOM;
  ctx->device = device_id;
  cudaDeviceProp prop;
  GM_CUDA(cudaGetDeviceProperties(&prop, device_id));
  ctx->sm_count = prop.multiProcessorCount;
  {
    // GM_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity (measured: no effect on the MSM gathers; default untouched)
    const char* fg = getenv("GM_L2_FETCH");
    if (fg && atoi(fg) > 0) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(fg)); cudaGetLastError(); }
  }
  GM_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  GM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (auto& ev : ctx->ev) GM_CUDA(cudaEventCreate(&ev));
  GM_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  ctx->pinned_bytes = 4096 + 1024 * sizeof(ScMailbox);  // first 4 KB: call results; then the mailboxes of the sumcheck handles
  GM_CUDA(cudaHostAlloc(&ctx->pinned, ctx->pinned_bytes, cudaHostAllocDefault));
  for (uint32_t k = (uint32_t)((ctx->pinned_bytes - 4096) / sizeof(ScMailbox)); k-- > 0;) ctx->free_slots.push_back(k);
  // short-lived vectors (prover state, DeviceFr temporaries) come from the stream-ordered pool: keep freed
  // blocks cached instead of returning them to the driver at every synchronisation
  cudaMemPool_t pool;
  GM_CUDA(cudaDeviceGetDefaultMemPool(&pool, device_id));
  unsigned long long keep = ~0ull;
  GM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  *out_ctx = ctx;
  return GM_OK;
}

// Drops the caller's reference.  Handles created from this context stay valid (their *_free works in any order
// relative to gm_shutdown); calls that need the context itself fail with GM_ERR_STATE from here on.  The context
// pointer must not be passed to gm_shutdown twice.
int gm_shutdown(gm_ctx* ctx) {
  if (!ctx) return GM_OK;
  {
    std::lock_guard<std::recursive_mutex> guard(ctx->mu);
    if (ctx->closed.exchange(true)) return GM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    comm_destroy(ctx);
    if (ctx->host_stream) { gm_msm_stream_free(ctx->host_stream); ctx->host_stream = nullptr; }
    ctx->msm.release();
    if (ctx->d_result) cudaFree(ctx->d_result);
    if (ctx->d_flush) cudaFree(ctx->d_flush);
    ctx->d_result = ctx->d_flush = nullptr;
    if (ctx->bounce) cudaFreeHost(ctx->bounce);
    ctx->bounce = nullptr;
    ctx->bounce_bytes = 0;
    ctx->fr_red.release();
    ctx->fr_div.release();
  }
  ctx_release(ctx);
  return GM_OK;
}

uint64_t gm_launch_count(const gm_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }
float gm_last_device_ms(const gm_ctx* ctx, int phase) { return (ctx && phase >= 0 && phase < 4) ? ctx->last_ms[phase] : -1.f; }
int gm_device_synchronize(gm_ctx* ctx) {
  GM_ARG(ctx, "ctx is NULL");
  GM_ENTER(ctx);
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GM_OK;
}

int gm_timer_start(gm_ctx* ctx) {
  GM_ARG(ctx, "ctx is NULL");
  GM_ENTER(ctx);
  GM_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
  return GM_OK;
}
int gm_timer_stop(gm_ctx* ctx, float* out_ms) {
  GM_ARG(ctx && out_ms, "NULL argument");
  GM_ENTER(ctx);
  GM_CUDA(cudaEventRecord(ctx->ev[7], ctx->stream));
  GM_CUDA(cudaEventSynchronize(ctx->ev[7]));
  GM_CUDA(cudaEventElapsedTime(out_ms, ctx->ev[6], ctx->ev[7]));
  return GM_OK;
}
int gm_l2_flush(gm_ctx* ctx) {
  GM_ARG(ctx, "ctx is NULL");
  GM_ENTER(ctx);
  const size_t bytes = (size_t)256 << 20;  // > 126 MB L2
  if (!ctx->d_flush) GM_CUDA(cudaMalloc(&ctx->d_flush, bytes));
  GM_CUDA(cudaMemsetAsync(ctx->d_flush, 0x5a, bytes, ctx->stream));
  return GM_OK;
}

// ---- raw buffers -------------------------------------------------------------------------
int gm_dev_alloc(gm_ctx* ctx, size_t bytes, void** out_dev) {
  GM_ARG(ctx && out_dev, "NULL argument");
  GM_ENTER(ctx);
  GM_CUDA(cudaMallocAsync(out_dev, bytes ? bytes : 16, ctx->stream));
  return GM_OK;
}
int gm_dev_free(gm_ctx* ctx, void* dev) {
  GM_ARG(ctx, "ctx is NULL");
  GM_ENTER(ctx);
  if (dev) GM_CUDA(cudaFreeAsync(dev, ctx->stream));  // stream ordered: queued work that uses it finishes first
  return GM_OK;
}
int gm_dev_upload(gm_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  GM_ARG(ctx, "ctx is NULL");
  GM_ENTER(ctx);
  GM_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GM_OK;
}
int gm_dev_download(gm_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  GM_ARG(ctx, "ctx is NULL");
  GM_ENTER(ctx);
  GM_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GM_OK;
}
int gm_fr_random_dev(gm_ctx* ctx, void* out_dev, size_t n, uint64_t seed) {
  GM_ARG(ctx && (out_dev || n == 0), "NULL argument");
  GM_ENTER(ctx);
  GM_TRY(fr_random_dev(ctx, reinterpret_cast<Fr*>(out_dev), n, seed));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GM_OK;
}

// ---- SRS ---------------------------------------------------------------------------------
static int srs_alloc(gm_ctx* ctx, size_t n, gm_srs** out) {
  gm_srs* s = new (std::nothrow) gm_srs();
  if (!s) return GM_ERR_OOM;
  s->ctx = ctx;
  s->n = n;
  cudaError_t e = cudaMalloc(&s->d_points, std::max<size_t>(n, 1) * sizeof(Affine));
  if (e != cudaSuccess) {
    delete s;
    set_error("cudaMalloc of %zu SRS points failed: %s", n, cudaGetErrorString(e));
    return GM_ERR_OOM;
  }
  ctx_retain(ctx);
  *out = s;
  return GM_OK;
}

static int upload_points(gm_ctx* ctx, const void* points, size_t n, size_t stride, long inf_offset, DevBuf& raw, Affine* d_out,
                         cudaStream_t copy_on) {
  GM_ARG(stride >= 96, "stride_bytes must be >= 96");
  GM_ARG(inf_offset < (long)stride, "inf_offset outside the record");
  GM_ARG(inf_offset < 0 || inf_offset >= 96, "inf_offset overlaps the coordinates");
  if (n == 0) return GM_OK;
  if (stride == 96 && inf_offset < 0) {
    GM_CUDA(cudaMemcpyAsync(d_out, points, n * 96, cudaMemcpyHostToDevice, copy_on));
    return GM_OK;
  }
  GM_TRY(raw.reserve(n * stride));
  GM_CUDA(cudaMemcpyAsync(raw.p, points, n * stride, cudaMemcpyHostToDevice, copy_on));
  return GM_OK;
}

int gm_srs_load_g1(gm_ctx* ctx, const void* points, size_t n, size_t stride_bytes, long inf_offset, gm_srs** out_srs) {
  GM_ARG(ctx && out_srs && (points || n == 0), "NULL argument");
  GM_ENTER(ctx);
  gm_srs* s = nullptr;
  GM_TRY(srs_alloc(ctx, n, &s));
  DevBuf raw;
  int rc = upload_points(ctx, points, n, stride_bytes, inf_offset, raw, reinterpret_cast<Affine*>(s->d_points), ctx->stream);
  if (rc == GM_OK && raw.p) rc = srs_pack(ctx, raw.as<uint8_t>(), n, stride_bytes, inf_offset, reinterpret_cast<Affine*>(s->d_points));
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  raw.release();
  if (rc != GM_OK || e != cudaSuccess) {
    if (rc == GM_OK) { set_error("srs load: %s", cudaGetErrorString(e)); rc = GM_ERR_CUDA; }
    gm_srs_free(s);
    return rc;
  }
  *out_srs = s;
  return GM_OK;
}

int gm_srs_generate_g1(gm_ctx* ctx, size_t n, uint64_t first_multiple, gm_srs** out_srs) {
  GM_ARG(ctx && out_srs, "NULL argument");
  GM_ENTER(ctx);
  gm_srs* s = nullptr;
  GM_TRY(srs_alloc(ctx, n, &s));
  int rc = srs_generate(ctx, n, first_multiple, reinterpret_cast<Affine*>(s->d_points));
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc != GM_OK || e != cudaSuccess) {
    if (rc == GM_OK) { set_error("srs generate: %s", cudaGetErrorString(e)); rc = GM_ERR_CUDA; }
    gm_srs_free(s);
    return rc;
  }
  *out_srs = s;
  return GM_OK;
}

int gm_srs_setup_g1(gm_ctx* ctx, const uint64_t g_xy[12], const uint64_t tau[4], size_t n, gm_srs** out_srs) {
  GM_ARG(ctx && out_srs && g_xy && tau, "NULL argument");
  GM_ENTER(ctx);
  gm_srs* s = nullptr;
  GM_TRY(srs_alloc(ctx, n, &s));
  Affine g;
  memcpy(&g, g_xy, 96);
  Fr t;
  memcpy(t.v, tau, 32);
  void* d_pow = nullptr;
  cudaError_t e = cudaMalloc(&d_pow, std::max<size_t>(n, 1) * 32);
  int rc = GM_OK;
  if (e != cudaSuccess) { set_error("srs setup: cudaMalloc failed: %s", cudaGetErrorString(e)); rc = GM_ERR_OOM; }
  if (rc == GM_OK) rc = fr_powers_dev(ctx, t, n, reinterpret_cast<Fr*>(d_pow));          // misc::powers(tau, n)
  if (rc == GM_OK) rc = srs_fixed_base(ctx, g, reinterpret_cast<const uint32_t*>(d_pow), n, reinterpret_cast<Affine*>(s->d_points));
  e = cudaStreamSynchronize(ctx->stream);
  if (d_pow) cudaFree(d_pow);
  if (rc != GM_OK || e != cudaSuccess) {
    if (rc == GM_OK) { set_error("srs setup: %s", cudaGetErrorString(e)); rc = GM_ERR_CUDA; }
    gm_srs_free(s);
    return rc;
  }
  *out_srs = s;
  return GM_OK;
}

int gm_srs_fill_g1(gm_ctx* ctx, const uint64_t point_xy[12], size_t n, gm_srs** out_srs) {
  GM_ARG(ctx && out_srs && point_xy, "NULL argument");
  GM_ENTER(ctx);
  gm_srs* s = nullptr;
  GM_TRY(srs_alloc(ctx, n, &s));
  Affine p;
  memcpy(&p, point_xy, 96);
  int rc = srs_fill(ctx, p, n, reinterpret_cast<Affine*>(s->d_points));
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc != GM_OK || e != cudaSuccess) {
    if (rc == GM_OK) { set_error("srs fill: %s", cudaGetErrorString(e)); rc = GM_ERR_CUDA; }
    gm_srs_free(s);
    return rc;
  }
  *out_srs = s;
  return GM_OK;
}

static void srs_drop_tables(gm_srs* srs) {
  for (int k = 0; k < srs->npre; k++) {
    if (srs->pre[k].d_table) cudaFree(srs->pre[k].d_table);
    srs->pre[k] = gm_srs::PreTable();
  }
  srs->npre = 0;
}

int gm_srs_precompute(gm_ctx* ctx, gm_srs* srs, size_t expected_msm_len) {
  GM_ARG(ctx && srs, "NULL argument");
  GM_ENTER(ctx);
  if (srs->n == 0) return GM_OK;
  srs_drop_tables(srs);
  // table 0 covers the whole SRS; tables 1..4 cover prefixes 8x, 64x, 512x and 4096x shorter (kept while >= 2^12 points),
  // each with the window size its own length calls for: a short commitment (the 23 fold levels of tensorcheck go down
  // to 2 coefficients) then reduces a few thousand buckets instead of the 2^21 of the full table
  size_t prefix = srs->n;
  size_t expect = expected_msm_len ? std::min(expected_msm_len, srs->n) : srs->n;
  for (int k = 0; k < 5; k++) {
    if (k > 0 && prefix < ((size_t)1 << 12)) break;
    const MsmPlan P = msm_plan_merged(std::min(expect, prefix), 0);
    GM_ARG((double)P.W * (double)prefix < 2147483648.0, "SRS too large for a precomputed table (W * n must stay below 2^31)");
    void* table = nullptr;
    // GM_TABLE_PAD=1 pads the records to one 128-byte line each: a third less DRAM traffic in the gathers (ncu:
    // 45.1 -> 32.3 GB for k_aff_finish<1> at 2^24) but no time gained - the gathers are latency-, not
    // bandwidth-bound - at 33 % more HBM, so packed 96-byte records stay the default.
    const char* pad_env = getenv("GM_TABLE_PAD");
    const int rec_q = (pad_env && atoi(pad_env)) ? 8 : 6;
    const size_t bytes = (size_t)P.W * prefix * rec_q * 16;
    cudaError_t e = cudaMalloc(&table, bytes);
    if (e != cudaSuccess) {
      set_error("precompute: cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
      srs_drop_tables(srs);
      return GM_ERR_OOM;
    }
    int rc = msm_precompute(ctx, reinterpret_cast<const Affine*>(srs->d_points), prefix, P.c, P.W, rec_q, reinterpret_cast<Affine*>(table));
    e = cudaStreamSynchronize(ctx->stream);
    if (rc != GM_OK || e != cudaSuccess) {
      if (rc == GM_OK) { set_error("precompute: %s", cudaGetErrorString(e)); rc = GM_ERR_CUDA; }
      cudaFree(table);
      srs_drop_tables(srs);
      return rc;
    }
    srs->pre[k].d_table = table;
    srs->pre[k].prefix = prefix;
    srs->pre[k].c = P.c;
    srs->pre[k].W = P.W;
    srs->pre[k].rec_q = rec_q;
    srs->npre = k + 1;
    prefix >>= 3;
    expect = prefix;
  }
  return GM_OK;
}

int gm_srs_precompute_info(const gm_srs* srs, int* out_window_bits, int* out_levels) {
  GM_ARG(srs, "NULL argument");
  if (out_window_bits) *out_window_bits = srs->npre ? srs->pre[0].c : 0;
  if (out_levels) *out_levels = srs->npre ? srs->pre[0].W : 0;
  return GM_OK;
}

size_t gm_srs_len(const gm_srs* srs) { return srs ? srs->n : 0; }

int gm_srs_read(gm_ctx* ctx, const gm_srs* srs, size_t offset, size_t n, uint64_t* out_xy) {
  GM_ARG(ctx && srs && out_xy, "NULL argument");
  GM_ARG(offset <= srs->n && n <= srs->n - offset, "range outside the SRS");
  GM_ENTER(ctx);
  GM_CUDA(cudaMemcpyAsync(out_xy, reinterpret_cast<const Affine*>(srs->d_points) + offset, n * sizeof(Affine),
                          cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GM_OK;
}

// Valid before or after gm_shutdown of the context the SRS came from (the handle keeps the context alive).
int gm_srs_free(gm_srs* srs) {
  if (!srs) return GM_OK;
  gm_ctx* ctx = srs->ctx;
  {
    std::lock_guard<std::recursive_mutex> guard(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);   // queued MSMs that read these points finish first
    if (srs->owned && srs->d_points) cudaFree(srs->d_points);
    srs_drop_tables(srs);
  }
  delete srs;
  ctx_release(ctx);
  return GM_OK;
}

// ---- MSM ---------------------------------------------------------------------------------
static int ensure_result(gm_ctx* ctx, ResultSlot** slot) {
  if (!ctx->d_result) {
    GM_CUDA(cudaMalloc(&ctx->d_result, sizeof(ResultSlot)));
    const uint32_t one[8] = {1u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    GM_CUDA(cudaMemcpy(reinterpret_cast<ResultSlot*>(ctx->d_result)->one, one, sizeof(one), cudaMemcpyHostToDevice));
  }
  *slot = reinterpret_cast<ResultSlot*>(ctx->d_result);
  return GM_OK;
}

static void record_phases(gm_ctx* ctx) {
  cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&ctx->last_ms[1], ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&ctx->last_ms[2], ctx->ev[3], ctx->ev[4]);
  cudaEventElapsedTime(&ctx->last_ms[3], ctx->ev[4], ctx->ev[5]);
}

// bases for an MSM over srs[base_offset, base_offset + n): the smallest precomputed table that covers the range
static MsmBases bases_of_srs(const gm_srs* srs, size_t base_offset, size_t n, bool full_table_only = false) {
  MsmBases b;
  b.points = reinterpret_cast<const Affine*>(srs->d_points);
  b.n = srs->n;
  for (int k = full_table_only ? 0 : srs->npre - 1; k >= 0; k--) {  // tables are ordered by decreasing prefix
    if (base_offset + n <= srs->pre[k].prefix) {
      b.table = reinterpret_cast<const Affine*>(srs->pre[k].d_table);
      b.n = srs->pre[k].prefix;
      b.c = srs->pre[k].c;
      b.W = srs->pre[k].W;
      b.rec_q = srs->pre[k].rec_q;
      break;
    }
  }
  return b;
}
static MsmBases bases_of_points(const Affine* pts, size_t n) {
  MsmBases b;
  b.points = pts;
  b.n = n;
  return b;
}

// `sharded`: this rank's partial sum is exchanged with the other ranks of the context's communicator (one
// ncclAllGather of 192-byte XYZZ points on the library stream) and every rank returns the same total.
static int exchange_partials(gm_ctx* ctx, XYZZ* d_acc) {
  if (comm_world(ctx) <= 1) return GM_OK;
  void* d_all = nullptr;
  GM_TRY(comm_allgather_dev(ctx, d_acc, sizeof(XYZZ), &d_all));
  return msm_acc_set_sum_xyzz(ctx, d_all, (size_t)comm_world(ctx), sizeof(XYZZ), d_acc);
}

// the accumulator of the running call -> (exchange between ranks) -> normalised 144-byte result on the host
static int msm_finish(gm_ctx* ctx, ResultSlot* slot, size_t n, uint64_t out[18], bool sharded) {
  if (sharded) GM_TRY(exchange_partials(ctx, &slot->acc));
  GM_TRY(msm_acc_normalize(ctx, &slot->acc, &slot->out));
  GM_CUDA(cudaMemcpyAsync(ctx->pinned, &slot->out, sizeof(Jacobian), cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->pinned, sizeof(Jacobian));
  if (n) record_phases(ctx); else cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
  return GM_OK;
}

// Constant scalar vectors (the reference's dummy_r1cs witness, circuit.rs:349-365): sum_i s P_i = s * (sum_i P_i) - n
// additions and one scalar multiplication instead of n * windows additions that all land in one bucket per window.
// Both steps run through the ordinary pipeline: the sum of the bases is the MSM with every scalar = 1 (scalar stride 0
// over a device constant), its result becomes a one-point base, and s * S is the MSM of that point with scalars[0].
// GM_CONST_SCALAR_MIN: smallest n that is tested for it (0 = never); the test costs one 4-byte round trip.
static size_t const_scalar_min() {
  static const size_t v = [] {
    const char* e = getenv("GM_CONST_SCALAR_MIN");
    return e ? (size_t)strtoull(e, nullptr, 10) : ((size_t)1 << 16);
  }();
  return v;
}

static int msm_common(gm_ctx* ctx, const MsmBases& bases, size_t base_offset, const uint32_t* d_scalars, size_t n, bool bigint, uint64_t out[18],
                      bool sharded = false) {
  ResultSlot* slot;
  GM_TRY(ensure_result(ctx, &slot));
  bool constant = false;
  if (const_scalar_min() && n >= const_scalar_min())
    GM_TRY(msm_scalars_all_equal(ctx, d_scalars, n, ctx->msm.scalar_stride, slot->flag, reinterpret_cast<uint32_t*>(ctx->pinned) + 64, &constant));
  GM_TRY(msm_acc_reset(ctx, &slot->acc));
  if (constant) {
    const size_t stride = ctx->msm.scalar_stride;
    ctx->msm.scalar_stride = 0;                                   // every term reads the same scalar: 1
    int rc = msm_accumulate(ctx, bases, base_offset, slot->one, n, /*bigint=*/true, &slot->acc);
    ctx->msm.scalar_stride = 1;
    if (rc == GM_OK) rc = msm_acc_to_affine(ctx, &slot->acc, &slot->sum_point);
    if (rc == GM_OK) rc = msm_acc_reset(ctx, &slot->acc);
    MsmBases one_point;
    one_point.points = &slot->sum_point;
    one_point.n = 1;
    if (rc == GM_OK) rc = msm_accumulate(ctx, one_point, 0, d_scalars, 1, bigint, &slot->acc);
    ctx->msm.scalar_stride = stride;
    GM_TRY(rc);
  } else {
    GM_TRY(msm_accumulate(ctx, bases, base_offset, d_scalars, n, bigint, &slot->acc));
  }
  return msm_finish(ctx, slot, n, out, sharded);
}

int gm_msm_g1_dev(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const void* scalars_dev, size_t n,
                  int scalars_are_bigint, uint64_t out_jacobian[18]) {
  GM_ARG(ctx && srs && out_jacobian && (scalars_dev || n == 0), "NULL argument");
  GM_ARG(base_offset <= srs->n, "base_offset beyond the SRS");
  GM_ENTER(ctx);
  n = std::min(n, srs->n - base_offset);  // msm_unchecked truncates to the shorter input
  GM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  return msm_common(ctx, bases_of_srs(srs, base_offset, n), base_offset, reinterpret_cast<const uint32_t*>(scalars_dev), n, scalars_are_bigint != 0,
                    out_jacobian);
}

static int stream_create(gm_ctx* ctx, const gm_srs* srs_or_null, size_t chunk_cap, bool holds_ctx_ref, gm_msm_stream** out);
static int stream_flush(gm_msm_stream* s);

// Host scalars.  A large input is fed through the context's internal msm stream in a few chunks: the H2D copy of
// chunk k+1 (copy stream, double-buffered staging) overlaps the sort and bucket accumulation of chunk k, the buckets stay
// resident and the reduction runs once - msm_chunks (src/kzg/space.rs:22-55) applied to our own entry point.  It is what
// makes a call from ordinary pageable memory (an arkworks &[Fr]) cost little more than one from pinned memory: the
// driver's staged pageable copy (about 11 GB/s) hides behind the kernels instead of preceding them.
// GM_E2E_CHUNKS overrides the number of chunks (1 = one copy, then the one-shot pipeline).
static int msm_host(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const uint64_t* scalars, size_t n, bool bigint, uint64_t out[18],
                    bool sharded) {
  GM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  int chunks = 1;
  if (n >= ((size_t)1 << 23)) chunks = 4;
  if (const char* e = getenv("GM_E2E_CHUNKS")) { if (atoi(e) > 0) chunks = atoi(e); }
  if (chunks <= 1 || n < (size_t)chunks) {
    GM_TRY(ctx->msm.scalars.reserve(std::max<size_t>(n, 1) * 32));
    if (n) GM_CUDA(cudaMemcpyAsync(ctx->msm.scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    return msm_common(ctx, bases_of_srs(srs, base_offset, n), base_offset, ctx->msm.scalars.as<uint32_t>(), n, bigint, out, sharded);
  }
  // PINNED host memory: one-shot pipeline fed piecewise - the digits of a piece are extracted while the next piece is on
  // the bus, and nothing else waits for the transfer (measured at 2^24: 77.3 -> see profiles/r02_summary.md)
  if (!getenv("GM_E2E_CHUNKS")) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, scalars) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
      ResultSlot* slot;
      GM_TRY(ensure_result(ctx, &slot));
      GM_TRY(msm_acc_reset(ctx, &slot->acc));
      const int rc = msm_accumulate_pinned(ctx, bases_of_srs(srs, base_offset, n), base_offset, scalars, n, bigint, /*pieces=*/8, &slot->acc);
      if (rc == GM_OK) return msm_finish(ctx, slot, n, out, sharded);
      if (rc != GM_ERR_ARG) return rc;          // GM_ERR_ARG: several passes needed - the streamed path below handles any size
    } else {
      cudaGetLastError();
    }
  }
  if (!ctx->host_stream) GM_TRY(stream_create(ctx, nullptr, 0, /*holds_ctx_ref=*/false, &ctx->host_stream));
  gm_msm_stream* s = ctx->host_stream;
  const size_t step = (n + chunks - 1) / chunks;
  GM_TRY(stream_flush(s));                      // nothing pending: resets the plan
  GM_CUDA(cudaMemsetAsync(s->d_acc, 0, sizeof(XYZZ), ctx->stream));
  s->srs = srs;
  s->chunk_cap = step;
  // PAGEABLE memory (what an arkworks &[Fr] is): the driver's own bounce copy runs at about 11 GB/s on one thread.  Each
  // chunk is instead gathered into a pinned bounce buffer of the context by a few host threads and DMA'd from there -
  // both behind the kernels of the previous chunk.  GM_HOST_COPY_THREADS = 0 leaves the copy to the driver.
  static const int copy_threads = [] {
    const char* e = getenv("GM_HOST_COPY_THREADS");
    const int hw = (int)std::thread::hardware_concurrency();
    return e ? atoi(e) : std::max(1, std::min(8, hw / 2));
  }();
  uint64_t* bounce = nullptr;
  if (copy_threads > 0) {
    if (ctx->bounce_bytes < step * 32) {
      if (ctx->bounce) cudaFreeHost(ctx->bounce);
      ctx->bounce = nullptr;
      ctx->bounce_bytes = 0;
      if (cudaHostAlloc(&ctx->bounce, step * 32, cudaHostAllocDefault) == cudaSuccess) ctx->bounce_bytes = step * 32;
      else cudaGetLastError();               // no pinned memory to spare: the driver's path still works
    }
    bounce = reinterpret_cast<uint64_t*>(ctx->bounce);
  }
  for (size_t off = 0; off < n; off += step) {
    const size_t m = std::min(step, n - off);
    const uint64_t* src = scalars + 4 * off;
    if (bounce != nullptr) {
      const size_t bytes = m * 32, per = (bytes / copy_threads + 4095) & ~(size_t)4095;
      std::vector<std::thread> pool;
      for (int t = 1; t < copy_threads; t++) {
        const size_t b0 = std::min(bytes, (size_t)t * per), b1 = std::min(bytes, b0 + per);
        if (b1 > b0) pool.emplace_back([=] { memcpy(reinterpret_cast<uint8_t*>(bounce) + b0, reinterpret_cast<const uint8_t*>(src) + b0, b1 - b0); });
      }
      memcpy(bounce, src, std::min(bytes, per));
      for (auto& th : pool) th.join();
      src = bounce;                          // gm_msm_stream_push returns once its copy has left the buffer
    }
    GM_TRY(gm_msm_stream_push(s, nullptr, 0, -1, base_offset + off, src, m, bigint ? 1 : 0));
  }
  const int rc = sharded ? gm_msm_stream_finalize_sharded(s, out) : gm_msm_stream_finalize(s, out);
  s->srs = nullptr;
  GM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  GM_CUDA(cudaEventSynchronize(ctx->ev[1]));
  cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
  return rc;
}

int gm_msm_g1(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const uint64_t* scalars, size_t n,
              int scalars_are_bigint, uint64_t out_jacobian[18]) {
  GM_ARG(ctx && srs && out_jacobian && (scalars || n == 0), "NULL argument");
  GM_ARG(base_offset <= srs->n, "base_offset beyond the SRS");
  GM_ENTER(ctx);
  n = std::min(n, srs->n - base_offset);
  return msm_host(ctx, srs, base_offset, scalars, n, scalars_are_bigint != 0, out_jacobian, /*sharded=*/false);
}

// Multi-GPU MSM: `srs` holds THIS rank's contiguous range of the points and `scalars` the matching
// range of the scalars; every rank calls with its own shard and gets the sum over all ranks.
int gm_msm_g1_sharded_dev(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const void* scalars_dev, size_t n,
                          int scalars_are_bigint, uint64_t out_jacobian[18]) {
  GM_ARG(ctx && srs && out_jacobian && (scalars_dev || n == 0), "NULL argument");
  GM_ARG(base_offset <= srs->n, "base_offset beyond the SRS");
  GM_ENTER(ctx);
  n = std::min(n, srs->n - base_offset);
  GM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  return msm_common(ctx, bases_of_srs(srs, base_offset, n), base_offset, reinterpret_cast<const uint32_t*>(scalars_dev), n, scalars_are_bigint != 0,
                    out_jacobian, /*sharded=*/true);
}

// General device-scalar entry: term i uses scalars_dev[i * scalar_stride].  With a world of W ranks dealing points out
// cyclically (rank r holds P_r, P_{r+W}, ...: every vector, whatever its length, splits evenly - the fold levels of
// tensorcheck halve 23 times), rank r passes scalars_dev = v + r, scalar_stride = W, n = ceil((len - r) / W).
int gm_msm_g1_strided_dev(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const void* scalars_dev, size_t n, size_t scalar_stride,
                          int scalars_are_bigint, int sharded, uint64_t out_jacobian[18]) {
  GM_ARG(ctx && srs && out_jacobian && (scalars_dev || n == 0), "NULL argument");
  GM_ARG(base_offset <= srs->n && scalar_stride >= 1, "base_offset beyond the SRS, or a zero stride");
  GM_ENTER(ctx);
  n = std::min(n, srs->n - base_offset);
  GM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  ctx->msm.scalar_stride = scalar_stride;
  const int rc = msm_common(ctx, bases_of_srs(srs, base_offset, n), base_offset, reinterpret_cast<const uint32_t*>(scalars_dev), n,
                            scalars_are_bigint != 0, out_jacobian, sharded != 0);
  ctx->msm.scalar_stride = 1;
  return rc;
}

// every `stride`-th point of an SRS, starting at `first`: the cyclic shard of one rank, cut from a resident key
int gm_srs_subsample(gm_ctx* ctx, const gm_srs* srs, size_t first, size_t stride, size_t count, gm_srs** out_srs) {
  GM_ARG(ctx && srs && out_srs && stride >= 1, "bad argument");
  GM_ARG(count == 0 || first + (count - 1) * stride < srs->n, "range outside the SRS");
  GM_ENTER(ctx);
  gm_srs* s = nullptr;
  GM_TRY(srs_alloc(ctx, count, &s));
  cudaError_t e = count ? cudaMemcpy2DAsync(s->d_points, sizeof(Affine), reinterpret_cast<const Affine*>(srs->d_points) + first,
                                            stride * sizeof(Affine), sizeof(Affine), count, cudaMemcpyDeviceToDevice, ctx->stream)
                        : cudaSuccess;
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { set_error("srs subsample: %s", cudaGetErrorString(e)); gm_srs_free(s); return GM_ERR_CUDA; }
  *out_srs = s;
  return GM_OK;
}

int gm_msm_g1_sharded(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const uint64_t* scalars, size_t n,
                      int scalars_are_bigint, uint64_t out_jacobian[18]) {
  GM_ARG(ctx && srs && out_jacobian && (scalars || n == 0), "NULL argument");
  GM_ARG(base_offset <= srs->n, "base_offset beyond the SRS");
  GM_ENTER(ctx);
  n = std::min(n, srs->n - base_offset);
  return msm_host(ctx, srs, base_offset, scalars, n, scalars_are_bigint != 0, out_jacobian, /*sharded=*/true);
}

int gm_msm_g1_checked(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, size_t bases_len, const uint64_t* scalars,
                      size_t scalars_len, uint64_t out_jacobian[18], size_t* out_min_len) {
  GM_ARG(ctx && srs && out_jacobian, "NULL argument");
  GM_ARG(base_offset <= srs->n && bases_len <= srs->n - base_offset, "bases range outside the SRS");
  if (bases_len != scalars_len) {
    if (out_min_len) *out_min_len = std::min(bases_len, scalars_len);
    set_error("msm: bases.len() = %zu != scalars.len() = %zu", bases_len, scalars_len);
    return GM_ERR_LENGTH;
  }
  return gm_msm_g1(ctx, srs, base_offset, scalars, scalars_len, 0, out_jacobian);
}

int gm_msm_g1_hostbases(gm_ctx* ctx, const void* points, size_t stride_bytes, long inf_offset, const uint64_t* scalars,
                        size_t n, int scalars_are_bigint, uint64_t out_jacobian[18]) {
  GM_ARG(ctx && out_jacobian && ((points && scalars) || n == 0), "NULL argument");
  GM_ENTER(ctx);
  GM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  MsmScratch& S = ctx->msm;
  GM_TRY(S.bases_tmp.reserve(std::max<size_t>(n, 1) * sizeof(Affine)));
  DevBuf raw;
  int rc = upload_points(ctx, points, n, stride_bytes, inf_offset, raw, S.bases_tmp.as<Affine>(), ctx->stream);
  if (rc == GM_OK && raw.p) rc = srs_pack(ctx, raw.as<uint8_t>(), n, stride_bytes, inf_offset, S.bases_tmp.as<Affine>());
  if (rc == GM_OK) rc = S.scalars.reserve(std::max<size_t>(n, 1) * 32);
  if (rc == GM_OK && n) {
    cudaError_t e = cudaMemcpyAsync(S.scalars.p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { set_error("H2D scalars: %s", cudaGetErrorString(e)); rc = GM_ERR_CUDA; }
  }
  if (rc == GM_OK) rc = msm_common(ctx, bases_of_points(S.bases_tmp.as<Affine>(), n), 0, S.scalars.as<uint32_t>(), n, scalars_are_bigint != 0, out_jacobian);
  cudaStreamSynchronize(ctx->stream);
  raw.release();
  return rc;
}

int gm_g1_sum(gm_ctx* ctx, const uint64_t* jacobians, size_t k, uint64_t out_jacobian[18]) {
  GM_ARG(ctx && out_jacobian && (jacobians || k == 0), "NULL argument");
  GM_ENTER(ctx);
  ResultSlot* slot;
  GM_TRY(ensure_result(ctx, &slot));
  GM_TRY(ctx->msm.bases_tmp.reserve(std::max<size_t>(k, 1) * sizeof(Jacobian)));
  if (k) GM_CUDA(cudaMemcpyAsync(ctx->msm.bases_tmp.p, jacobians, k * sizeof(Jacobian), cudaMemcpyHostToDevice, ctx->stream));
  GM_TRY(msm_acc_reset(ctx, &slot->acc));
  if (k) GM_TRY(msm_acc_add_jacobians(ctx, ctx->msm.bases_tmp.as<Jacobian>(), k, &slot->acc));
  GM_TRY(msm_acc_normalize(ctx, &slot->acc, &slot->out));
  GM_CUDA(cudaMemcpyAsync(ctx->pinned, &slot->out, sizeof(Jacobian), cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out_jacobian, ctx->pinned, sizeof(Jacobian));
  return GM_OK;
}

// ---- streamed MSM ------------------------------------------------------------------------
static int stream_create(gm_ctx* ctx, const gm_srs* srs_or_null, size_t chunk_cap, bool holds_ctx_ref, gm_msm_stream** out);
static int stream_flush(gm_msm_stream* s) {
  if (s->plan_set && s->dirty) {
    GM_TRY(msm_stream_reduce(s->ctx, s->plan, s->buckets.as<XYZZ>(), s->live.as<uint32_t>(), s->d_acc));
    GM_CUDA(cudaMemsetAsync(s->live.p, 0, msm_plan_buckets(s->plan) * 4, s->ctx->stream));
  }
  s->dirty = false;
  s->plan_set = false;
  return GM_OK;
}

int gm_msm_stream_new(gm_ctx* ctx, const gm_srs* srs_or_null, size_t chunk_cap, gm_msm_stream** out) {
  GM_ARG(ctx && out, "NULL argument");
  GM_ENTER(ctx);
  return stream_create(ctx, srs_or_null, chunk_cap, /*holds_ctx_ref=*/true, out);
}

static int stream_create(gm_ctx* ctx, const gm_srs* srs_or_null, size_t chunk_cap, bool holds_ctx_ref, gm_msm_stream** out) {
  gm_msm_stream* s = new (std::nothrow) gm_msm_stream();
  if (!s) return GM_ERR_OOM;
  s->ctx = ctx;
  s->holds_ctx_ref = holds_ctx_ref;
  if (holds_ctx_ref) ctx_retain(ctx);
  s->srs = srs_or_null;
  s->chunk_cap = chunk_cap;
  cudaError_t e = cudaMalloc(&s->d_acc, sizeof(XYZZ));
  if (e == cudaSuccess) e = cudaMemsetAsync(s->d_acc, 0, sizeof(XYZZ), ctx->stream);
  for (int k = 0; k < 2 && e == cudaSuccess; k++) {
    e = cudaEventCreateWithFlags(&s->copied[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->consumed[k], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    set_error("msm stream: %s", cudaGetErrorString(e));
    gm_msm_stream_free(s);
    return GM_ERR_CUDA;
  }
  *out = s;
  return GM_OK;
}

// scalars already resident on the device (the quotients and fold levels of the elastic prover never leave HBM):
// bases are the SRS range [base_offset, base_offset + m); no staging copy, the chunk is consumed in place
int gm_msm_stream_push_dev(gm_msm_stream* s, size_t base_offset, const void* scalars_dev, size_t m, int scalars_are_bigint) {
  GM_ARG(s && (scalars_dev || m == 0), "NULL argument");
  gm_ctx* ctx = s->ctx;
  GM_ENTER(ctx);
  if (m == 0) return GM_OK;
  GM_ARG(s->srs != nullptr, "stream has no SRS");
  GM_ARG(base_offset <= s->srs->n && m <= s->srs->n - base_offset, "base range outside the SRS");
  const MsmBases bases = bases_of_srs(s->srs, base_offset, m, /*full_table_only=*/true);
  if (s->plan_set && (!s->plan_srs || m > std::max<size_t>(s->chunk_cap, 1))) GM_TRY(stream_flush(s));
  if (!s->plan_set) {
    s->plan = msm_stream_plan(bases, std::max(s->chunk_cap, m));
    s->chunk_cap = std::max(s->chunk_cap, m);
    s->plan_srs = true;
    s->plan_set = true;
    const size_t M = msm_plan_buckets(s->plan);
    GM_TRY(s->buckets.reserve(M * sizeof(XYZZ)));
    GM_TRY(s->live.reserve(M * 4));
    GM_CUDA(cudaMemsetAsync(s->live.p, 0, M * 4, ctx->stream));
  }
  GM_TRY(msm_stream_push(ctx, bases, base_offset, reinterpret_cast<const uint32_t*>(scalars_dev), m, scalars_are_bigint != 0, s->plan,
                         s->buckets.as<XYZZ>(), s->live.as<uint32_t>()));
  s->dirty = true;
  return GM_OK;
}

int gm_msm_stream_push(gm_msm_stream* s, const void* points, size_t stride_bytes, long inf_offset, size_t base_offset,
                       const uint64_t* scalars, size_t m, int scalars_are_bigint) {
  GM_ARG(s && (scalars || m == 0), "NULL argument");
  gm_ctx* ctx = s->ctx;
  GM_ENTER(ctx);
  if (m == 0) return GM_OK;
  MsmBases bases;
  size_t boff = 0;
  const unsigned b = s->turn & 1u;
  // the staging buffers of this slot may still be read by the chunk pushed two calls ago
  if (s->used[b]) GM_CUDA(cudaEventSynchronize(s->consumed[b]));
  GM_TRY(s->scal[b].reserve(m * 32));
  GM_CUDA(cudaMemcpyAsync(s->scal[b].p, scalars, m * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
  bool need_pack = false;
  const bool from_srs = points == nullptr;
  if (!from_srs) {
    GM_TRY(s->pts[b].reserve(m * sizeof(Affine)));
    GM_TRY(upload_points(ctx, points, m, stride_bytes, inf_offset, s->pts_raw[b], s->pts[b].as<Affine>(), ctx->copy_stream));
    need_pack = !(stride_bytes == 96 && inf_offset < 0);
    bases = bases_of_points(s->pts[b].as<Affine>(), m);
  } else {
    GM_ARG(s->srs != nullptr, "stream has no SRS and no points were supplied");
    GM_ARG(base_offset <= s->srs->n && m <= s->srs->n - base_offset, "base range outside the SRS");
    bases = bases_of_srs(s->srs, base_offset, m, /*full_table_only=*/true);
    boff = base_offset;
  }
  GM_CUDA(cudaEventRecord(s->copied[b], ctx->copy_stream));
  // the caller may reuse its buffers as soon as we return: wait for the copies (the previous chunk's
  // kernels keep running on ctx->stream meanwhile - this is the H2D / compute overlap)
  GM_CUDA(cudaEventSynchronize(s->copied[b]));
  GM_CUDA(cudaStreamWaitEvent(ctx->stream, s->copied[b], 0));
  if (need_pack) GM_TRY(srs_pack(ctx, s->pts_raw[b].as<uint8_t>(), m, stride_bytes, inf_offset, s->pts[b].as<Affine>()));
  // one bucket layout per stream; a change of the kind of bases (SRS range <-> ad-hoc points) or a chunk larger
  // than the plan was made for first folds the current buckets into the accumulator
  if (s->plan_set && (s->plan_srs != from_srs || m > std::max<size_t>(s->chunk_cap, 1))) GM_TRY(stream_flush(s));
  if (!s->plan_set) {
    s->plan = msm_stream_plan(bases, std::max(s->chunk_cap, m));
    s->chunk_cap = std::max(s->chunk_cap, m);
    s->plan_srs = from_srs;
    s->plan_set = true;
    const size_t M = msm_plan_buckets(s->plan);
    GM_TRY(s->buckets.reserve(M * sizeof(XYZZ)));
    GM_TRY(s->live.reserve(M * 4));
    GM_CUDA(cudaMemsetAsync(s->live.p, 0, M * 4, ctx->stream));
  }
  GM_TRY(msm_stream_push(ctx, bases, boff, s->scal[b].as<uint32_t>(), m, scalars_are_bigint != 0, s->plan, s->buckets.as<XYZZ>(),
                         s->live.as<uint32_t>()));
  s->dirty = true;
  GM_CUDA(cudaEventRecord(s->consumed[b], ctx->stream));
  s->used[b] = true;
  s->turn++;
  return GM_OK;
}

static int stream_finalize(gm_msm_stream* s, uint64_t out_jacobian[18], bool sharded);
int gm_msm_stream_finalize(gm_msm_stream* s, uint64_t out_jacobian[18]) { return stream_finalize(s, out_jacobian, false); }
// streamed MSM of a multi-GPU job (config 5): every rank streamed its own range; the totals are exchanged once, here
int gm_msm_stream_finalize_sharded(gm_msm_stream* s, uint64_t out_jacobian[18]) { return stream_finalize(s, out_jacobian, true); }

static int stream_finalize(gm_msm_stream* s, uint64_t out_jacobian[18], bool sharded) {
  GM_ARG(s && out_jacobian, "NULL argument");
  gm_ctx* ctx = s->ctx;
  GM_ENTER(ctx);
  ResultSlot* slot;
  GM_TRY(ensure_result(ctx, &slot));
  GM_TRY(stream_flush(s));
  if (sharded) GM_TRY(exchange_partials(ctx, s->d_acc));
  GM_TRY(msm_acc_normalize(ctx, s->d_acc, &slot->out));
  GM_CUDA(cudaMemcpyAsync(ctx->pinned, &slot->out, sizeof(Jacobian), cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out_jacobian, ctx->pinned, sizeof(Jacobian));
  return GM_OK;
}

int gm_msm_stream_free(gm_msm_stream* s) {
  if (!s) return GM_OK;
  gm_ctx* ctx = s->ctx;
  {
    std::lock_guard<std::recursive_mutex> guard(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    if (s->d_acc) cudaFree(s->d_acc);
    for (int k = 0; k < 2; k++) {
      s->scal[k].release(); s->pts_raw[k].release(); s->pts[k].release();
      if (k == 0) { s->buckets.release(); s->live.release(); }
      if (s->copied[k]) cudaEventDestroy(s->copied[k]);
      if (s->consumed[k]) cudaEventDestroy(s->consumed[k]);
    }
  }
  const bool release = s->holds_ctx_ref;
  delete s;
  if (release) ctx_release(ctx);
  return GM_OK;
}

// ---- Fr folds ----------------------------------------------------------------------------
int gm_fr_fold_dev(gm_ctx* ctx, const void* f_dev, size_t n, const uint64_t r[4], void* out_dev) {
  GM_ARG(ctx && r && ((f_dev && out_dev) || n == 0), "NULL argument");
  GM_ENTER(ctx);
  Fr rr;
  fr_from_u64(rr, r);
  GM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  GM_TRY(fr_fold_dev(ctx, reinterpret_cast<const Fr*>(f_dev), n, rr, reinterpret_cast<Fr*>(out_dev)));
  GM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
  return GM_OK;
}

int gm_fr_fold(gm_ctx* ctx, const uint64_t* f, size_t n, const uint64_t r[4], uint64_t* out) {
  GM_ARG(ctx && r && ((f && out) || n == 0), "NULL argument");
  GM_ENTER(ctx);
  if (n == 0) return GM_OK;
  const size_t half = (n + 1) / 2;
  DevBuf& in = ctx->msm.scalars;
  DevBuf& res = ctx->msm.bases_tmp;
  GM_TRY(in.reserve(n * 32));
  GM_TRY(res.reserve(half * 32));
  Fr rr;
  fr_from_u64(rr, r);
  GM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  GM_CUDA(cudaMemcpyAsync(in.p, f, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  GM_TRY(fr_fold_dev(ctx, in.as<Fr>(), n, rr, res.as<Fr>()));
  GM_CUDA(cudaMemcpyAsync(out, res.p, half * 32, cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
  return GM_OK;
}

size_t gm_fr_fold_chain_len(size_t n, size_t k) {
  size_t tot = 0;
  for (size_t j = 0; j < k; j++) { n = (n + 1) / 2; tot += n; }
  return tot;
}

int gm_fr_fold_chain(gm_ctx* ctx, const uint64_t* f, size_t n, const uint64_t* challenges, size_t k, uint64_t* out_levels) {
  GM_ARG(ctx && ((f && out_levels) || n == 0 || k == 0) && (challenges || k == 0), "NULL argument");
  GM_ENTER(ctx);
  if (n == 0 || k == 0) return GM_OK;
  const size_t tot = gm_fr_fold_chain_len(n, k);
  DevBuf& in = ctx->msm.scalars;
  DevBuf& res = ctx->msm.bases_tmp;
  GM_TRY(in.reserve(n * 32));
  GM_TRY(res.reserve(tot * 32));
  GM_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  GM_CUDA(cudaMemcpyAsync(in.p, f, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  const Fr* src = in.as<Fr>();
  Fr* dst = res.as<Fr>();
  size_t len = n;
  for (size_t j = 0; j < k; j++) {
    Fr rr;
    fr_from_u64(rr, challenges + 4 * j);
    GM_TRY(fr_fold_dev(ctx, src, len, rr, dst));
    src = dst;
    len = (len + 1) / 2;
    dst += len;
  }
  GM_CUDA(cudaMemcpyAsync(out_levels, res.p, tot * 32, cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev[0], ctx->ev[1]);
  return GM_OK;
}

// ---- sumcheck ----------------------------------------------------------------------------
// A prover handle owns its stream, events, vectors and pinned message slot: its calls never take the context lock, so
// several provers can be driven from different host threads at once (Sumcheck::prove_batch, proof.rs:85).
static Lane lane_of(gm_sumcheck* p) { return Lane{p->stream, &p->ctx->launches}; }

static int sumcheck_alloc(gm_ctx* ctx, size_t f_len, size_t g_len, const uint64_t twist[4], int flavour, gm_sumcheck** out) {
  GM_ARG(flavour == GM_SUMCHECK_GEMINI_TIME || flavour == GM_SUMCHECK_HERRING_F || flavour == GM_SUMCHECK_GEMINI_SPACE, "unknown flavour");
  gm_sumcheck* p = new (std::nothrow) gm_sumcheck();
  if (!p) return GM_ERR_OOM;
  p->ctx = ctx;
  ctx_retain(ctx);
  p->nf = f_len;
  p->ng = g_len;
  p->flavour = flavour;
  fr_from_u64(p->twist, twist);
  // time_prover.rs:35-38 (max) vs herring/time_prover.rs:36-39 and space_prover.rs:76-79 (min)
  p->tot_rounds = flavour == GM_SUMCHECK_GEMINI_TIME ? ceil_log2(std::max(f_len, g_len)) : ceil_log2(std::min(f_len, g_len));
  const size_t ctas = sc_max_ctas(f_len, g_len);
  cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  for (int k = 0; k < 2 && e == cudaSuccess; k++) {
    e = cudaEventCreate(&p->ev[k]);
    if (e == cudaSuccess) e = cudaEventCreate(&p->tm[k]);
  }
  auto alloc = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMallocAsync(ptr, std::max<size_t>(bytes, 32), p->stream); };
  alloc((void**)&p->f[0], f_len * 32);
  alloc((void**)&p->f[1], ((f_len + 1) / 2) * 32);
  alloc((void**)&p->g[0], g_len * 32);
  alloc((void**)&p->g[1], ((g_len + 1) / 2) * 32);
  alloc((void**)&p->d_partials, ctas * 64);
  alloc((void**)&p->d_ticket, 16);
  alloc((void**)&p->d_out, 64);
  if (e == cudaSuccess) e = cudaMemsetAsync(p->d_ticket, 0, 16, p->stream);
  if (e == cudaSuccess) {
    std::lock_guard<std::recursive_mutex> guard(ctx->mu);
    if (!ctx->free_slots.empty()) {
      p->slot = (int)ctx->free_slots.back();
      ctx->free_slots.pop_back();
      p->mbox = reinterpret_cast<ScMailbox*>(reinterpret_cast<uint8_t*>(ctx->pinned) + 4096 + sizeof(ScMailbox) * (size_t)p->slot);
    }
  }
  if (e == cudaSuccess && p->slot < 0) e = cudaHostAlloc((void**)&p->mbox, sizeof(ScMailbox), cudaHostAllocDefault);
  if (e == cudaSuccess) { memset(p->mbox, 0, sizeof(ScMailbox)); p->h_out = p->mbox->scratch; }
  if (e != cudaSuccess) {
    set_error("sumcheck alloc: %s", cudaGetErrorString(e));
    gm_sumcheck_free(p);
    return e == cudaErrorMemoryAllocation ? GM_ERR_OOM : GM_ERR_CUDA;
  }
  *out = p;
  return GM_OK;
}

static int sumcheck_load(gm_ctx* ctx, const void* f, size_t f_len, const void* g, size_t g_len, const uint64_t twist[4], int flavour,
                         cudaMemcpyKind kind, gm_sumcheck** out, bool big_endian = false) {
  gm_sumcheck* p = nullptr;
  {
    // the only part that touches the context: device selection, the closed check and - for device-resident inputs -
    // an event that orders this prover's stream after the work already queued on the context's stream
    GM_ENTER(ctx);
    GM_TRY(sumcheck_alloc(ctx, f_len, g_len, twist, flavour, &p));
    cudaError_t e = cudaSuccess;
    if (kind == cudaMemcpyDeviceToDevice) {
      e = cudaEventRecord(ctx->ev_join, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(p->stream, ctx->ev_join, 0);
    }
    if (e != cudaSuccess) {
      set_error("sumcheck join: %s", cudaGetErrorString(e));
      gm_sumcheck_free(p);
      return GM_ERR_CUDA;
    }
  }
  cudaError_t e = cudaSuccess;
  if (f_len) e = cudaMemcpyAsync(p->f[0], f, f_len * 32, kind, p->stream);
  if (e == cudaSuccess && g_len) e = cudaMemcpyAsync(p->g[0], g, g_len * 32, kind, p->stream);
  if (e == cudaSuccess && big_endian) {
    // streams arrive highest-degree first (space_prover.rs:38-58): the prover's own copies are reversed in place
    const Lane ln = lane_of(p);
    int rc = fr_reverse_dev(ln, ctx->sm_count, p->f[0], f_len, p->f[0]);
    if (rc == GM_OK) rc = fr_reverse_dev(ln, ctx->sm_count, p->g[0], g_len, p->g[0]);
    if (rc != GM_OK) { gm_sumcheck_free(p); return rc; }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);   // inputs are copied at construction: no aliasing afterwards
  if (e != cudaSuccess) {
    set_error("sumcheck copy: %s", cudaGetErrorString(e));
    gm_sumcheck_free(p);
    return GM_ERR_CUDA;
  }
  *out = p;
  return GM_OK;
}

int gm_sumcheck_new(gm_ctx* ctx, const uint64_t* f, size_t f_len, const uint64_t* g, size_t g_len, const uint64_t twist[4],
                    int flavour, gm_sumcheck** out) {
  GM_ARG(ctx && out && twist && (f || f_len == 0) && (g || g_len == 0), "NULL argument");
  return sumcheck_load(ctx, f, f_len, g, g_len, twist, flavour, cudaMemcpyHostToDevice, out);
}

int gm_sumcheck_new_dev(gm_ctx* ctx, const void* f_dev, size_t f_len, const void* g_dev, size_t g_len, const uint64_t twist[4],
                        int flavour, gm_sumcheck** out) {
  GM_ARG(ctx && out && twist && (f_dev || f_len == 0) && (g_dev || g_len == 0), "NULL argument");
  return sumcheck_load(ctx, f_dev, f_len, g_dev, g_len, twist, flavour, cudaMemcpyDeviceToDevice, out);
}

int gm_sumcheck_new_ex(gm_ctx* ctx, const void* f, size_t f_len, const void* g, size_t g_len, const uint64_t twist[4], int flavour,
                       int input_flags, gm_sumcheck** out) {
  GM_ARG(ctx && out && twist && (f || f_len == 0) && (g || g_len == 0), "NULL argument");
  GM_ARG((input_flags & ~(GM_INPUT_DEVICE | GM_INPUT_BIG_ENDIAN)) == 0, "unknown input flag");
  return sumcheck_load(ctx, f, f_len, g, g_len, twist, flavour, (input_flags & GM_INPUT_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                       out, (input_flags & GM_INPUT_BIG_ENDIAN) != 0);
}

// From<&SpaceProver> for TimeProver (space_prover.rs:269-307) / ElasticProver::fold (elastic_prover.rs:44-57): the folded
// vectors are already resident, only the flavour (final_foldings semantics) changes; round counters are kept
int gm_sumcheck_set_flavour(gm_sumcheck* p, int flavour) {
  GM_ARG(p, "NULL argument");
  GM_ARG(flavour == GM_SUMCHECK_GEMINI_TIME || flavour == GM_SUMCHECK_HERRING_F || flavour == GM_SUMCHECK_GEMINI_SPACE, "unknown flavour");
  p->flavour = flavour;
  return GM_OK;
}

static bool sc_use_twist(const gm_sumcheck* p, const Fr& tw) {
  return p->flavour != GM_SUMCHECK_HERRING_F && tw != Fr::one();
}

int gm_sumcheck_fold(gm_sumcheck* p, const uint64_t r[4]) {
  GM_ARG(p && r, "NULL argument");
  GM_CUDA(cudaSetDevice(p->ctx->device));
  Fr rg;
  fr_from_u64(rg, r);
  const Fr rf = rg * p->twist;  // time_prover.rs:77 / herring/time_prover.rs:85
  const int nxt = p->cur ^ 1;
  const Lane ln = lane_of(p);
  GM_TRY(fr_fold_dev(ln, p->ctx->sm_count, p->f[p->cur], p->nf, rf, p->f[nxt]));
  GM_TRY(fr_fold_dev(ln, p->ctx->sm_count, p->g[p->cur], p->ng, rg, p->g[nxt]));
  p->cur = nxt;
  p->nf = (p->nf + 1) / 2;
  p->ng = (p->ng + 1) / 2;
  p->twist = p->twist.sqr();
  return GM_OK;
}

int gm_sumcheck_next_message(gm_sumcheck* p, const uint64_t* challenge_or_null, uint64_t out_ab[8], int* out_has_msg) {
  GM_ARG(p && out_ab && out_has_msg, "NULL argument");
  GM_CUDA(cudaSetDevice(p->ctx->device));
  if (p->round > p->tot_rounds) {  // time_prover.rs:84 assert
    set_error("next_message: more rounds than needed");
    return GM_ERR_STATE;
  }
  const bool last = p->round == p->tot_rounds;
  if (challenge_or_null && last) {
    GM_TRY(gm_sumcheck_fold(p, challenge_or_null));
  }
  if (last) {
    *out_has_msg = 0;
    return GM_OK;
  }
  const Lane ln = lane_of(p);
  const uint32_t seq = ++p->seq;       // the kernel publishes (a, b) and this number in the pinned mailbox
  if (challenge_or_null) {
    Fr rg;
    fr_from_u64(rg, challenge_or_null);
    const Fr rf = rg * p->twist;
    const Fr new_twist = p->twist.sqr();
    const int nxt = p->cur ^ 1;
    GM_TRY(sc_fold_message_dev(ln, p->f[p->cur], p->nf, p->g[p->cur], p->ng, rf, rg, p->f[nxt], p->g[nxt], new_twist,
                               sc_use_twist(p, new_twist), p->d_partials, p->d_ticket, p->d_out, p->mbox, seq));
    p->cur = nxt;
    p->nf = (p->nf + 1) / 2;
    p->ng = (p->ng + 1) / 2;
    p->twist = new_twist;
  } else {
    GM_TRY(sc_message_dev(ln, p->f[p->cur], p->nf, p->g[p->cur], p->ng, p->twist, sc_use_twist(p, p->twist), p->d_partials,
                          p->d_ticket, p->d_out, p->mbox, seq));
  }
  // no D2H copy, no stream synchronisation: the last CTA of the kernel wrote the message into pinned memory
  if (!sc_wait_message(p->stream, p->mbox, seq)) {
    GM_CUDA(cudaStreamSynchronize(p->stream));
    set_error("sumcheck round: the message never arrived");
    return GM_ERR_CUDA;
  }
  memcpy(out_ab, (const void*)p->mbox->msg, 64);
  p->round++;
  *out_has_msg = 1;
  return GM_OK;
}

size_t gm_sumcheck_rounds(const gm_sumcheck* p) { return p ? p->tot_rounds : 0; }
size_t gm_sumcheck_round(const gm_sumcheck* p) { return p ? p->round : 0; }
int gm_sumcheck_set_rounds(gm_sumcheck* p, size_t round, size_t tot_rounds) {
  GM_ARG(p, "NULL argument");
  p->round = round;
  p->tot_rounds = tot_rounds;
  return GM_OK;
}
int gm_sumcheck_timer_start(gm_sumcheck* p) {
  GM_ARG(p, "NULL argument");
  GM_CUDA(cudaSetDevice(p->ctx->device));
  GM_CUDA(cudaEventRecord(p->tm[0], p->stream));
  return GM_OK;
}
int gm_sumcheck_timer_stop(gm_sumcheck* p, float* out_ms) {
  GM_ARG(p && out_ms, "NULL argument");
  GM_CUDA(cudaSetDevice(p->ctx->device));
  GM_CUDA(cudaEventRecord(p->tm[1], p->stream));
  GM_CUDA(cudaEventSynchronize(p->tm[1]));
  GM_CUDA(cudaEventElapsedTime(out_ms, p->tm[0], p->tm[1]));
  return GM_OK;
}

int gm_sumcheck_final_foldings(gm_sumcheck* p, uint64_t out_fg[8], int* out_has) {
  GM_ARG(p && out_fg && out_has, "NULL argument");
  GM_CUDA(cudaSetDevice(p->ctx->device));
  if (p->round != p->tot_rounds) { *out_has = 0; return GM_OK; }
  if (p->nf == 0 || p->ng == 0) {
    set_error("final_foldings on an empty vector");
    return GM_ERR_STATE;
  }
  // TimeProver / herring: f[0], g[0] (time_prover.rs:135-137); SpaceProver: the HEAD of the folded big-endian streams
  // (space_prover.rs:260-266) = the last coefficient of the resident little-endian vectors
  const bool space = p->flavour == GM_SUMCHECK_GEMINI_SPACE;
  GM_CUDA(cudaMemcpyAsync(p->h_out, p->f[p->cur] + (space ? p->nf - 1 : 0), 32, cudaMemcpyDeviceToHost, p->stream));
  GM_CUDA(cudaMemcpyAsync(p->h_out + 1, p->g[p->cur] + (space ? p->ng - 1 : 0), 32, cudaMemcpyDeviceToHost, p->stream));
  GM_CUDA(cudaStreamSynchronize(p->stream));
  memcpy(out_fg, p->h_out, 64);
  *out_has = 1;
  return GM_OK;
}

int gm_sumcheck_read_state(gm_sumcheck* p, uint64_t* out_f, size_t* f_len, uint64_t* out_g, size_t* g_len, uint64_t out_twist[4]) {
  GM_ARG(p, "NULL argument");
  GM_CUDA(cudaSetDevice(p->ctx->device));
  if (f_len) *f_len = p->nf;
  if (g_len) *g_len = p->ng;
  if (out_twist) memcpy(out_twist, p->twist.v, 32);
  if (out_f && p->nf) GM_CUDA(cudaMemcpyAsync(out_f, p->f[p->cur], p->nf * 32, cudaMemcpyDeviceToHost, p->stream));
  if (out_g && p->ng) GM_CUDA(cudaMemcpyAsync(out_g, p->g[p->cur], p->ng * 32, cudaMemcpyDeviceToHost, p->stream));
  GM_CUDA(cudaStreamSynchronize(p->stream));
  return GM_OK;
}

// device pointers of the current (folded) vectors: hand-off to callers that keep working on the device
int gm_sumcheck_state_dev(gm_sumcheck* p, const void** out_f_dev, size_t* f_len, const void** out_g_dev, size_t* g_len) {
  GM_ARG(p, "NULL argument");
  GM_CUDA(cudaSetDevice(p->ctx->device));
  GM_CUDA(cudaStreamSynchronize(p->stream));
  if (out_f_dev) *out_f_dev = p->f[p->cur];
  if (out_g_dev) *out_g_dev = p->g[p->cur];
  if (f_len) *f_len = p->nf;
  if (g_len) *g_len = p->ng;
  return GM_OK;
}

// Valid before or after gm_shutdown of the context (the handle keeps it alive).
int gm_sumcheck_free(gm_sumcheck* p) {
  if (!p) return GM_OK;
  gm_ctx* ctx = p->ctx;
  cudaSetDevice(ctx->device);
  if (p->stream) {
    for (int k = 0; k < 2; k++) { if (p->f[k]) cudaFreeAsync(p->f[k], p->stream); if (p->g[k]) cudaFreeAsync(p->g[k], p->stream); }
    if (p->d_partials) cudaFreeAsync(p->d_partials, p->stream);
    if (p->d_ticket) cudaFreeAsync(p->d_ticket, p->stream);
    if (p->d_out) cudaFreeAsync(p->d_out, p->stream);
    cudaStreamSynchronize(p->stream);
    cudaStreamDestroy(p->stream);
  }
  for (int k = 0; k < 2; k++) { if (p->ev[k]) cudaEventDestroy(p->ev[k]); if (p->tm[k]) cudaEventDestroy(p->tm[k]); }
  if (p->slot >= 0) {
    std::lock_guard<std::recursive_mutex> guard(ctx->mu);
    ctx->free_slots.push_back((uint32_t)p->slot);
  } else if (p->mbox) {
    cudaFreeHost(p->mbox);
  }
  delete p;
  ctx_release(ctx);
  return GM_OK;
}

}  // extern "C"
