/* gemini_b200 - C ABI of the B200-native Gemini prover hot path.
 *
 * This header is the drop-in boundary.  The reference (arkworks-rs/gemini @ 844a85e5)
 * has no FFI of its own: its seams are Rust traits and inherent methods.  Each entry
 * point below names the reference interface it replaces (paths relative to
 * /root/reference); INTEGRATION.md shows the Rust `extern "C"` block and the wrapper
 * types that bind them to ark_ec::VariableBaseMSM / sumcheck::Prover.
 *
 * Conventions
 *  - every function returns 0 on success or a GM_ERR_* code; nothing unwinds;
 *    gm_last_error() returns a thread-local description of the last failure.
 *  - field elements are little-endian u64 limbs in MONTGOMERY form exactly as
 *    arkworks' Fp<MontBackend,N> stores them: Fr = 4 limbs (32 B), Fq = 6 limbs (48 B).
 *  - an affine G1 point is x | y (96 B).  The identity is flagged either by a byte at
 *    `inf_offset` inside each record (arkworks' repr: stride 104, flag at 96) or by
 *    x = y = 0 when inf_offset < 0.
 *  - a projective G1 result is Jacobian X | Y | Z (18 limbs, 144 B) = arkworks'
 *    short_weierstrass::Projective.  Results are returned NORMALISED (Z = 1, identity
 *    = (1,1,0)), so equal group elements are equal byte strings.
 *  - host pointers unless the name ends in _dev (device pointers of the ctx's GPU).
 *  - threads: a gm_sumcheck handle owns its CUDA stream, events, vectors and pinned message slot,
 *    so DISTINCT provers may be driven concurrently from different host threads (the reference
 *    drives them from rayon workers, src/subprotocols/sumcheck/proof.rs:85); one handle is never
 *    shared mutably.  Entry points that take a gm_ctx* (MSM, folds, Fr vector helpers, streamed
 *    MSM pushes) share the context's stream and scratch arena: they are thread-safe but
 *    serialised by a per-context lock - use one context per thread for concurrent MSMs.
 *  - lifetime: the context is reference counted; every gm_srs / gm_sumcheck / gm_msm_stream
 *    handle keeps it alive, so handles may be freed before OR after gm_shutdown (Rust Drop
 *    order is arbitrary).  After gm_shutdown, calls that need the context return GM_ERR_STATE.
 *  - there is no CPU fallback: without a CUDA device gm_init fails with GM_ERR_CUDA.
 */
#ifndef GEMINI_B200_H
#define GEMINI_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM_OK 0
#define GM_ERR_CUDA 1      /* CUDA runtime failure (see gm_last_error) */
#define GM_ERR_ARG 2       /* invalid argument */
#define GM_ERR_LENGTH 3    /* gm_msm_g1_checked: bases/scalars length mismatch (== Err(min_len)) */
#define GM_ERR_STATE 4     /* call not valid in the handle's state (e.g. next_message past the last round) */
#define GM_ERR_OOM 5

typedef struct gm_ctx gm_ctx;
typedef struct gm_srs gm_srs;
typedef struct gm_msm_stream gm_msm_stream;
typedef struct gm_sumcheck gm_sumcheck;

/* ---- context ---------------------------------------------------------------------- */
int gm_init(int device_id, gm_ctx** out_ctx);
/* drops the caller's reference (call once per gm_init); live handles stay valid and freeable */
int gm_shutdown(gm_ctx* ctx);
const char* gm_last_error(void);
int gm_abi_version(void);
/* Host-only introspection of the MSM planner (no device needed): out = {window bits c, windows W, buckets per set,
 * merged bucket set (precomputed table) 0/1, affine levels, slots per thread G and warps of the first level, buckets per
 * running-sum slice}.  Honours the same GM_MSM_* / GM_AFF_* environment overrides as the MSM itself. */
int gm_msm_describe_plan(size_t n, int with_table, int sm_count, int out[8]);
/* number of kernels this library has launched on this ctx since gm_init (bench gpu_launches) */
uint64_t gm_launch_count(const gm_ctx* ctx);
/* device elapsed ms of the last gm_msm_* / gm_sumcheck_* call, by CUDA events on the ctx stream;
 * phase 0 = whole call; for MSM calls 1 = digits + counting sort + work list, 2 = k_accumulate (the
 * dominant kernel), 3 = split combine + bucket reduction + finish */
float gm_last_device_ms(const gm_ctx* ctx, int phase);
int gm_device_synchronize(gm_ctx* ctx);
/* CUDA-event stopwatch on the ctx stream (the stream every kernel of this library is launched on)
 * and an L2 flush (writes a 256 MB scratch buffer) - measurement helpers for bench.py */
int gm_timer_start(gm_ctx* ctx);
int gm_timer_stop(gm_ctx* ctx, float* out_ms);
int gm_l2_flush(gm_ctx* ctx);

/* ---- SRS (CommitterKey::powers_of_g, src/kzg/time.rs:24-27, resident on the device) ---- */
int gm_srs_load_g1(gm_ctx* ctx, const void* points, size_t n, size_t stride_bytes, long inf_offset,
                   gm_srs** out_srs);
/* synthetic bases P_i = [first_multiple + i] * G generated on the device (bench / tests) */
int gm_srs_generate_g1(gm_ctx* ctx, size_t n, uint64_t first_multiple, gm_srs** out_srs);
/* CommitterKey::new, src/kzg/time.rs:49-72 (G1 half): powers_of_g[i] = tau^i * g for i < n - misc::powers, the
 * fixed-base MSM of ark-ec (FixedBase::get_window_table / FixedBase::msm) and normalize_batch, all on the device.
 * g_xy: affine generator (Montgomery x|y); tau: Montgomery Fr.  powers_of_g2 (G2) is outside this path. */
int gm_srs_setup_g1(gm_ctx* ctx, const uint64_t g_xy[12], const uint64_t tau[4], size_t n, gm_srs** out_srs);
/* n copies of one point: DummyStreamer(G1::generator(), n), examples/snark.rs:62-65 */
int gm_srs_fill_g1(gm_ctx* ctx, const uint64_t point_xy[12], size_t n, gm_srs** out_srs);
/* Optional, one-time (key setup, like CommitterKey::new): store 2^(c*w) * P_i for every window w next to
 * the points (levels x n x 96 B of HBM; 1.3 GB for n = 2^20, 19 GB for n = 2^24).  MSMs over this SRS
 * then use ONE shared bucket set for all windows: no per-window running sums and no serial 2^(c*w)
 * doubling tail.  expected_msm_len (0 = the SRS length) picks the window size.  Up to four further tables over the
 * first n/8, n/64, n/512 and n/4096 points (with smaller windows) serve short commitments against a long SRS - the
 * fold levels committed by tensorcheck.  Results are identical. */
int gm_srs_precompute(gm_ctx* ctx, gm_srs* srs, size_t expected_msm_len);
int gm_srs_precompute_info(const gm_srs* srs, int* out_window_bits, int* out_levels);
size_t gm_srs_len(const gm_srs* srs);
int gm_srs_read(gm_ctx* ctx, const gm_srs* srs, size_t offset, size_t n, uint64_t* out_xy /* n*12 */);
int gm_srs_free(gm_srs* srs);

/* ---- MSM: ark_ec::VariableBaseMSM (called at src/kzg/time.rs:82,129; src/kzg/space.rs:52) ---- */
/* msm_unchecked / msm_bigint: sum_{i<n} scalars[i] * srs[base_offset + i].
 * n is clamped to the bases available (min(len) truncation of msm_unchecked).
 * scalars_are_bigint = 0: Montgomery Fr (msm_unchecked); 1: canonical BigInt<4> (msm_bigint).
 * `scalars` is ordinary host memory.  Pageable memory (a Rust &[Fr]) is gathered chunk by chunk into a pinned bounce
 * buffer of the context by a few host threads, behind the kernels of the previous chunk; memory that is already pinned
 * (cudaHostAlloc / cudaHostRegister) is recognised and copied piecewise, the digits of a piece extracted while the next one
 * is on the bus.  The result is the same group element either way; a constant scalar vector (>= 2^16 terms, device entry
 * points) is computed as s * (sum of the bases). */
int gm_msm_g1(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const uint64_t* scalars, size_t n,
              int scalars_are_bigint, uint64_t out_jacobian[18]);
int gm_msm_g1_dev(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const void* scalars_dev, size_t n,
                  int scalars_are_bigint, uint64_t out_jacobian[18]);
/* VariableBaseMSM::msm: returns GM_ERR_LENGTH and *out_min_len = min(lens) on a length mismatch */
int gm_msm_g1_checked(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, size_t bases_len,
                      const uint64_t* scalars, size_t scalars_len, uint64_t out_jacobian[18],
                      size_t* out_min_len);
/* ad-hoc bases supplied with the call (herring/module.rs:100, kzg/mod.rs:163) */
int gm_msm_g1_hostbases(gm_ctx* ctx, const void* points, size_t stride_bytes, long inf_offset,
                        const uint64_t* scalars, size_t n, int scalars_are_bigint, uint64_t out_jacobian[18]);

/* ---- streamed MSM: msm_chunks (src/kzg/space.rs:22-55) and ChunkedPippenger
 *      (src/kzg/msm/stream_pippenger.rs:209-271).  The accumulator stays on the device;
 *      each push is one pipelined chunk; finalize performs the single 144-byte D2H. ---- */
int gm_msm_stream_new(gm_ctx* ctx, const gm_srs* srs_or_null, size_t chunk_cap, gm_msm_stream** out);
/* points == NULL: bases are srs[base_offset .. base_offset+m) ; else m ad-hoc points */
int gm_msm_stream_push(gm_msm_stream* s, const void* points, size_t stride_bytes, long inf_offset,
                       size_t base_offset, const uint64_t* scalars, size_t m, int scalars_are_bigint);
/* scalars resident on the device (quotients / fold levels of the elastic prover): bases = srs[base_offset .. +m) */
int gm_msm_stream_push_dev(gm_msm_stream* s, size_t base_offset, const void* scalars_dev, size_t m, int scalars_are_bigint);
int gm_msm_stream_finalize(gm_msm_stream* s, uint64_t out_jacobian[18]);
int gm_msm_stream_free(gm_msm_stream* s);

/* sum of k Jacobian points (Commitment / EvaluationProof `Add`, `Sum`: src/kzg/mod.rs:114-126) */
int gm_g1_sum(gm_ctx* ctx, const uint64_t* jacobians /* k*18 */, size_t k, uint64_t out_jacobian[18]);

/* ---- multi-GPU (the reference is single-process): one process per GPU, one gm_ctx per process, the
 *      SRS split by contiguous point range.  Rank 0 draws an id (ncclGetUniqueId) and hands the 128 bytes to the
 *      other ranks out of band (MPI, a TCP store, torch.distributed ...); every rank then calls gm_comm_init.  The
 *      exchange of an MSM is ONE ncclAllGather of the 192-byte partial accumulators queued on the library's stream
 *      between the bucket reduction and the normalisation, followed by world-1 curve additions on every rank. ---- */
#define GM_COMM_ID_BYTES 128
int gm_comm_unique_id(uint8_t out_id[GM_COMM_ID_BYTES]);
int gm_comm_init(gm_ctx* ctx, const uint8_t id[GM_COMM_ID_BYTES], int rank, int world);
int gm_comm_rank(const gm_ctx* ctx);
int gm_comm_world(const gm_ctx* ctx);
int gm_comm_nccl_version(void);
int gm_comm_barrier(gm_ctx* ctx);
/* every rank contributes `bytes` (<= 256) host bytes; recv_all receives world * bytes in rank order (the 64-byte round
 * messages and the per-rank last coefficients of the sharded sumcheck) */
int gm_comm_allgather(gm_ctx* ctx, const void* send, size_t bytes, void* recv_all);
/* msm_unchecked over ALL ranks' shards: `srs` / `scalars` are this rank's contiguous range; every rank gets the total */
int gm_msm_g1_sharded(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const uint64_t* scalars, size_t n,
                      int scalars_are_bigint, uint64_t out_jacobian[18]);
int gm_msm_g1_sharded_dev(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const void* scalars_dev, size_t n,
                          int scalars_are_bigint, uint64_t out_jacobian[18]);
/* term i uses scalars_dev[i * scalar_stride]; sharded != 0 adds the exchange.  Cyclic sharding (rank r holds every world-th
 * point from r on) splits a vector of ANY length evenly: rank r passes scalars_dev = v + r, scalar_stride = world,
 * n = ceil((len - r) / world) - the shape of a multi-GPU CommitterKey::commit (src/kzg/time.rs:81-83) */
int gm_msm_g1_strided_dev(gm_ctx* ctx, const gm_srs* srs, size_t base_offset, const void* scalars_dev, size_t n, size_t scalar_stride,
                          int scalars_are_bigint, int sharded, uint64_t out_jacobian[18]);
/* every stride-th point of a resident SRS from `first` on, as a new SRS handle (a rank's cyclic shard) */
int gm_srs_subsample(gm_ctx* ctx, const gm_srs* srs, size_t first, size_t stride, size_t count, gm_srs** out_srs);
/* msm_chunks across ranks (config 5): every rank streams its own range; the exchange happens once, here.  The handle
 * must not be pushed to afterwards. */
int gm_msm_stream_finalize_sharded(gm_msm_stream* s, uint64_t out_jacobian[18]);

/* ---- Fr folds: misc::fold_polynomial (src/misc.rs:52-56), herring split_fold
 *      (src/herring/time_prover.rs:72-76), tensorcheck::foldings_polynomial
 *      (src/subprotocols/tensorcheck/mod.rs:124-133) ---- */
/* out[i] = f[2i] + r * f[2i+1], i < ceil(n/2); a missing odd element is zero */
int gm_fr_fold(gm_ctx* ctx, const uint64_t* f, size_t n, const uint64_t r[4], uint64_t* out);
int gm_fr_fold_dev(gm_ctx* ctx, const void* f_dev, size_t n, const uint64_t r[4], void* out_dev);
/* k successive folds by challenges[0..k); level j (1-based) has ceil(n / 2^j) elements and is
 * written at out + level_offset(j) elements, levels concatenated in order 1..k */
int gm_fr_fold_chain(gm_ctx* ctx, const uint64_t* f, size_t n, const uint64_t* challenges, size_t k,
                     uint64_t* out_levels);
/* total number of Fr elements gm_fr_fold_chain writes */
size_t gm_fr_fold_chain_len(size_t n, size_t k);

/* ---- sumcheck provers: trait Prover (src/subprotocols/sumcheck/prover.rs:30-45) ---- */
#define GM_SUMCHECK_GEMINI_TIME 0 /* TimeProver, sumcheck/time_prover.rs:42-137: rounds from max len, twisted message */
#define GM_SUMCHECK_HERRING_F 1   /* herring TimeProver<FModule>, herring/time_prover.rs:44-137: rounds from min len */
#define GM_SUMCHECK_GEMINI_SPACE 2 /* SpaceProver, sumcheck/space_prover.rs:38-266: messages of TimeProver, rounds from the MIN
                                     length (:76-79), final foldings = head of the folded big-endian streams (:260-266).  The folded
                                     vectors stay resident instead of re-streaming the input every round. */
#define GM_INPUT_DEVICE 1         /* gm_sumcheck_new_ex: f, g are device pointers of the context's GPU (else host: pageable or pinned) */
#define GM_INPUT_BIG_ENDIAN 2     /* f, g arrive highest-degree coefficient first, the stream order of the space / elastic provers */
int gm_sumcheck_new(gm_ctx* ctx, const uint64_t* f, size_t f_len, const uint64_t* g, size_t g_len,
                    const uint64_t twist[4], int flavour, gm_sumcheck** out);
int gm_sumcheck_new_dev(gm_ctx* ctx, const void* f_dev, size_t f_len, const void* g_dev, size_t g_len,
                        const uint64_t twist[4], int flavour, gm_sumcheck** out);
int gm_sumcheck_new_ex(gm_ctx* ctx, const void* f, size_t f_len, const void* g, size_t g_len, const uint64_t twist[4], int flavour,
                       int input_flags, gm_sumcheck** out);
/* ElasticProver::fold's Space -> Time switch (elastic_prover.rs:44-57; From<&SpaceProver> for TimeProver, space_prover.rs:269-307) */
int gm_sumcheck_set_flavour(gm_sumcheck* p, int flavour);
/* next_message(Option<F>): challenge may be NULL (first call).  *out_has_msg = 0 <=> None. */
int gm_sumcheck_next_message(gm_sumcheck* p, const uint64_t* challenge_or_null, uint64_t out_ab[8],
                             int* out_has_msg);
int gm_sumcheck_fold(gm_sumcheck* p, const uint64_t r[4]);
size_t gm_sumcheck_rounds(const gm_sumcheck* p);
size_t gm_sumcheck_round(const gm_sumcheck* p);
/* override the round counters (From<&SpaceProver> for TimeProver, space_prover.rs:269-307) */
int gm_sumcheck_set_rounds(gm_sumcheck* p, size_t round, size_t tot_rounds);
int gm_sumcheck_final_foldings(gm_sumcheck* p, uint64_t out_fg[8], int* out_has);
/* a CUDA-event stopwatch on the handle's own stream */
int gm_sumcheck_timer_start(gm_sumcheck* p);
int gm_sumcheck_timer_stop(gm_sumcheck* p, float* out_ms);
/* device pointers of the current (folded) vectors, valid until the next call on the handle */
int gm_sumcheck_state_dev(gm_sumcheck* p, const void** out_f_dev, size_t* f_len, const void** out_g_dev, size_t* g_len);
/* copy out the current (folded) vectors - used by tests and by the elastic hand-off */
int gm_sumcheck_read_state(gm_sumcheck* p, uint64_t* out_f, size_t* f_len, uint64_t* out_g, size_t* g_len,
                           uint64_t out_twist[4]);
int gm_sumcheck_free(gm_sumcheck* p);

/* ---- Fiat-Shamir: merlin::Transcript + GeminiTranscript (src/transcript.rs:8-34), host-side, no device needed ---- */
typedef struct gm_transcript gm_transcript;
/* merlin::Transcript::new(label) */
int gm_transcript_new(const uint8_t* label, size_t label_len, gm_transcript** out);
int gm_transcript_clone(const gm_transcript* t, gm_transcript** out);
int gm_transcript_free(gm_transcript* t);
int gm_transcript_append_message(gm_transcript* t, const uint8_t* label, size_t label_len, const uint8_t* msg, size_t len);
int gm_transcript_challenge_bytes(gm_transcript* t, const uint8_t* label, size_t label_len, uint8_t* out, size_t n);
/* append_serializable of 1 or 2 Fr (Montgomery limbs in, 32-byte little-endian canonical integers hashed) */
int gm_transcript_append_fr(gm_transcript* t, const uint8_t* label, size_t label_len, const uint64_t* mont, size_t count);
/* get_challenge::<Fr>: 64 PRF bytes -> Fr::from_random_bytes with retry; Montgomery limbs out */
int gm_transcript_get_challenge_fr(gm_transcript* t, const uint8_t* label, size_t label_len, uint64_t out_mont[4]);
/* Sumcheck::prove (src/subprotocols/sumcheck/proof.rs:36-66): the whole Fiat-Shamir loop as one call.  out_msgs receives
 * rounds x (a | b) and out_challenges rounds x Fr (Montgomery limbs; `capacity` rounds of room, gm_sumcheck_rounds(p) is
 * enough), out_final the final foldings, which are also appended to the transcript (b"final-folding") as the reference does. */
int gm_sumcheck_prove(gm_sumcheck* p, gm_transcript* t, uint64_t* out_msgs, uint64_t* out_challenges, size_t capacity, size_t* out_rounds,
                      uint64_t out_final[8]);

/* ---- raw device buffers for callers that keep vectors resident (bench, pipelines) ---- */
int gm_dev_alloc(gm_ctx* ctx, size_t bytes, void** out_dev);
int gm_dev_free(gm_ctx* ctx, void* dev);
int gm_dev_upload(gm_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int gm_dev_download(gm_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
/* fill a device buffer with n pseudo-random canonical-range Fr elements in Montgomery form
 * (splitmix64 counter stream; bench / full-size property tests) */
int gm_fr_random_dev(gm_ctx* ctx, void* out_dev, size_t n, uint64_t seed);

/* ---- Fr vectors of the time prover, device resident.  All pointers are device
 *      pointers of the ctx's GPU; calls are queued on the ctx stream and return without synchronising
 *      unless they hand a scalar back to the host. ---- */
int gm_dev_memset(gm_ctx* ctx, void* dev, int byte, size_t bytes);
int gm_dev_copy(gm_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);
/* out[i] = in[n-1-i] (in place when out == in): big-endian stream order <-> resident little-endian order */
int gm_fr_reverse_dev(gm_ctx* ctx, const void* in_dev, size_t n, void* out_dev);
/* out[i] = x^i, i < n: misc::powers (src/misc.rs:59-65) */
int gm_fr_powers_dev(gm_ctx* ctx, const uint64_t x[4], size_t n, void* out_dev);
/* out = (E, O) = (sum_{i even} f_i x^i, sum_{i odd} f_i x^i): f(x) = E + O and f(-x) = E - O;
 * misc::evaluate_le (src/misc.rs:194-199) at beta and -beta in one pass */
int gm_fr_eval_dev(gm_ctx* ctx, const void* f_dev, size_t n, const uint64_t x[4], uint64_t out_even_odd[8]);
/* out[idx] = prod_{bit j of idx} rho[j], 2^k elements: misc::tensor (src/misc.rs:133-149) */
int gm_fr_tensor_dev(gm_ctx* ctx, const uint64_t* rho, size_t k, void* out_dev);
/* out[i] = a[i] * b[i]: misc::hadamard (src/misc.rs:205-208) */
int gm_fr_hadamard_dev(gm_ctx* ctx, const void* a_dev, const void* b_dev, size_t n, void* out_dev);
/* acc[i] += c * x[i], i < n: one term of misc::linear_combination (src/misc.rs:37-48) */
int gm_fr_axpy_dev(gm_ctx* ctx, void* acc_dev, const void* x_dev, size_t n, const uint64_t c[4]);
/* y = M x for a CSR matrix (u32 rowptr[nrows+1], u32 col[], Fr vals[]): misc::product_matrix_vector
 * (src/misc.rs:100-110); with the transposed matrices, the abc_tensored sums of snark/time_prover.rs:63-81 */
int gm_fr_spmv_dev(gm_ctx* ctx, const void* rowptr_dev, const void* col_dev, const void* vals_dev, size_t nrows,
                   const void* x_dev, void* y_dev);
/* f = q * (X - a) + rem: q (n-1 coefficients, little-endian) and rem; the quotients of CommitterKey::open /
 * open_multi_points (src/kzg/time.rs:112-145) as a parallel suffix-Horner scan */
int gm_fr_div_linear_dev(gm_ctx* ctx, const void* f_dev, size_t n, const uint64_t a[4], void* q_dev, uint64_t out_rem[4]);
/* gm_fr_fold_chain with input and output resident on the device */
int gm_fr_fold_chain_dev(gm_ctx* ctx, const void* f_dev, size_t n, const uint64_t* challenges, size_t k, void* out_levels_dev);

/* ---- self-test kernels (parity tests of the device field / curve arithmetic) ---- */
/* op: 0 mul, 1 add, 2 sub, 3 inv (binary almost-inverse), 4 from_mont, 5 to_mont, 6 sqr, 7 inv (division steps: the one the
 * MSM kernels use);  field: 0 = Fq (12 u32), 1 = Fr (8 u32) */
int gm_selftest_field(gm_ctx* ctx, int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* r, size_t n);
/* op: 0 xyzz += affine, 1 xyzz += -affine, 2 xyzz += xyzz, 3 xyzz = 2*xyzz; out = normalised Jacobian (36 u32) */
int gm_selftest_curve(gm_ctx* ctx, int op, const uint32_t* acc_xyzz, const uint32_t* other, uint32_t* out_jac, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* GEMINI_B200_H */
