// Fold-in-half reductions over Fr (BLS12-381 scalar field) for sm_100a.
//
//   k_fr_fold           misc::fold_polynomial (/root/reference/src/misc.rs:52-56), herring split_fold
//                       (/root/reference/src/herring/time_prover.rs:72-76)             [48 B / input element]
//   k_sc_message        first round message of TimeProver::next_message
//                       (/root/reference/src/subprotocols/sumcheck/time_prover.rs:98-118)
//   k_sc_fold_message   fold of f and g (time_prover.rs:75-80) FUSED with the next round's message:
//                       the folded vectors are produced, written once and consumed from registers
//                       (the CPU reference makes two passes and allocates two fresh Vecs per round)
//
// Message, for pair i of the (folded) vectors with t_i = twist^(2i):
//   gemini  : a += f[2i] g[2i] t_i ; b += (f[2i] g[2i+1] + g[2i] f[2i+1] twist) t_i
//   herring : the same with twist = 1 (herring/time_prover.rs:91-123 applies the twist only in fold)
// Missing elements (odd tails, unequal lengths) read as zero, exactly like the reference's
// `unwrap_or(&zero)` / zip-to-shorter (a product with a zero partner vanishes).
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "fp.cuh"
#include "fr.cuh"

namespace gm {

static constexpr int SC_THREADS = 256;
static constexpr int SC_K = 8;                       // pairs per thread
static constexpr int SC_TILE = SC_THREADS * SC_K;    // pairs per CTA

__device__ __forceinline__ Fr load_fr(const Fr* p) {
  Fr r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4 lo = s[0], hi = s[1];
  r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
  r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
  return r;
}
__device__ __forceinline__ Fr load_fr_or_zero(const Fr* base, size_t idx, size_t n) {
  return idx < n ? load_fr(base + idx) : Fr::zero();
}
__device__ __forceinline__ void store_fr(Fr* p, const Fr& v) {
  uint4* d = reinterpret_cast<uint4*>(p);
  d[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
  d[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fr_fold(const Fr* __restrict__ f, size_t n, Fr r, Fr* __restrict__ out) {
  const size_t half = (n + 1) / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    Fr e = load_fr(f + 2 * i);
    if (2 * i + 1 < n) {
      Fr o = load_fr(f + 2 * i + 1);
      e = e + r * o;
    }
    store_fr(out + i, e);
  }
}

// ---------------------------------------------------------------------------------------------
struct PowTable {
  Fr p[40];  // p[k] = (twist^2)^(2^k)
};

__device__ __forceinline__ Fr pow_from_table(const PowTable& tab, uint64_t e) {
  Fr r = Fr::one();
  for (int k = 0; k < 40 && (e >> k); k++)
    if ((e >> k) & 1ull) r = r * tab.p[k];
  return r;
}

__device__ __forceinline__ Fr warp_sum_fr(Fr v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Fr other;
#pragma unroll
    for (int j = 0; j < 8; j++) other.v[j] = __shfl_down_sync(0xffffffffu, v.v[j], o);
    v = v + other;
  }
  return v;
}

// (a, b) leave the device: result slot in HBM and, when the prover has one, its pinned mailbox (msg, then msg_seq = seq),
// where the host is spinning for it - no D2H copy and no stream synchronisation on the round's critical path (measured at
// 2^24: 2.37 -> 1.98 ms for the 24 rounds).  A persistent single-CTA kernel for the last rounds, fed challenges through the
// same mailbox, was measured on top of this and did not pay (2.03 ms: one CTA is slower than a small grid, and the
// launch it saves is all that was left): removed.
__device__ __forceinline__ void sc_publish(const Fr& x, const Fr& y, Fr* out, ScMailbox* mb, uint32_t seq) {
  store_fr(out, x);
  store_fr(out + 1, y);
  if (mb != nullptr) {
    volatile uint32_t* dst = reinterpret_cast<volatile uint32_t*>(&mb->msg[0]);
#pragma unroll
    for (int j = 0; j < 8; j++) { dst[j] = x.v[j]; dst[8 + j] = y.v[j]; }
    __threadfence_system();
    mb->msg_seq = seq;
  }
}

// CTA sum of (a, b); then the last CTA to arrive sums all CTA partials and publishes the message.
__device__ __forceinline__ void sc_reduce_and_publish(Fr a, Fr b, Fr* partials, unsigned int* ticket, Fr* out, ScMailbox* mb = nullptr,
                                                      uint32_t seq = 0) {
  __shared__ Fr sh[2 * (SC_THREADS / 32)];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  a = warp_sum_fr(a);
  b = warp_sum_fr(b);
  if (lane == 0) { sh[2 * wid] = a; sh[2 * wid + 1] = b; }
  __syncthreads();
  if (wid == 0) {
    Fr x = lane < SC_THREADS / 32 ? sh[2 * lane] : Fr::zero();
    Fr y = lane < SC_THREADS / 32 ? sh[2 * lane + 1] : Fr::zero();
    x = warp_sum_fr(x);
    y = warp_sum_fr(y);
    if (lane == 0) {
      if (gridDim.x == 1) {
        // the last rounds of every sumcheck run on one CTA: its sum IS the message - no partials, no fence, no ticket
        // (they are half of the dependency chain of a round that short)
        sc_publish(x, y, out, mb, seq);
      } else {
        store_fr(partials + 2 * blockIdx.x, x);
        store_fr(partials + 2 * blockIdx.x + 1, y);
        __threadfence();
        unsigned int tk = atomicAdd(ticket, 1u);
        is_last = (tk == gridDim.x - 1);
      }
    }
  }
  if (gridDim.x == 1) return;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  Fr x = Fr::zero(), y = Fr::zero();
  for (unsigned int k = threadIdx.x; k < gridDim.x; k += blockDim.x) {
    x = x + load_fr(partials + 2 * k);
    y = y + load_fr(partials + 2 * k + 1);
  }
  x = warp_sum_fr(x);
  y = warp_sum_fr(y);
  __syncthreads();
  if (lane == 0) { sh[2 * wid] = x; sh[2 * wid + 1] = y; }
  __syncthreads();
  if (wid == 0) {
    x = lane < SC_THREADS / 32 ? sh[2 * lane] : Fr::zero();
    y = lane < SC_THREADS / 32 ? sh[2 * lane + 1] : Fr::zero();
    x = warp_sum_fr(x);
    y = warp_sum_fr(y);
    if (lane == 0) {
      *ticket = 0;  // re-arm for the next round
      sc_publish(x, y, out, mb, seq);
    }
  }
}

// The sums of the round messages are accumulated with LAZY reduction (FrAcc, fp.cuh - like ark-ff's sum_of_products
// behind misc::ip_unsafe): the last product of every term keeps its full 512 bits (64 IMAD.WIDE instead of the 128 of
// a Montgomery product) and the Montgomery reduction is paid once per thread, not once per term.
//   twist == 1:  a += fe*ge ; b += fe*go + ge*fo                          3 lazy products per pair
//   twisted:     u = fe*t, v = ge*(t*twist) ; a += u*ge ; b += u*go + v*fo   2 full + 3 lazy (+ 2 to advance t, t*twist)
using ScAcc = FrAcc;
__device__ __forceinline__ Fr sc_acc_value(const ScAcc& a) { return a.reduce(); }
template <bool TW>
__device__ __forceinline__ void pair_contrib(ScAcc& a, ScAcc& b, const Fr& fe, const Fr& fo, const Fr& ge, const Fr& go,
                                             const Fr& t, const Fr& tt) {
  static_assert(ScAcc::UNREDUCED_RUN >= 2, "b takes two unreduced products per pair");
  if (TW) {
    const Fr u = fe * t;
    const Fr v = ge * tt;
    a.mul_add_unreduced(u, ge);
    b.mul_add_unreduced(u, go);
    b.mul_add_unreduced(v, fo);
  } else {
    a.mul_add_unreduced(fe, ge);
    b.mul_add_unreduced(fe, go);
    b.mul_add_unreduced(ge, fo);
  }
  a.normalize();
  b.normalize();
}

template <bool TW>
__global__ void __launch_bounds__(SC_THREADS, 2)
k_sc_message(const Fr* __restrict__ f, size_t nf, const Fr* __restrict__ g, size_t ng, Fr twist, PowTable tab, int kpt,
             Fr* partials, unsigned int* ticket, Fr* out, ScMailbox* mb, uint32_t seq) {
  const size_t npairs = min((nf + 1) / 2, (ng + 1) / 2);
  const size_t i0 = (size_t)blockIdx.x * SC_THREADS * kpt + threadIdx.x;
  ScAcc a = ScAcc::zero(), b = ScAcc::zero();
  Fr t = Fr::one(), tt = Fr::one(), step = Fr::one();
  if (TW && i0 < npairs) { t = pow_from_table(tab, i0); tt = t * twist; step = tab.p[8]; }  // step = (twist^2)^SC_THREADS
#pragma unroll 1
  for (int k = 0; k < kpt; k++) {
    const size_t i = i0 + (size_t)k * SC_THREADS;
    if (i >= npairs) break;
    Fr fe = load_fr_or_zero(f, 2 * i, nf), fo = load_fr_or_zero(f, 2 * i + 1, nf);
    Fr ge = load_fr_or_zero(g, 2 * i, ng), go = load_fr_or_zero(g, 2 * i + 1, ng);
    pair_contrib<TW>(a, b, fe, fo, ge, go, t, tt);
    if (TW) { t = t * step; tt = tt * step; }
  }
  sc_reduce_and_publish(sc_acc_value(a), sc_acc_value(b), partials, ticket, out, mb, seq);
}

// fold by (rf, rg) and message of the folded vectors with the squared twist `twist` (already squared
// by the host) in one pass.  nf/ng are the lengths BEFORE the fold.
template <bool TW>
__global__ void __launch_bounds__(SC_THREADS, 2)   // 2 CTAs (16 warps) per SM: at most 128 registers
k_sc_fold_message(const Fr* __restrict__ f, size_t nf, const Fr* __restrict__ g, size_t ng, Fr rf, Fr rg,
                  Fr* __restrict__ f_out, Fr* __restrict__ g_out, Fr twist, PowTable tab, int kpt,
                  Fr* partials, unsigned int* ticket, Fr* out, ScMailbox* mb, uint32_t seq) {
  const size_t nf2 = (nf + 1) / 2, ng2 = (ng + 1) / 2;
  const size_t npairs = max((nf2 + 1) / 2, (ng2 + 1) / 2);  // every element must be folded
  const size_t i0 = (size_t)blockIdx.x * SC_THREADS * kpt + threadIdx.x;
  ScAcc a = ScAcc::zero(), b = ScAcc::zero();
  Fr t = Fr::one(), tt = Fr::one(), step = Fr::one();
  if (TW && i0 < npairs) { t = pow_from_table(tab, i0); tt = t * twist; step = tab.p[8]; }
#pragma unroll 1
  for (int k = 0; k < kpt; k++) {
    const size_t i = i0 + (size_t)k * SC_THREADS;
    if (i >= npairs) break;
    Fr fe = load_fr_or_zero(f, 4 * i, nf);
    if (4 * i + 1 < nf) fe = fe + rf * load_fr(f + 4 * i + 1);
    Fr fo = load_fr_or_zero(f, 4 * i + 2, nf);
    if (4 * i + 3 < nf) fo = fo + rf * load_fr(f + 4 * i + 3);
    Fr ge = load_fr_or_zero(g, 4 * i, ng);
    if (4 * i + 1 < ng) ge = ge + rg * load_fr(g + 4 * i + 1);
    Fr go = load_fr_or_zero(g, 4 * i + 2, ng);
    if (4 * i + 3 < ng) go = go + rg * load_fr(g + 4 * i + 3);
    if (2 * i < nf2) store_fr(f_out + 2 * i, fe);
    if (2 * i + 1 < nf2) store_fr(f_out + 2 * i + 1, fo);
    if (2 * i < ng2) store_fr(g_out + 2 * i, ge);
    if (2 * i + 1 < ng2) store_fr(g_out + 2 * i + 1, go);
    pair_contrib<TW>(a, b, fe, fo, ge, go, t, tt);
    if (TW) { t = t * step; tt = tt * step; }
  }
  sc_reduce_and_publish(sc_acc_value(a), sc_acc_value(b), partials, ticket, out, mb, seq);
}

// ---------------------------------------------------------------------------------------------
// Staged variants of the two round kernels for long vectors.  ncu (profiles/r02_ncu_sumcheck_summary.txt) shows the
// register-fed kernels above at 63-73 % of the multiplier pipe with 16 warps per SM: whenever a warp waits for its
// loads (long scoreboard, 29 % of the stall cycles) the three others cannot cover the dependent IMAD.WIDE chains.
// Here every WARP streams its own slices of f and g through a private ring of shared-memory buffers with
// cp.async (LDGSTS, 16 bytes per lane, 512 contiguous bytes per instruction): the copies of the next two half-steps
// are in flight while the current one is multiplied, they cost no registers, and because a buffer is only ever read
// by the warp that filled it the only synchronisation is cp.async.wait_group + __syncwarp.  Elements beyond the end
// of a vector are zero-filled by the copy (src-size 0), which is the reference's `unwrap_or(zero)` / zip-to-shorter.
//   EPP = elements of one vector per "pair": 2 (message) or 4 (fold + message); a half-step = 32 * EPP elements.
//   Layout of a buffer: thread t's 2*EPP chunks of 16 bytes are contiguous, chunk index XOR-swizzled by t so that
//   the LDS.128 of a quarter warp touch all 32 banks.
// ---------------------------------------------------------------------------------------------
template <int EPP>
struct ScStage {
  static constexpr int C = 2 * EPP;                       // 16-byte chunks per thread and half-step
  static constexpr uint32_t BYTES = 32u * C * 16u;        // one half-step of one warp
  __device__ __forceinline__ static uint32_t slot(uint32_t t, uint32_t c) {
    const uint32_t sw = (C == 8) ? (t & 7u) : ((t >> 1) & 3u);
    return (t * C + (c ^ sw)) * 16u;
  }
};

// the warp copies elements [e0, e0 + 32 * EPP) of v (n elements; zero beyond) into the buffer at shared address `buf`
template <int EPP>
__device__ __forceinline__ void sc_stage_issue(uint32_t buf, const Fr* __restrict__ v, size_t n, size_t e0, uint32_t lane) {
  using St = ScStage<EPP>;
#pragma unroll
  for (int j = 0; j < St::C; j++) {
    const uint32_t q = (uint32_t)j * 32u + lane;          // chunk of the slice: lanes copy 512 contiguous bytes
    const size_t e = e0 + (q >> 1);
    const bool ok = e < n;
    const char* src = reinterpret_cast<const char*>(v + (ok ? e : 0)) + (q & 1u) * 16u;
    const uint32_t dst = buf + St::slot(q / St::C, q % St::C);
    const uint32_t sz = ok ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
  }
}
__device__ __forceinline__ void sc_stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void sc_stage_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
  __syncwarp();
}
// element k (0 <= k < EPP) of this lane's part of a landed half-step
template <int EPP>
__device__ __forceinline__ Fr sc_stage_read(uint32_t buf, uint32_t lane, int k) {
  using St = ScStage<EPP>;
  Fr r;
  uint4 lo, hi;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "r"(buf + St::slot(lane, 2 * k)));
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "r"(buf + St::slot(lane, 2 * k + 1)));
  r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
  r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
  return r;
}

static constexpr int SC_RING = 3;   // half-step buffers per warp: two copies in flight while one is consumed
// Measured and rejected (2^24, round 2): three CTAs per SM (24 warps at 80 registers, 76-560 bytes spilled, a ring of two
// in the fold kernel so that 3 x 64 KB fit): 1.87 ms against 1.72 ms for the whole sumcheck.

// FOLD = false: message of (f, g).  FOLD = true: fold by (rf, rg), write the folded vectors, message of the folded
// vectors (nf / ng are the lengths BEFORE the fold, npairs counts pairs of the vectors the message is taken of).
template <bool TW, bool FOLD>
__global__ void __launch_bounds__(SC_THREADS, 2)
k_sc_staged(const Fr* __restrict__ f, size_t nf, const Fr* __restrict__ g, size_t ng, Fr rf, Fr rg, Fr* __restrict__ f_out,
            Fr* __restrict__ g_out, size_t npairs, Fr twist, PowTable tab, Fr step, Fr* partials, unsigned int* ticket, Fr* out,
            ScMailbox* mb, uint32_t seq) {
  constexpr int EPP = FOLD ? 4 : 2;
  using St = ScStage<EPP>;
  extern __shared__ uint4 sc_ring_raw[];
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(sc_ring_raw) + wid * (SC_RING * St::BYTES);
  // PERSISTENT grid (two CTAs per SM): CTA b takes the tiles b, b + gridDim.x, ... of SC_THREADS pairs, so the two
  // Montgomery reductions, the shuffle trees and the ticket of the epilogue are paid once per CTA and not once per
  // 8 pairs (they were a quarter of the instructions of the 8-pairs-per-thread grid), and the tiles divide evenly.
  const size_t stride = (size_t)gridDim.x * SC_THREADS;                    // pairs between two tiles of this CTA
  const size_t i0 = (size_t)blockIdx.x * SC_THREADS + threadIdx.x;         // this thread's first pair
  const size_t w0 = i0 - lane;                                             // the warp's first pair
  // iterations of this warp (uniform across its lanes): pairs w0 + k * stride < npairs
  int iters = 0;
  if (w0 < npairs) iters = (int)((npairs - w0 + stride - 1) / stride);
  const int hsteps = 2 * iters;                                            // f slice, g slice, f slice, ...
  auto issue = [&](int h) {
    if (h < hsteps) {
      const size_t e0 = (w0 + (size_t)(h >> 1) * stride) * EPP;
      if (h & 1) sc_stage_issue<EPP>(ring + (uint32_t)(h % SC_RING) * St::BYTES, g, ng, e0, lane);
      else sc_stage_issue<EPP>(ring + (uint32_t)(h % SC_RING) * St::BYTES, f, nf, e0, lane);
    }
    sc_stage_commit();   // (possibly empty) group: keeps the wait_group distance constant
  };
  issue(0);
  issue(1);
  ScAcc a = ScAcc::zero(), b = ScAcc::zero();
  Fr t = Fr::one(), tt = Fr::one();
  if (TW && iters > 0) { t = pow_from_table(tab, i0); tt = t * twist; }    // step = (twist^2)^stride, from the host
  const size_t nf2 = (nf + 1) / 2, ng2 = (ng + 1) / 2;
#pragma unroll 1
  for (int k = 0; k < iters; k++) {
    const size_t i = i0 + (size_t)k * stride;
    Fr fe, fo, ge, go;
    // ---- f slice
    sc_stage_wait<1>();
    issue(2 * k + 2);
    {
      const uint32_t buf = ring + (uint32_t)((2 * k) % SC_RING) * St::BYTES;
      if (FOLD) {
        fe = sc_stage_read<EPP>(buf, lane, 0) + rf * sc_stage_read<EPP>(buf, lane, 1);
        fo = sc_stage_read<EPP>(buf, lane, 2) + rf * sc_stage_read<EPP>(buf, lane, 3);
        if (2 * i < nf2) store_fr(f_out + 2 * i, fe);
        if (2 * i + 1 < nf2) store_fr(f_out + 2 * i + 1, fo);
      } else {
        fe = sc_stage_read<EPP>(buf, lane, 0);
        fo = sc_stage_read<EPP>(buf, lane, 1);
      }
    }
    // ---- g slice
    sc_stage_wait<1>();
    issue(2 * k + 3);
    {
      const uint32_t buf = ring + (uint32_t)((2 * k + 1) % SC_RING) * St::BYTES;
      if (FOLD) {
        ge = sc_stage_read<EPP>(buf, lane, 0) + rg * sc_stage_read<EPP>(buf, lane, 1);
        go = sc_stage_read<EPP>(buf, lane, 2) + rg * sc_stage_read<EPP>(buf, lane, 3);
        if (2 * i < ng2) store_fr(g_out + 2 * i, ge);
        if (2 * i + 1 < ng2) store_fr(g_out + 2 * i + 1, go);
      } else {
        ge = sc_stage_read<EPP>(buf, lane, 0);
        go = sc_stage_read<EPP>(buf, lane, 1);
      }
    }
    pair_contrib<TW>(a, b, fe, fo, ge, go, t, tt);
    if (TW) { t = t * step; tt = tt * step; }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  sc_reduce_and_publish(sc_acc_value(a), sc_acc_value(b), partials, ticket, out, mb, seq);
}

// out[i] = in[n - 1 - i]: the big-endian streams of the elastic prover (Reverse(..) in /root/reference/src/kzg/space.rs:288-297,
// src/snark/elastic_prover.rs:174-188) are the resident little-endian vectors read backwards.  In place when out == in.
__global__ void __launch_bounds__(256)
k_fr_reverse(const Fr* in, size_t n, Fr* out) {
  const size_t half = (n + 1) / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
    const size_t j = n - 1 - i;
    const Fr a = load_fr(in + i), b = load_fr(in + j);
    store_fr(out + i, b);
    store_fr(out + j, a);
  }
}

// splitmix64 counter stream -> Fr elements (the 255-bit value, minus r if needed, is used directly as
// the Montgomery representative).  tests/util.py holds the numpy restatement.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void k_fr_random(Fr* __restrict__ out, size_t n, uint64_t seed) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr v;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint64_t w = splitmix64(seed + 4 * i + j);
      v.v[2 * j] = (uint32_t)w;
      v.v[2 * j + 1] = (uint32_t)(w >> 32);
    }
    v.v[7] &= 0x7FFFFFFFu;
    detail::cond_sub_p<FrParams>(v.v, v.v);
    store_fr(out + i, v);
  }
}

// =============================================================================================
// Fr vector helpers of the time prover: everything between the MSMs and the
// sumchecks of snark::Proof::new_time (/root/reference/src/snark/time_prover.rs:19-117) and
// TensorcheckProof::new_time (/root/reference/src/subprotocols/tensorcheck/mod.rs:190-275) so that the
// prover's vectors never leave HBM.  The CPU reference runs all of these as serial O(n) loops.
// =============================================================================================

// out[i] = x^i  (misc::powers, src/misc.rs:59-65); tab.p[k] = x^(2^k)
__global__ void __launch_bounds__(SC_THREADS)
k_fr_powers(Fr* __restrict__ out, size_t n, PowTable tab) {
  const size_t i0 = (size_t)blockIdx.x * SC_TILE + threadIdx.x;
  if (i0 >= n) return;
  Fr t = pow_from_table(tab, i0);
  const Fr step = tab.p[8];  // x^SC_THREADS
#pragma unroll 1
  for (int k = 0; k < SC_K; k++) {
    const size_t i = i0 + (size_t)k * SC_THREADS;
    if (i >= n) break;
    store_fr(out + i, t);
    t = t * step;
  }
}

// (E, O) = (sum_{i even} f_i x^i, sum_{i odd} f_i x^i): f(x) = E + O and f(-x) = E - O in one pass
// (misc::evaluate_le, src/misc.rs:194-199; tensorcheck evaluates every polynomial at beta and -beta,
// tensorcheck/mod.rs:228-247).  tab.p[k] = (x^2)^(2^k).
__global__ void __launch_bounds__(SC_THREADS)
k_fr_eval_even_odd(const Fr* __restrict__ f, size_t n, Fr x, PowTable tab, int kpt, Fr* partials, unsigned int* ticket, Fr* out) {
  const size_t npairs = (n + 1) / 2;
  const size_t i0 = (size_t)blockIdx.x * SC_THREADS * kpt + threadIdx.x;
  Fr e = Fr::zero(), o = Fr::zero();
  Fr t = Fr::one(), step = Fr::one();
  if (i0 < npairs) { t = pow_from_table(tab, i0); step = tab.p[8]; }
#pragma unroll 1
  for (int k = 0; k < kpt; k++) {
    const size_t i = i0 + (size_t)k * SC_THREADS;
    if (i >= npairs) break;
    e = e + load_fr(f + 2 * i) * t;
    if (2 * i + 1 < n) o = o + load_fr(f + 2 * i + 1) * t;
    t = t * step;
  }
  o = o * x;  // sum f_{2i+1} x^(2i) times x
  sc_reduce_and_publish(e, o, partials, ticket, out);
}

// out[idx] = prod_{bit j of idx set} rho_j  (misc::tensor, src/misc.rs:133-149).  Each thread owns 16
// consecutive indices: the product over the high bits once, then the 16 combinations of rho_0..rho_3.
struct TensorArgs {
  Fr rho[32];
  int k;
};
__global__ void __launch_bounds__(256)
k_fr_tensor(Fr* __restrict__ out, size_t n, TensorArgs args) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t base = g * 16;
  if (base >= n) return;
  Fr hi = Fr::one();
  for (int j = 4; j < args.k; j++)
    if ((base >> j) & 1) hi = hi * args.rho[j];
  Fr low[16];
  low[0] = Fr::one();
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (j < args.k) {
#pragma unroll
      for (int m = 0; m < (1 << j); m++) low[(1 << j) + m] = low[m] * args.rho[j];
    }
  }
#pragma unroll
  for (int m = 0; m < 16; m++)
    if (base + m < n) store_fr(out + base + m, m == 0 ? hi : hi * low[m]);
}

// out[i] = a[i] * b[i]  (misc::hadamard, src/misc.rs:205-208)
__global__ void __launch_bounds__(256)
k_fr_hadamard(const Fr* __restrict__ a, const Fr* __restrict__ b, size_t n, Fr* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    store_fr(out + i, load_fr(a + i) * load_fr(b + i));
}

// acc[i] += c * x[i], i < n  (misc::linear_combination, src/misc.rs:37-48, one term at a time)
__global__ void __launch_bounds__(256)
k_fr_axpy(Fr* __restrict__ acc, const Fr* __restrict__ x, size_t n, Fr c, int c_is_one) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fr v = load_fr(x + i);
    if (!c_is_one) v = v * c;
    store_fr(acc + i, load_fr(acc + i) + v);
  }
}

// y[r] = sum_k vals[k] * x[col[k]] over the CSR row r  (misc::product_matrix_vector, src/misc.rs:100-110;
// with the transposed matrices it is the abc_tensored accumulation of snark/time_prover.rs:63-81)
__global__ void __launch_bounds__(256)
k_fr_spmv(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ col, const Fr* __restrict__ vals, size_t nrows,
          const Fr* __restrict__ x, Fr* __restrict__ y) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (size_t)gridDim.x * blockDim.x) {
    Fr acc = Fr::zero();
    for (uint32_t k = rowptr[r]; k < rowptr[r + 1]; k++) acc = acc + load_fr(vals + k) * load_fr(x + col[k]);
    store_fr(y + r, acc);
  }
}

// Synthetic division by (X - a) as a hierarchical suffix-Horner scan (KZG quotients, kzg/time.rs:112-145).
//   S_i = sum_{k>=i} f_k a^(k-i):  quotient q_j = S_{j+1}, remainder = S_0.
// up-sweep:  next[j] = sum_{k<DIV_K} cur[j*DIV_K + k] * a_l^k        (a_l = a^(DIV_K^level))
// down-sweep: within run j start from the carry S_next[j+1] and Horner down, writing every S.
// A run is DIV_K = 8 consecutive elements = 256 contiguous bytes per thread: the whole run is loaded up front with
// 128-bit accesses (two full 128-byte lines per thread, nothing left to the L1) and the dependent Horner chain is 8 long.
// Round 2 started with runs of 32 walked by a load-multiply loop: neighbouring threads 1 KB apart, one load in flight per
// thread, 2.7 ms per division at 2^24 against 0.7 ms of memory time (profiles/r02_prover_pieces.txt: three_divisions).
static constexpr int DIV_K = 8;
static constexpr int DIV_K_LOG = 3;
__device__ __forceinline__ void div_load_run(Fr* run, const Fr* __restrict__ cur, size_t lo, size_t n) {
#pragma unroll
  for (int k = 0; k < DIV_K; k++) run[k] = lo + k < n ? load_fr(cur + lo + k) : Fr::zero();
}
__global__ void __launch_bounds__(128)
k_fr_div_up(const Fr* __restrict__ cur, size_t n, Fr a_l, Fr* __restrict__ next) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t lo = j * DIV_K;
  if (lo >= n) return;
  Fr run[DIV_K];
  div_load_run(run, cur, lo, n);       // elements beyond n are zero: they add nothing to the aggregate
  Fr acc = run[DIV_K - 1];
#pragma unroll
  for (int k = DIV_K - 2; k >= 0; k--) acc = run[k] + a_l * acc;
  store_fr(next + j, acc);
}
// suffix: S of the next level (nullptr at the top level); out_shift = 1 at level 0 (q_j = S_{j+1}) else 0
__global__ void __launch_bounds__(128)
k_fr_div_down(const Fr* __restrict__ cur, size_t n, Fr a_l, const Fr* __restrict__ suffix_next, size_t n_next,
              Fr* __restrict__ out, int out_shift, Fr* __restrict__ rem) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t lo = j * DIV_K;
  if (lo >= n) return;
  Fr run[DIV_K];
  div_load_run(run, cur, lo, n);
  Fr acc = (suffix_next != nullptr && j + 1 < n_next) ? load_fr(suffix_next + j + 1) : Fr::zero();
#pragma unroll
  for (int k = DIV_K - 1; k >= 0; k--) {
    if (lo + k >= n) continue;          // the last run of a level may be short (acc is still the carry = 0 there)
    acc = run[k] + a_l * acc;
    if (out_shift) {
      if (lo + k >= 1) store_fr(out + lo + k - 1, acc);
      else store_fr(rem, acc);
    } else {
      store_fr(out + lo + k, acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define LAUNCH(ctx, kernel, grid, block, shmem, ...)                \
  do {                                                              \
    kernel<<<grid, block, shmem, (ctx)->stream>>>(__VA_ARGS__);     \
    (ctx)->launches++;                                              \
  } while (0)

#define LAUNCH_LN(ln, kernel, grid, block, shmem, ...)             \
  do {                                                              \
    kernel<<<grid, block, shmem, (ln).stream>>>(__VA_ARGS__);       \
    (*(ln).launches)++;                                             \
  } while (0)

static inline unsigned fold_grid_sm(int sm_count, size_t outputs) {
  size_t blocks = (outputs + 255) / 256;
  size_t cap = (size_t)sm_count * 16;
  return (unsigned)std::max<size_t>(1, std::min(blocks, cap));
}
static inline unsigned fold_grid(const gm_ctx* ctx, size_t outputs) { return fold_grid_sm(ctx->sm_count, outputs); }

int fr_fold_dev(const Lane& ln, int sm_count, const Fr* d_f, size_t n, const Fr& r, Fr* d_out) {
  if (n == 0) return GM_OK;
  LAUNCH_LN(ln, k_fr_fold, fold_grid_sm(sm_count, (n + 1) / 2), 256, 0, d_f, n, r, d_out);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int fr_reverse_dev(const Lane& ln, int sm_count, const Fr* d_in, size_t n, Fr* d_out) {
  if (n == 0) return GM_OK;
  LAUNCH_LN(ln, k_fr_reverse, fold_grid_sm(sm_count, (n + 1) / 2), 256, 0, d_in, n, d_out);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int fr_random_dev(gm_ctx* ctx, Fr* d_out, size_t n, uint64_t seed) {
  if (n == 0) return GM_OK;
  LAUNCH(ctx, k_fr_random, fold_grid(ctx, n), 256, 0, d_out, n, seed);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

static PowTable make_pow_table(const Fr& twist, size_t npairs) {
  PowTable tab;
  Fr x = twist.sqr();
  int need = 9;  // p[8] is always read as the per-thread step
  while (need < 40 && (npairs >> need)) need++;
  for (int k = 0; k < 40; k++) {
    if (k < need) { tab.p[k] = x; x = x.sqr(); }
    else tab.p[k] = Fr::one();
  }
  return tab;
}

// pairs per thread: 8 amortises the power-table walk on long vectors; short vectors (the late, latency-bound
// rounds) use every thread for one pair so that a round costs one short dependency chain
static inline int sc_pairs_per_thread(size_t npairs) {
  return npairs > ((size_t)1 << 19) ? 8 : npairs > ((size_t)1 << 17) ? 4 : npairs > ((size_t)1 << 16) ? 2 : 1;
}
static inline unsigned sc_grid(size_t npairs, int kpt) {
  return (unsigned)std::max<size_t>(1, (npairs + (size_t)SC_THREADS * kpt - 1) / ((size_t)SC_THREADS * kpt));
}

size_t sc_max_ctas(size_t nf, size_t ng) {
  size_t npairs = std::max((nf + 1) / 2, (ng + 1) / 2);
  return std::max<size_t>(1024, (npairs + SC_TILE - 1) / SC_TILE);  // >= the largest grid of any later round
}

// Vectors of more than 2^sc_staged_min_log() pairs go through k_sc_staged (GM_SC_STAGED_MIN_LOG=64 switches it off).
// Measured at 2^24 (profiles/r02_sumcheck_staged_ab.txt): whole sumcheck 1.877 ms without it, 1.749 / 1.731 / 1.722 ms
// with the threshold at 2^19 / 2^17 / 2^15 pairs.
static int sc_staged_min_log() {
  static const int v = [] {
    const char* e = getenv("GM_SC_STAGED_MIN_LOG");
    return e ? atoi(e) : 15;
  }();
  return v;
}
static int sc_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  return sms;
}
template <bool TW, bool FOLD>
static int sc_staged_launch(const Lane& ctx, const Fr* d_f, size_t nf, const Fr* d_g, size_t ng, const Fr& rf, const Fr& rg, Fr* d_f_out,
                            Fr* d_g_out, size_t npairs, const Fr& twist, const PowTable& tab, Fr* d_partials, unsigned int* d_ticket,
                            Fr* d_out, ScMailbox* mb, uint32_t seq) {
  constexpr size_t shmem = (size_t)(SC_THREADS / 32) * SC_RING * ScStage<FOLD ? 4 : 2>::BYTES;
  // per device and cheap next to a round of this size: set on every launch rather than cached per process
  GM_CUDA(cudaFuncSetAttribute(k_sc_staged<TW, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
  const size_t tiles = (npairs + SC_THREADS - 1) / SC_THREADS;
  const unsigned grid = (unsigned)std::min<size_t>(tiles, (size_t)2 * sc_sm_count());   // <= 1024 = the partials' capacity
  Fr step = Fr::one();
  if (TW) {   // (twist^2)^(grid * SC_THREADS) from the table of (twist^2)^(2^k)
    const uint64_t e = (uint64_t)grid * SC_THREADS;
    for (int k = 0; k < 40 && (e >> k); k++)
      if ((e >> k) & 1ull) step = step * tab.p[k];
  }
  LAUNCH_LN(ctx, (k_sc_staged<TW, FOLD>), grid, SC_THREADS, shmem, d_f, nf, d_g, ng, rf, rg, d_f_out, d_g_out, npairs, twist, tab, step,
            d_partials, d_ticket, d_out, mb, seq);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int sc_message_dev(const Lane& ctx, const Fr* d_f, size_t nf, const Fr* d_g, size_t ng, const Fr& twist, bool use_twist,
                   Fr* d_partials, unsigned int* d_ticket, Fr* d_out, ScMailbox* mb, uint32_t seq) {
  const size_t npairs = std::min((nf + 1) / 2, (ng + 1) / 2);
  const int kpt = sc_pairs_per_thread(npairs);
  const unsigned grid = sc_grid(npairs, kpt);
  const bool staged = sc_staged_min_log() < 63 && npairs > ((size_t)1 << sc_staged_min_log());
  if (use_twist) {
    PowTable tab = make_pow_table(twist, npairs);
    if (staged)
      return sc_staged_launch<true, false>(ctx, d_f, nf, d_g, ng, twist, twist, nullptr, nullptr, npairs, twist, tab, d_partials,
                                           d_ticket, d_out, mb, seq);
    LAUNCH_LN(ctx, k_sc_message<true>, grid, SC_THREADS, 0, d_f, nf, d_g, ng, twist, tab, kpt, d_partials, d_ticket, d_out, mb, seq);
  } else {
    PowTable tab;  // unused
    if (staged)
      return sc_staged_launch<false, false>(ctx, d_f, nf, d_g, ng, twist, twist, nullptr, nullptr, npairs, twist, tab, d_partials,
                                            d_ticket, d_out, mb, seq);
    LAUNCH_LN(ctx, k_sc_message<false>, grid, SC_THREADS, 0, d_f, nf, d_g, ng, twist, tab, kpt, d_partials, d_ticket, d_out, mb, seq);
  }
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int sc_fold_message_dev(const Lane& ctx, const Fr* d_f, size_t nf, const Fr* d_g, size_t ng, const Fr& rf, const Fr& rg,
                        Fr* d_f_out, Fr* d_g_out, const Fr& new_twist, bool use_twist, Fr* d_partials,
                        unsigned int* d_ticket, Fr* d_out, ScMailbox* mb, uint32_t seq) {
  const size_t nf2 = (nf + 1) / 2, ng2 = (ng + 1) / 2;
  const size_t npairs = std::max((nf2 + 1) / 2, (ng2 + 1) / 2);
  const int kpt = sc_pairs_per_thread(npairs);
  const unsigned grid = sc_grid(npairs, kpt);
  const bool staged = sc_staged_min_log() < 63 && npairs > ((size_t)1 << sc_staged_min_log());
  if (use_twist) {
    PowTable tab = make_pow_table(new_twist, npairs);
    if (staged)
      return sc_staged_launch<true, true>(ctx, d_f, nf, d_g, ng, rf, rg, d_f_out, d_g_out, npairs, new_twist, tab, d_partials,
                                          d_ticket, d_out, mb, seq);
    LAUNCH_LN(ctx, k_sc_fold_message<true>, grid, SC_THREADS, 0, d_f, nf, d_g, ng, rf, rg, d_f_out, d_g_out, new_twist, tab, kpt,
           d_partials, d_ticket, d_out, mb, seq);
  } else {
    PowTable tab;
    if (staged)
      return sc_staged_launch<false, true>(ctx, d_f, nf, d_g, ng, rf, rg, d_f_out, d_g_out, npairs, new_twist, tab, d_partials,
                                           d_ticket, d_out, mb, seq);
    LAUNCH_LN(ctx, k_sc_fold_message<false>, grid, SC_THREADS, 0, d_f, nf, d_g, ng, rf, rg, d_f_out, d_g_out, new_twist, tab, kpt,
           d_partials, d_ticket, d_out, mb, seq);
  }
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

// ---- vector helpers ---------------------------------------------------------------------------
static PowTable make_pow_table_base(const Fr& base, size_t count) {
  PowTable tab;
  Fr x = base;
  int need = 9;
  while (need < 40 && (count >> need)) need++;
  for (int k = 0; k < 40; k++) {
    if (k < need) { tab.p[k] = x; x = x.sqr(); }
    else tab.p[k] = Fr::one();
  }
  return tab;
}

int fr_powers_dev(gm_ctx* ctx, const Fr& x, size_t n, Fr* d_out) {
  if (n == 0) return GM_OK;
  PowTable tab = make_pow_table_base(x, n);
  LAUNCH(ctx, k_fr_powers, (unsigned)((n + SC_TILE - 1) / SC_TILE), SC_THREADS, 0, d_out, n, tab);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int fr_eval_even_odd_dev(gm_ctx* ctx, const Fr* d_f, size_t n, const Fr& x, Fr* d_partials, unsigned int* d_ticket, Fr* d_out) {
  const size_t npairs = (n + 1) / 2;
  PowTable tab = make_pow_table_base(x.sqr(), npairs);
  const int kpt = sc_pairs_per_thread(npairs);
  const unsigned grid = sc_grid(npairs, kpt);
  LAUNCH(ctx, k_fr_eval_even_odd, grid, SC_THREADS, 0, d_f, n, x, tab, kpt, d_partials, d_ticket, d_out);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int fr_tensor_dev(gm_ctx* ctx, const Fr* rho, int k, Fr* d_out) {
  if (k < 0 || k > 32) { set_error("tensor: 0..32 challenges supported"); return GM_ERR_ARG; }
  TensorArgs args;
  for (int j = 0; j < 32; j++) args.rho[j] = j < k ? rho[j] : Fr::one();
  args.k = k;
  const size_t n = (size_t)1 << k;
  const size_t threads = (n + 15) / 16;
  LAUNCH(ctx, k_fr_tensor, (unsigned)((threads + 255) / 256), 256, 0, d_out, n, args);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int fr_hadamard_dev(gm_ctx* ctx, const Fr* d_a, const Fr* d_b, size_t n, Fr* d_out) {
  if (n == 0) return GM_OK;
  LAUNCH(ctx, k_fr_hadamard, fold_grid(ctx, n), 256, 0, d_a, d_b, n, d_out);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int fr_axpy_dev(gm_ctx* ctx, Fr* d_acc, const Fr* d_x, size_t n, const Fr& c) {
  if (n == 0) return GM_OK;
  LAUNCH(ctx, k_fr_axpy, fold_grid(ctx, n), 256, 0, d_acc, d_x, n, c, c == Fr::one() ? 1 : 0);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int fr_spmv_dev(gm_ctx* ctx, const uint32_t* d_rowptr, const uint32_t* d_col, const Fr* d_vals, size_t nrows, const Fr* d_x, Fr* d_y) {
  if (nrows == 0) return GM_OK;
  LAUNCH(ctx, k_fr_spmv, fold_grid(ctx, nrows), 256, 0, d_rowptr, d_col, d_vals, nrows, d_x, d_y);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

size_t fr_div_scratch_elems(size_t n) {
  size_t tot = 0;
  while (n > 1) { n = (n + DIV_K - 1) / DIV_K; tot += 2 * n; }
  return tot + 2;
}

// q (n-1 elements) and remainder of f / (X - a); d_scratch holds fr_div_scratch_elems(n) elements
int fr_div_linear_dev(gm_ctx* ctx, const Fr* d_f, size_t n, const Fr& a, Fr* d_q, Fr* d_rem, Fr* d_scratch) {
  if (n == 0) { GM_CUDA(cudaMemsetAsync(d_rem, 0, 32, ctx->stream)); return GM_OK; }
  // level arrays: A^l (aggregates) and S^l (suffixes), l >= 1
  std::vector<size_t> sizes{n};
  std::vector<Fr> apow{a};
  while (sizes.back() > 1) {
    sizes.push_back((sizes.back() + DIV_K - 1) / DIV_K);
    Fr p = apow.back();
    for (int k = 0; k < DIV_K_LOG; k++) p = p.sqr();  // ^DIV_K
    apow.push_back(p);
  }
  const int levels = (int)sizes.size() - 1;
  std::vector<Fr*> A(levels + 1), S(levels + 1);
  A[0] = const_cast<Fr*>(d_f);
  Fr* cursor = d_scratch;
  for (int l = 1; l <= levels; l++) { A[l] = cursor; cursor += sizes[l]; S[l] = cursor; cursor += sizes[l]; }
  for (int l = 0; l < levels; l++) {
    const size_t runs = sizes[l + 1];
    LAUNCH(ctx, k_fr_div_up, (unsigned)((runs + 127) / 128), 128, 0, A[l], sizes[l], apow[l], A[l + 1]);
  }
  // top level has one element: its suffix is itself
  if (levels >= 1) GM_CUDA(cudaMemcpyAsync(S[levels], A[levels], 32, cudaMemcpyDeviceToDevice, ctx->stream));
  for (int l = levels - 1; l >= 0; l--) {
    const size_t runs = sizes[l + 1];
    if (l == 0)
      LAUNCH(ctx, k_fr_div_down, (unsigned)((runs + 127) / 128), 128, 0, A[0], sizes[0], apow[0], S[1], sizes[1], d_q, 1, d_rem);
    else
      LAUNCH(ctx, k_fr_div_down, (unsigned)((runs + 127) / 128), 128, 0, A[l], sizes[l], apow[l], S[l + 1], sizes[l + 1], S[l], 0, d_rem);
  }
  if (levels == 0) {  // n == 1: quotient empty, remainder f_0
    GM_CUDA(cudaMemcpyAsync(d_rem, d_f, 32, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

}  // namespace gm
