// G1 multi-scalar multiplication for sm_100a: sum_i s_i * P_i over BLS12-381.
//
// Replaces ark_ec::VariableBaseMSM::{msm_unchecked, msm_bigint} (ark-ec 0.4.2, not vendored in
// the reference; written spec: /root/reference/src/kzg/msm/variable_base.rs:16-177) behind
// CommitterKey::commit (/root/reference/src/kzg/time.rs:81-83) and msm_chunks
// (/root/reference/src/kzg/space.rs:22-55).  Same signed-digit bucket method, re-shaped for a GPU:
//
//   1 k_digits_hist     Fr Montgomery -> canonical (into_bigint), signed radix-2^c digits
//                       (variable_base.rs:21-61), per-bucket histogram            [HBM: 32 B/term in]
//   2 scan              exclusive prefix sum of the W * 2^(c-1) bucket counts
//   3 k_scatter         counting sort: point references grouped by (window, bucket)
//   3b k_aff_prepare / k_aff_invert / k_aff_finish  (x R levels)
//                       pairwise bucket sums in AFFINE coordinates with shared inversions (g1_affine.cuh):
//                       6 instead of 10 Fq products per addition; the survivors go on to rows 4-6.
//                       Every bucket's run of references starts at a multiple of 2^R (the scan pads the counts), so
//                       pair s of a level is positions (2s, 2s+1) of the level below: no per-slot search, no
//                       per-level counts / scans, coalesced loads; level outputs are x / y planes (SoA)
//   4 k_classify / k_worklist_fill
//                       buckets are cut into work items of <= SPLIT references and the items are
//                       ordered by size (largest first) so that the 32 lanes of a warp run equally
//                       long loops; a bucket holding *every* point (all-equal scalars, the
//                       reference's dummy_r1cs default, src/circuit.rs:349-365) becomes thousands of items
//   5 k_accumulate      one thread per work item: XYZZ += affine base (gathered 96 B/point,
//                       128-bit loads), complete formulas (identity, P+P, P-P)
//   6 k_split_combine   one CTA per split bucket: tree-sum of its partial sums
//   7 k_bucket_chunks   running-sum reduction (variable_base.rs:150-166) over slices of L buckets
//   8 k_rowcol          the chunk sums S_t carry weights t*L: row / column sums of the t = hi*L2 + lo matrix
//   9 k_weighted, k_window_total   one short scalar multiplication per row and per column, CTA tree-sums,
//                       per-window total times 2^(c*w) (variable_base.rs:168-175)
//  10 k_final           sum of windows, += device-resident accumulator, optional normalisation
// Rows 1-6 live in this file, rows 7-10 in msm_reduce.cu, the key-setup kernels in srs.cu.
//
// The window size c is chosen by a cost model for the GPU (not arkworks' ln(n)+2): the result is a
// unique group element, so any c gives the bit-identical normalised output.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "g1.cuh"
#include "g1_affine.cuh"
#include "msm.cuh"
#include "msm_mem.cuh"

namespace gm {

static constexpr uint32_t SKIP = 0xFFFFFFFFu;
static constexpr uint32_t NONE = 0xFFFFFFFFu;
static constexpr int SPLIT = 1024;         // upper bound of the point references per work item (the runtime value is a kernel argument)
static constexpr int ACC_THREADS = 128;

// Warp-aggregated atomicAdd: lanes that target the same counter elect a leader which adds the group size
// once; every lane gets its own slot.  For uniformly random keys this is a no-op in cost terms, for the
// reference's default all-equal scalars (one hot bucket per window) it removes 31/32 of the contended
// L2 atomics.  Must be called by all 32 lanes (inactive lanes pass active = false).
__device__ __forceinline__ uint32_t warp_agg_atomic_inc(uint32_t* counters, size_t key, bool active) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  // cheap filter: a hot counter shows up as equal keys in neighbouring lanes; uniformly random keys almost
  // never do, and then the plain atomic is cheaper than the match
  const unsigned long long nb_key = __shfl_down_sync(0xffffffffu, (unsigned long long)key, 1);
  const bool nb_active = __shfl_down_sync(0xffffffffu, active ? 1 : 0, 1) != 0;
  const bool dup = active && nb_active && lane < 31 && nb_key == (unsigned long long)key;
  if (!__any_sync(0xffffffffu, dup)) return active ? atomicAdd(counters + key, 1u) : 0u;
  if (!active) return 0;
  const unsigned peers = __match_any_sync(mask, (unsigned long long)key);
  const int leader = __ffs(peers) - 1;
  const unsigned rank = __popc(peers & ((1u << lane) - 1u));
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(counters + key, (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  return base + rank;
}

struct Meta {
  uint32_t n_items;
  uint32_t n_split;
  uint32_t n_partials;
  uint32_t pad;
  uint32_t size_hist[SPLIT + 1];
  uint32_t size_base[SPLIT + 1];
  uint32_t size_fill[SPLIT + 1];
};

// -------------------------------------------------------------------------------------------
// 1. digits + histogram
// -------------------------------------------------------------------------------------------
// stride: term i reads scalar i * stride (1 = dense; world size when a vector is dealt out cyclically to the ranks)
// [i0, i1): the terms this launch converts (the whole vector, or the piece of a host vector that has just arrived)
__global__ void k_digits_hist(const uint32_t* __restrict__ scalars, uint32_t i0, uint32_t i1, uint32_t n, size_t stride, int is_bigint, int c, int W,
                              int merged, uint32_t* __restrict__ digits, uint32_t* __restrict__ counts) {
  const uint32_t i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < i1;
  Fr s = Fr::zero();
  if (live) {
    const uint4* p = reinterpret_cast<const uint4*>(scalars + (size_t)i * stride * 8);
    uint4 lo = __ldg(p), hi = __ldg(p + 1);
    s.v[0] = lo.x; s.v[1] = lo.y; s.v[2] = lo.z; s.v[3] = lo.w;
    s.v[4] = hi.x; s.v[5] = hi.y; s.v[6] = hi.z; s.v[7] = hi.w;
  }
  if (!is_bigint) {
    s = s.from_mont();  // into_bigint
  } else {
    // msm_bigint takes canonical BigInt<4>; fold any 256-bit value into [0, r) (2^256 < 3r)
    detail::cond_sub_p<FrParams>(s.v, s.v);
    detail::cond_sub_p<FrParams>(s.v, s.v);
  }
  uint32_t limb[9];
#pragma unroll
  for (int j = 0; j < 8; j++) limb[j] = s.v[j];
  limb[8] = 0;
  const uint32_t nb = 1u << (c - 1);
  const uint32_t mask = (1u << c) - 1;
  uint32_t carry = 0;
  for (int w = 0; w < W; w++) {
    const int bit = w * c;
    const int k = bit >> 5, sh = bit & 31;
    uint32_t raw = 0;
    if (k < 8) {
      uint64_t two = ((uint64_t)limb[k + 1] << 32) | limb[k];
      raw = (uint32_t)(two >> sh) & mask;
    }
    uint32_t coef = raw + carry;
    uint32_t code;
    if (w == W - 1) {
      // last window keeps its carry (variable_base.rs:58): digit = coef >= 0
      code = coef == 0 ? SKIP : (coef - 1);
      carry = 0;
    } else {
      carry = coef >= nb ? 1u : 0u;
      if (coef == 0 || coef == (1u << c)) code = SKIP;       // digit 0
      else if (carry) code = ((1u << c) - coef - 1) | 0x80000000u;  // digit = coef - 2^c < 0
      else code = coef - 1;
    }
    if (live) digits[(size_t)w * n + i] = code;
    // merged: every window shares one bucket set (the bases are pre-multiplied by 2^(c*w))
    warp_agg_atomic_inc(counts, (merged ? 0 : (size_t)w * nb) + (code & 0x7FFFFFFFu), live && code != SKIP);
  }
}

// -------------------------------------------------------------------------------------------
// 2. exclusive scan (three small kernels; M = W * 2^(c-1) <= a few million)
// -------------------------------------------------------------------------------------------
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 8;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* sh, uint32_t* total) {
  // sh: SCAN_THREADS/32 + 1 words
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t t = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    sh[lane] = t;  // inclusive warp totals
  }
  __syncthreads();
  uint32_t warp_off = wid ? sh[wid - 1] : 0;
  *total = sh[(blockDim.x >> 5) - 1];
  uint32_t r = warp_off + x - v;
  __syncthreads();
  return r;
}

// pad = 2^R - 1: every count is rounded up to a multiple of 2^R, so every start is a multiple of 2^R (affine levels)
__global__ void k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ tile_sums, uint32_t M, uint32_t pad) {
  __shared__ uint32_t sh[33];
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < M) ? ((in[base + k] + pad) & ~pad) : 0; s += v[k]; }
  uint32_t total;
  uint32_t off = block_exclusive_scan(s, sh, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < M) out[base + k] = off; off += v[k]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void k_scan_tile_sums(uint32_t* tile_sums, uint32_t ntiles) {
  __shared__ uint32_t sh[33];
  uint32_t running = 0;
  for (uint32_t b = 0; b < ntiles; b += blockDim.x) {
    uint32_t i = b + threadIdx.x;
    uint32_t v = i < ntiles ? tile_sums[i] : 0, total;
    uint32_t e = block_exclusive_scan(v, sh, &total);
    if (i < ntiles) tile_sums[i] = running + e;
    running += total;
  }
}
__global__ void k_scan_add(uint32_t* __restrict__ out, uint32_t* __restrict__ copy, const uint32_t* __restrict__ tile_sums, uint32_t M) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  uint32_t v = out[i] + tile_sums[i / SCAN_TILE];
  out[i] = v;
  copy[i] = v;
}

// -------------------------------------------------------------------------------------------
// 3. counting-sort scatter
// -------------------------------------------------------------------------------------------
// Measured on B200 (profiles/r02_summary.md): 2^24 terms = 201 M references in 6.2 ms.  The kernel waits on the cursor atomics
// (they return the position the store depends on); 4 independent atomics per thread, and a two-pass variant that first groups
// the references by the top 8 bits of the bucket index so that the scatter works inside an L2-sized window, were both
// measured and gave 6.2 and 7.5 ms: the rate of returning L2 atomics is the bound, not latency or DRAM traffic.  A
// register-free prefetch.global.L2 ring for the table gathers of the first affine level was measured too: 65.8 vs 55.1 ms,
// and so was staging those gathers through shared memory with cp.async, four slots per thread in flight instead of one
// (k_aff_prepare_staged, parity green): 58.9 vs 54.8 ms at 2^24, 4.52 vs 4.28 ms at 2^20.  More requests in flight make
// the first level SLOWER: its 2 x 10^8 random 48-byte reads run at the rate the memory system takes them, not at a latency.
__global__ void k_scatter(const uint32_t* __restrict__ digits, uint32_t n, int W, uint32_t nb, int merged, uint32_t ref_offset,
                          uint32_t ref_stride, uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = blockIdx.y;
  const uint32_t code = i < n ? digits[(size_t)w * n + i] : SKIP;
  const bool live = code != SKIP;
  const uint32_t pos = warp_agg_atomic_inc(cursor, (merged ? 0 : (size_t)w * nb) + (code & 0x7FFFFFFFu), live);
  // reference into the base table: level w of the precomputed table when merged
  if (live) sorted[pos] = (i + ref_offset + (uint32_t)w * ref_stride) | (code & 0x80000000u);
}

// -------------------------------------------------------------------------------------------
// 3b. affine pre-reduction: the references of every bucket are summed pairwise, level by level, in AFFINE
//     coordinates with shared inversions (g1_affine.cuh): 6 Fq products per addition instead of the 10 of the
//     XYZZ mixed addition of k_accumulate.  One level = three launches:
//       k_aff_prepare  classify each pair, denominators, per-thread exclusive prefix products (to HBM),
//                      per-warp butterfly giving every lane the product of the OTHER lanes' totals
//       k_aff_invert   one Kaliski inversion per warp total
//       k_aff_finish   back-substitution through the prefix products, chord / tangent formulas, 96 B out
//     Output slot s of bucket gb holds in[2j] + in[2j+1] (j = s - out_starts[gb]); an odd leftover passes
//     through.  A warp owns 32*G consecutive slots, lane l the slots base + k*32 + l (coalesced).
//     After R levels the surviving points (ceil(cnt / 2^R) per bucket) go through the XYZZ work-list path.
// -------------------------------------------------------------------------------------------
static constexpr int AFF_THREADS = 128;

// A level's input: level 0 gathers table points through the sorted references (NONE marks the padding between
// buckets), later levels read the x / y planes the level below wrote ((0,0) = identity = padding).
struct AffIn {
  const Affine* table;       // level 0
  const uint32_t* refs;      // level 0: sorted references, 2 * n_slots of them
  const Fq* x;               // level >= 1
  const Fq* y;
  int rec_q;
};

__device__ __forceinline__ Fq shfl_xor_fq(const Fq& v, int m) {
  Fq r;
#pragma unroll
  for (int k = 0; k < 12; k++) r.v[k] = __shfl_xor_sync(0xffffffffu, v.v[k], m);
  return r;
}

template <bool FIRST>
__device__ __forceinline__ Affine aff_load_point(const AffIn& in, uint32_t e, uint32_t ref) {
  Affine p;
  if (FIRST) {
    if (ref == NONE) { p.x = Fq::zero(); p.y = Fq::zero(); return p; }
    p = load_ro(rec_at(in.table, ref & 0x7FFFFFFFu, in.rec_q));
    if (ref >> 31) p.y = p.y.neg();   // -(0,0) = (0,0): the identity stays the identity
    return p;
  }
  p.x = load_ro(in.x + e);
  p.y = load_ro(in.y + e);
  return p;
}

// what k_aff_prepare knows about a slot after the cheap part: the two references (level 0) and the x coordinates
struct AffSlot {
  uint2 ref;      // level 0 only
  Fq x1, x2;
  bool valid;
};

template <bool FIRST>
__device__ __forceinline__ void aff_fetch_slot(AffSlot& sl, const AffIn& in, uint32_t s, uint32_t n_slots) {
  sl.valid = s < n_slots;
  if (!sl.valid) return;
  if (FIRST) {
    sl.ref = __ldg(reinterpret_cast<const uint2*>(in.refs) + s);
    sl.x1 = sl.ref.x == NONE ? Fq::zero() : load_ro(&rec_at(in.table, sl.ref.x & 0x7FFFFFFFu, in.rec_q)->x);
    sl.x2 = sl.ref.y == NONE ? Fq::zero() : load_ro(&rec_at(in.table, sl.ref.y & 0x7FFFFFFFu, in.rec_q)->x);
  } else {
    sl.x1 = load_ro(in.x + 2u * s);
    sl.x2 = load_ro(in.x + 2u * s + 1u);
  }
}

// slots of one level = ceil(S / 2^(level+1)), S = padded length of the sorted reference array (device side)
__device__ __forceinline__ uint32_t aff_level_slots(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ counts, uint32_t M,
                                                    uint32_t pad, int level) {
  const uint32_t S = __ldg(starts + M - 1) + ((__ldg(counts + M - 1) + pad) & ~pad);
  return S >> (level + 1);
}

template <bool FIRST>
__global__ void __launch_bounds__(AFF_THREADS, 4)
k_aff_prepare(AffIn in, const uint32_t* __restrict__ starts, const uint32_t* __restrict__ counts, uint32_t M, uint32_t pad, int level, int G,
              Fq* __restrict__ prefix, uint8_t* __restrict__ kinds, Fq* __restrict__ others, Fq* __restrict__ warp_totals) {
  const uint32_t n_slots = aff_level_slots(starts, counts, M, pad, level);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t base64 = (uint64_t)warp * 32u * (uint32_t)G;
  if (base64 >= n_slots) return;                      // whole warp out of range
  const uint32_t base = (uint32_t)base64;
  Fq run = Fq::one();
  // the loads of slot k+1 (references, x coordinates) are issued before the product of slot k
  AffSlot cur, nxt;
  aff_fetch_slot<FIRST>(cur, in, base + lane, n_slots);
#pragma unroll 1
  for (int k = 0; k < G; k++) {
    if (!cur.valid) break;
    const uint32_t s = base + (uint32_t)k * 32u + lane;
    nxt.valid = false;
    if (k + 1 < G) aff_fetch_slot<FIRST>(nxt, in, s + 32u, n_slots);
    Fq den = cur.x2 - cur.x1;
    uint32_t kind = PK_ADD;
    if (FIRST && cur.ref.y == NONE) {
      kind = cur.ref.x == NONE ? PK_ZERO : PK_PASS1;      // padding / odd leftover of a bucket
    } else if (den.is_zero() || cur.x1.is_zero() || cur.x2.is_zero()) {
      // rare (common only in the padding of the upper levels): equal x (P + P, P - P) or a possible identity (0, 0)
      const Affine p1 = aff_load_point<FIRST>(in, 2u * s, FIRST ? cur.ref.x : 0u);
      const Affine p2 = aff_load_point<FIRST>(in, 2u * s + 1u, FIRST ? cur.ref.y : 0u);
      kind = aff_pair_kind(p1, p2, true, den);
      if (kind == PK_PASS1 && p1.is_identity()) kind = PK_ZERO;
    }
    kinds[s] = (uint8_t)kind;
    if (aff_kind_needs_inverse(kind)) {
      store_rw(prefix + s, run);
      run = run * den;
    }
    cur = nxt;
  }
  // butterfly: g = product of the lanes of my group, o = product of the group WITHOUT my own total
  Fq g = run, o = Fq::one();
#pragma unroll 1
  for (int m = 1; m < 32; m <<= 1) {
    const Fq pg = shfl_xor_fq(g, m);
    o = (m == 1) ? pg : o * pg;
    g = g * pg;
  }
  store_rw(others + (size_t)warp * 32u + lane, o);
  if (lane == 0) store_rw(warp_totals + warp, g);
}

// spread = 32: one inversion per WARP (lane 0): with few inversions (they are a serial stage of the level) it is faster
// to leave 31 lanes idle than to run 32 data-dependent loops in lock step.  spread = 1: one inversion per thread.
__global__ void __launch_bounds__(128)
k_aff_invert(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ counts, uint32_t M, uint32_t pad, int level, int G, int spread,
             Fq* __restrict__ warp_totals) {
  const uint32_t n_slots = aff_level_slots(starts, counts, M, pad, level);
  const uint64_t n_warps = ((uint64_t)n_slots + 32u * (uint32_t)G - 1u) / (32u * (uint32_t)G);
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t % (uint32_t)spread) return;
  const uint64_t i = t / (uint32_t)spread;
  if (i >= n_warps) return;
  store_rw(warp_totals + i, fp_inv_serial(load_rw(warp_totals + i)));
}

template <bool FIRST>
__global__ void __launch_bounds__(AFF_THREADS, 4)
k_aff_finish(AffIn in, const uint32_t* __restrict__ starts, const uint32_t* __restrict__ counts, uint32_t M, uint32_t pad, int level, int G,
             const Fq* __restrict__ prefix, const uint8_t* __restrict__ kinds, const Fq* __restrict__ others,
             const Fq* __restrict__ warp_totals, Fq* __restrict__ out_x, Fq* __restrict__ out_y) {
  const uint32_t n_slots = aff_level_slots(starts, counts, M, pad, level);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t base64 = (uint64_t)warp * 32u * (uint32_t)G;
  if (base64 >= n_slots) return;
  const uint32_t base = (uint32_t)base64;
  // 1 / (my total) = 1 / (warp total) * (product of the other lanes' totals)
  Fq inv = load_rw(warp_totals + warp) * load_rw(others + (size_t)warp * 32u + lane);
#pragma unroll 1
  for (int k = G - 1; k >= 0; k--) {
    const uint32_t s = base + (uint32_t)k * 32u + lane;
    if (s >= n_slots) continue;
    const uint32_t kind = kinds[s];
    Affine r;
    if (kind == PK_ZERO) { r.x = Fq::zero(); r.y = Fq::zero(); }
    else {
      uint2 ref = make_uint2(0u, 0u);
      if (FIRST) ref = __ldg(reinterpret_cast<const uint2*>(in.refs) + s);
      if (kind == PK_PASS1) r = aff_load_point<FIRST>(in, 2u * s, ref.x);
      else if (kind == PK_PASS2) r = aff_load_point<FIRST>(in, 2u * s + 1u, ref.y);
      else {
        const Affine p1 = aff_load_point<FIRST>(in, 2u * s, ref.x);
        const Affine p2 = aff_load_point<FIRST>(in, 2u * s + 1u, ref.y);
        const Fq den = (kind == PK_ADD) ? (p2.x - p1.x) : p1.y.dbl();
        const Fq inv_den = inv * load_rw(prefix + s);
        inv = inv * den;
        r = aff_pair_finish(kind, p1, p2, inv_den);
      }
    }
    store_rw(out_x + s, r.x);
    store_rw(out_y + s, r.y);
  }
}

// -------------------------------------------------------------------------------------------
// 4. work list, largest items first
// -------------------------------------------------------------------------------------------
// shift = number of affine levels that ran: bucket gb now holds ceil(counts[gb] / 2^shift) points
__device__ __forceinline__ uint32_t eff_count(const uint32_t* __restrict__ counts, uint32_t gb, int shift) {
  return (counts[gb] + (1u << shift) - 1u) >> shift;
}

__global__ void k_classify(const uint32_t* __restrict__ counts, uint32_t M, int shift, uint32_t split, uint32_t* __restrict__ poff,
                           uint32_t* __restrict__ split_list, Meta* meta) {
  __shared__ uint32_t sh[SPLIT + 1];
  for (uint32_t k = threadIdx.x; k <= split; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  const uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x;
  if (gb < M) {
    const uint32_t cnt = eff_count(counts, gb, shift);
    uint32_t po = NONE;
    if (cnt) {
      const uint32_t m = (cnt + split - 1) / split;
      const uint32_t tail = cnt - (m - 1) * split;
      atomicAdd(&sh[tail], 1u);
      if (m > 1) {
        atomicAdd(&sh[split], m - 1);
        const uint32_t si = atomicAdd(&meta->n_split, 1u);
        po = atomicAdd(&meta->n_partials, m);
        split_list[si] = gb;
      }
    }
    poff[gb] = po;
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k <= split; k += blockDim.x)
    if (sh[k]) atomicAdd(&meta->size_hist[k], sh[k]);
}

__global__ void k_size_scan(Meta* meta, uint32_t split) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint32_t run = 0;
    for (int s = (int)split; s >= 1; s--) { meta->size_base[s] = run; run += meta->size_hist[s]; }
    meta->size_base[0] = run;
    meta->n_items = run;
  }
}

__global__ void k_worklist_fill(const uint32_t* __restrict__ counts, uint32_t M, int shift, uint32_t split, uint2* __restrict__ work, Meta* meta) {
  __shared__ uint32_t blk_cnt[SPLIT + 1];
  __shared__ uint32_t blk_base[SPLIT + 1];
  for (uint32_t k = threadIdx.x; k <= split; k += blockDim.x) blk_cnt[k] = 0;
  __syncthreads();
  const uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t cnt = 0, m = 0, tail = 0, r_tail = 0, r_full = 0;
  if (gb < M) {
    cnt = eff_count(counts, gb, shift);
    if (cnt) {
      m = (cnt + split - 1) / split;
      tail = cnt - (m - 1) * split;
      r_tail = atomicAdd(&blk_cnt[tail], 1u);
      if (m > 1) r_full = atomicAdd(&blk_cnt[split], m - 1);
    }
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k <= split; k += blockDim.x)
    if (blk_cnt[k]) blk_base[k] = meta->size_base[k] + atomicAdd(&meta->size_fill[k], blk_cnt[k]);
  __syncthreads();
  if (cnt) {
    work[blk_base[tail] + r_tail] = make_uint2(gb, m - 1);
    if (m > 1) {
      const uint32_t p = blk_base[split] + r_full;
      for (uint32_t k = 0; k + 1 < m; k++) work[p + k] = make_uint2(gb, k);
    }
  }
}

// -------------------------------------------------------------------------------------------
// 5. bucket accumulation: one thread per work item
// -------------------------------------------------------------------------------------------
// DIRECT: the input is the x / y planes left by the affine levels (slot = position, no sign), not a reference list
template <bool DIRECT>
__global__ void __launch_bounds__(ACC_THREADS, 3)  // 3 CTAs/SM: at most 168 registers per thread
k_accumulate(const Affine* __restrict__ bases, const Fq* __restrict__ px, const Fq* __restrict__ py, const uint32_t* __restrict__ sorted,
             const uint32_t* __restrict__ counts, const uint32_t* __restrict__ starts, int shift, const uint32_t* __restrict__ poff,
             const uint2* __restrict__ work, const Meta* __restrict__ meta, uint32_t split, int rec_q, XYZZ* __restrict__ buckets,
             XYZZ* __restrict__ partials, uint32_t* __restrict__ live) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= meta->n_items) return;
  const uint2 item = work[j];
  const uint32_t gb = item.x, k = item.y;
  const uint32_t cnt = eff_count(counts, gb, shift);
  const uint32_t first = (starts[gb] >> shift) + k * split;
  const uint32_t len = min(split, cnt - k * split);
  const uint32_t po = poff[gb];
  // streamed MSM: the bucket persists across chunks (live[gb] != 0 once it has been written)
  XYZZ acc = (live != nullptr && po == NONE && live[gb]) ? load_rw(buckets + gb) : XYZZ::identity();
  for (uint32_t e = 0; e < len; e++) {
    Affine p;
    if (DIRECT) {
      p.x = load_ro(px + first + e);
      p.y = load_ro(py + first + e);
    } else {
      const uint32_t ref = __ldg(sorted + first + e);
      p = load_ro(rec_at(bases, ref & 0x7FFFFFFFu, rec_q));
      if (ref >> 31) p.y = p.y.neg();
    }
    xyzz_madd(acc, p);
  }
  XYZZ* dst = (po == NONE) ? (buckets + gb) : (partials + po + k);
  store_rw(dst, acc);
  if (live != nullptr && po == NONE) live[gb] = 1u;
}


// 6. split buckets: bucket = sum of its partials.  Buckets with at most 32 partials are summed by one warp
//    (shuffle tree, no CTA barrier); larger ones (a bucket that holds a large share of all points) by a CTA.
__device__ __forceinline__ XYZZ shfl_down_xyzz(const XYZZ& v, int delta) {
  XYZZ r;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&v);
  uint32_t* dst = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int k = 0; k < 48; k++) dst[k] = __shfl_down_sync(0xffffffffu, src[k], delta);
  return r;
}

__global__ void __launch_bounds__(RED_THREADS)
k_split_combine(const uint32_t* __restrict__ split_list, const uint32_t* __restrict__ counts, int shift, const uint32_t* __restrict__ poff,
                const Meta* __restrict__ meta, uint32_t split, const XYZZ* __restrict__ partials, XYZZ* __restrict__ buckets,
                uint32_t* __restrict__ live) {
  extern __shared__ uint4 sh_raw[];
  XYZZ* sh = reinterpret_cast<XYZZ*>(sh_raw);
  const uint32_t ns = meta->n_split;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  const uint32_t gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  for (uint32_t s = gwarp; s < ns; s += gridDim.x * warps_per_cta) {
    const uint32_t gb = split_list[s];
    const uint32_t m = (eff_count(counts, gb, shift) + split - 1) / split;
    if (m > 32) continue;
    XYZZ acc = XYZZ::identity();
    if (lane < m) acc = load_rw(partials + poff[gb] + lane);
#pragma unroll 1
    for (int d = 16; d > 0; d >>= 1) {
      XYZZ other = shfl_down_xyzz(acc, d);
      if (lane + d < 32) xyzz_add(acc, other);
    }
    if (lane == 0) {
      if (live != nullptr) {
        if (live[gb]) { XYZZ prev = load_rw(buckets + gb); xyzz_add(acc, prev); }
        live[gb] = 1u;
      }
      store_rw(buckets + gb, acc);
    }
  }
  for (uint32_t s = blockIdx.x; s < ns; s += gridDim.x) {
    const uint32_t gb = split_list[s];
    const uint32_t m = (eff_count(counts, gb, shift) + split - 1) / split;
    if (m <= 32) continue;
    const XYZZ* src = partials + poff[gb];
    XYZZ acc = XYZZ::identity();
    for (uint32_t k = threadIdx.x; k < m; k += blockDim.x) {
      XYZZ b = load_rw(src + k);
      xyzz_add(acc, b);
    }
    XYZZ tot = block_sum_xyzz(acc, sh);
    if (threadIdx.x == 0) {
      if (live != nullptr) {
        if (live[gb]) { XYZZ prev = load_rw(buckets + gb); xyzz_add(tot, prev); }
        live[gb] = 1u;
      }
      store_rw(buckets + gb, tot);
    }
  }
}


// -------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------
static void plan_chunks(MsmPlan& P) {
  // slices of L buckets for the running-sum reduction: enough threads to fill the chip
  const uint64_t total = (uint64_t)(P.merged ? 1 : P.W) * P.nb;
  uint32_t L = 64;
  while (L > 4 && total / L < 49152) L >>= 1;
  P.L = (int)std::min<uint32_t>(L, P.nb);
  P.nchunks = P.nb / P.L;
}

MsmPlan msm_plan(size_t n) {
  MsmPlan best{};
  double best_cost = 1e300;
  const char* env = getenv("GM_MSM_C");
  int forced = env ? atoi(env) : 0;
  for (int c = 4; c <= 22; c++) {
    if (forced && c != forced) continue;
    int W = (256 + c - 1) / c;
    double nb = (double)(1u << (c - 1));
    // Fq multiplications: bucket accumulation (10 / mixed add) + running sums (2 * 14 / bucket)
    // + a latency term for the serial tail of the reduction kernels
    double cost = (double)n * W * 10.0 + W * nb * 28.0 + 3.0e5 * W;
    if (cost < best_cost) {
      best_cost = cost;
      best.c = c; best.W = W; best.nb = 1u << (c - 1);
    }
  }
  best.merged = false;
  plan_chunks(best);
  return best;
}

// window size for a precomputed table: all windows share ONE bucket set, so the running-sum cost
// is paid once and larger windows (fewer gathered points per scalar) win
MsmPlan msm_plan_merged(size_t n, int c_forced) {
  MsmPlan best{};
  double best_cost = 1e300;
  const char* env = getenv("GM_MSM_C_PRE");
  int forced = c_forced ? c_forced : (env ? atoi(env) : 0);
  for (int c = 10; c <= 23; c++) {
    if (forced && c != forced) continue;
    int W = (256 + c - 1) / c;
    double nb = (double)(1u << (c - 1));
    double cost = (double)n * W * 10.0 + nb * 28.0;
    if (cost < best_cost) {
      best_cost = cost;
      best.c = c; best.W = W; best.nb = 1u << (c - 1);
    }
  }
  best.merged = true;
  plan_chunks(best);
  return best;
}

#define LAUNCH(ctx, kernel, grid, block, shmem, ...)                       \
  do {                                                                     \
    kernel<<<grid, block, shmem, (ctx)->stream>>>(__VA_ARGS__);            \
    (ctx)->launches++;                                                     \
  } while (0)

// Plan of one MSM pass over these bases: merged (one bucket set) when a precomputed table is present.
static MsmPlan plan_for(const MsmBases& B, size_t n) { return B.table != nullptr ? msm_plan_merged(n, B.c) : msm_plan(n); }

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

// Affine levels before the XYZZ accumulation.  GM_MSM_AFFINE = k forces k levels (0 = off, tests force them on tiny
// inputs); otherwise: none for small MSMs (each level costs three launches and one serial field inversion, about
// 0.1 ms), else as many as leave 4..8 points per bucket on average.
static int affine_levels(size_t refs, size_t M) {
  const char* env = getenv("GM_MSM_AFFINE");
  if (env && *env) { int v = atoi(env); if (v >= 0) return std::min(v, 8); }
  if (refs < ((size_t)1 << 22)) return 0;
  const size_t avg = refs / M;
  int r = 0;
  while (r < 8 && (avg >> (r + 1)) >= (size_t)env_int("GM_AFF_KEEP", 4)) r++;
  return r;
}

// Shape of one affine level: every thread walks G slots.  Long runs amortise the 10 products of the warp butterfly
// and mean fewer inversions; short runs give more warps to balance the SMs.  Measured on B200 (profiles/r01_summary.md):
// at least 64 warps per SM wins at 2^20 and 2^24; one or two big waves, or cutting a level in two parts to overlap
// the inversions on a second stream, were slower.  GM_AFF_WPS overrides the warps-per-SM floor.
struct AffShape {
  int G;
  uint32_t warps;   // a multiple of the 4 warps of a CTA
};
static AffShape aff_shape(const gm_ctx* ctx, size_t slots) {
  const size_t wps = (size_t)std::max(1, env_int("GM_AFF_WPS", 64));
  size_t G = slots / (32 * (size_t)ctx->sm_count * wps);
  G = std::min<size_t>(64, std::max<size_t>(4, G));
  const size_t warps = (slots + 32 * G - 1) / (32 * G);
  AffShape sh;
  sh.warps = (uint32_t)((warps + 3) / 4 * 4);
  sh.G = (int)G;
  return sh;
}

// Padded length of the sorted reference array when every bucket's run starts at a multiple of A = 2^levels: an upper bound
// from the host's point of view (each non-empty bucket pads by at most A - 1), itself a multiple of A.
static size_t padded_refs_bound(size_t refs, size_t M, int levels) {
  const size_t A = (size_t)1 << levels;
  return (refs + std::min(M, refs) * (A - 1) + A - 1) & ~(A - 1);
}

// HBM the affine levels of one pass need on top of the sort buffers: x / y planes of the first two level outputs,
// the prefix products and the pair kinds of the first (largest) level
static size_t affine_scratch_bytes(size_t refs, size_t M, int levels) {
  if (levels <= 0) return 0;
  const size_t S = padded_refs_bound(refs, M, levels);
  return (S >> 1) * (sizeof(Affine) + sizeof(Fq) + 1) + (levels > 1 ? (S >> 2) * sizeof(Affine) : 0);
}

// Phase A: digits, counting sort, affine levels, work list, bucket accumulation.  Buckets go to `buckets`; when `live` is
// given the buckets persist across calls (streamed MSM) and are updated in place, otherwise they are (re)written and the
// per-call counts (ctx->msm.counts) tell which ones are valid.
// `feed` (optional): the scalars are still in PINNED host memory.  They are copied in `feed->pieces` pieces on the context's
// copy stream, and the digits / histogram kernel of a piece starts as soon as that piece has landed: the conversion of the
// scalars and all the allocation-free set-up of the call hide behind the PCIe transfer, and the rest of the pipeline
// (scan, scatter, levels, ...) starts the moment the last piece is in.
struct HostFeed {
  const uint64_t* host = nullptr;
  int pieces = 1;
};
static int msm_sort_accumulate(gm_ctx* ctx, const MsmBases& B, size_t base_offset, const uint32_t* d_scalars, size_t n, bool bigint,
                               const MsmPlan& P, XYZZ* buckets, uint32_t* live, const HostFeed* feed = nullptr) {
  if (n == 0) return GM_OK;
  MsmScratch& S = ctx->msm;
  const bool merged = P.merged;
  const int Weff = merged ? 1 : P.W;            // bucket sets
  const size_t M = (size_t)Weff * P.nb;
  const Affine* d_bases = merged ? B.table : B.points + base_offset;
  const uint32_t ref_offset = merged ? (uint32_t)base_offset : 0u;
  const uint32_t ref_stride = merged ? (uint32_t)B.n : 0u;
  const int rec_q = merged ? B.rec_q : 6;
  const size_t refs = (size_t)P.W * n;

  // affine levels before the XYZZ accumulation (0 = none); their scratch is optional: without it the XYZZ path runs alone
  int levels = affine_levels(refs, M);
  if (levels > 0) {
    // the level scratch of a very large pass must not crowd out the buffers the XYZZ path needs anyway: skip the levels
    // when the part still to be allocated exceeds half of the free HBM (only looked at above 8 GB of scratch;
    // msm_accumulate sizes its passes so that this does not happen)
    const size_t need = affine_scratch_bytes(refs, M, levels);
    if (need > ((size_t)8 << 30)) {
      const size_t have = S.aff_a.cap + S.aff_b.cap + S.aff_prefix.cap + S.aff_meta.cap;
      size_t free_b = 0, total_b = 0;
      if (need > have && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && need - have > free_b / 2) levels = 0;
    }
  }
  size_t s_max = padded_refs_bound(refs, M, levels);
  if (levels > 0) {
    const size_t s1 = s_max >> 1, s2 = levels > 1 ? s_max >> 2 : 0;
    const size_t warps1 = (size_t)aff_shape(ctx, s1).warps + (size_t)ctx->sm_count * 64;   // the first level has the most warps
    const bool ok = S.aff_a.reserve(s1 * sizeof(Affine)) == GM_OK && S.aff_b.reserve(s2 * sizeof(Affine) + 16) == GM_OK &&
                    S.aff_prefix.reserve(s1 * sizeof(Fq)) == GM_OK && S.aff_meta.reserve(s1 + 16) == GM_OK &&
                    S.aff_others.reserve(warps1 * 32 * sizeof(Fq)) == GM_OK && S.aff_totals.reserve(warps1 * sizeof(Fq)) == GM_OK;
    if (!ok) {
      S.aff_a.release(); S.aff_b.release(); S.aff_prefix.release(); S.aff_meta.release();
      cudaGetLastError();
      levels = 0;
      s_max = refs;
    }
  }
  const uint32_t pad = (1u << levels) - 1u;
  const size_t refs_eff = s_max >> levels;   // upper bound of the points the XYZZ accumulation still has to add

  GM_TRY(S.digits.reserve(refs * 4));
  GM_TRY(S.sorted.reserve(s_max * 4 + 16));
  GM_TRY(S.counts.reserve(M * 4));
  GM_TRY(S.starts.reserve(M * 4));
  GM_TRY(S.cursor.reserve(M * 4));
  GM_TRY(S.poff.reserve(M * 4));
  const size_t ntiles = (M + SCAN_TILE - 1) / SCAN_TILE;
  GM_TRY(S.scan_tmp.reserve(ntiles * 4 + 16));
  static_assert(sizeof(Meta) <= 16384, "Meta fits its slot");
  GM_TRY(S.meta.reserve(16384));
  Meta* meta = S.meta.as<Meta>();

  // references per work item: enough items to keep every SM busy with several waves, but long enough that the
  // per-item partial sums of a hot bucket stay few
  uint32_t split = SPLIT;
  const size_t avg_load = refs_eff / M + 1;
  while (split > 32 && split / 2 >= 4 * avg_load && refs_eff / split < (size_t)ctx->sm_count * 384 * 4) split >>= 1;
  const size_t max_split = refs_eff / split + 1;
  const size_t max_partials = 2 * max_split + 1;
  const size_t max_items = std::min<size_t>(M, refs_eff) + max_split + 1;
  GM_TRY(S.partials.reserve(max_partials * sizeof(XYZZ)));
  GM_TRY(S.work.reserve(max_items * sizeof(uint2)));
  GM_TRY(S.split.reserve(max_split * 4));

  cudaStream_t st = ctx->stream;
  GM_CUDA(cudaMemsetAsync(S.counts.p, 0, M * 4, st));
  GM_CUDA(cudaMemsetAsync(meta, 0, sizeof(Meta), st));
  // the padding between the buckets' runs reads as NONE (only the levels look at it)
  if (levels > 0) GM_CUDA(cudaMemsetAsync(S.sorted.p, 0xFF, s_max * 4, st));

  const uint32_t n32 = (uint32_t)n;
  const uint32_t M32 = (uint32_t)M;
  const uint32_t* counts = S.counts.as<uint32_t>();
  const uint32_t* starts = S.starts.as<uint32_t>();
  GM_CUDA(cudaEventRecord(ctx->ev[2], st));
  if (feed != nullptr && feed->pieces > 1) {
    // the copy stream starts after everything queued on the main stream so far (the previous call may still read d_scalars)
    GM_CUDA(cudaEventRecord(ctx->ev_join, st));
    GM_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_join, 0));
    const size_t step = (((n + feed->pieces - 1) / feed->pieces) + 255) & ~(size_t)255;
    for (size_t off = 0; off < n; off += step) {
      const size_t m = std::min(step, n - off);
      GM_CUDA(cudaMemcpyAsync(const_cast<uint32_t*>(d_scalars) + off * 8, feed->host + off * 4, m * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
      GM_CUDA(cudaEventRecord(ctx->ev_join, ctx->copy_stream));
      GM_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
      LAUNCH(ctx, k_digits_hist, (unsigned)((m + 255) / 256), 256, 0, d_scalars, (uint32_t)off, (uint32_t)(off + m), n32, S.scalar_stride, bigint ? 1 : 0, P.c,
             P.W, merged ? 1 : 0, S.digits.as<uint32_t>(), S.counts.as<uint32_t>());
    }
  } else {
    LAUNCH(ctx, k_digits_hist, (n32 + 255) / 256, 256, 0, d_scalars, 0u, n32, n32, S.scalar_stride, bigint ? 1 : 0, P.c, P.W, merged ? 1 : 0,
           S.digits.as<uint32_t>(), S.counts.as<uint32_t>());
  }
  LAUNCH(ctx, k_scan_tiles, (unsigned)ntiles, SCAN_THREADS, 0, counts, S.starts.as<uint32_t>(), S.scan_tmp.as<uint32_t>(), M32, pad);
  LAUNCH(ctx, k_scan_tile_sums, 1, 1024, 0, S.scan_tmp.as<uint32_t>(), (uint32_t)ntiles);
  LAUNCH(ctx, k_scan_add, (unsigned)((M + 255) / 256), 256, 0, S.starts.as<uint32_t>(), S.cursor.as<uint32_t>(), S.scan_tmp.as<uint32_t>(), M32);
  LAUNCH(ctx, k_scatter, dim3((n32 + 255) / 256, P.W), 256, 0, S.digits.as<uint32_t>(), n32, P.W, P.nb, merged ? 1 : 0, ref_offset, ref_stride, S.cursor.as<uint32_t>(), S.sorted.as<uint32_t>());
  GM_CUDA(cudaEventRecord(ctx->ev[3], st));

  // ---- affine levels: level r pairs positions (2s, 2s+1) of level r-1 (level 0: of the sorted references) ----
  Fq* cur_x = nullptr;
  Fq* cur_y = nullptr;
  for (int r = 0; r < levels; r++) {
    const size_t slots = s_max >> (r + 1);
    DevBuf& ob = (r & 1) ? S.aff_b : S.aff_a;
    Fq* out_x = ob.as<Fq>();
    Fq* out_y = out_x + slots;
    const AffShape shp = aff_shape(ctx, slots);
    const int G = shp.G;
    const unsigned ctas = shp.warps * 32 / AFF_THREADS;
    // few inversions: one per single-warp CTA (lane 0 only: data-dependent loops do not diverge)
    const int spread = shp.warps <= (uint32_t)ctx->sm_count * 24 ? 32 : 1;
    const unsigned inv_threads = spread == 32 ? 32u : 128u;
    const unsigned inv_ctas = (unsigned)(((size_t)shp.warps * spread + inv_threads - 1) / inv_threads);
    Fq* prefix = S.aff_prefix.as<Fq>();
    uint8_t* kinds = S.aff_meta.as<uint8_t>();
    Fq* others = S.aff_others.as<Fq>();
    Fq* totals = S.aff_totals.as<Fq>();
    AffIn in;
    in.table = d_bases; in.refs = S.sorted.as<uint32_t>(); in.x = cur_x; in.y = cur_y; in.rec_q = rec_q;
    if (r == 0)
      LAUNCH(ctx, k_aff_prepare<true>, ctas, AFF_THREADS, 0, in, starts, counts, M32, pad, r, G, prefix, kinds, others, totals);
    else
      LAUNCH(ctx, k_aff_prepare<false>, ctas, AFF_THREADS, 0, in, starts, counts, M32, pad, r, G, prefix, kinds, others, totals);
    LAUNCH(ctx, k_aff_invert, inv_ctas, inv_threads, 0, starts, counts, M32, pad, r, G, spread, totals);
    if (r == 0)
      LAUNCH(ctx, k_aff_finish<true>, ctas, AFF_THREADS, 0, in, starts, counts, M32, pad, r, G, prefix, kinds, others, totals, out_x, out_y);
    else
      LAUNCH(ctx, k_aff_finish<false>, ctas, AFF_THREADS, 0, in, starts, counts, M32, pad, r, G, prefix, kinds, others, totals, out_x, out_y);
    cur_x = out_x; cur_y = out_y;
  }

  LAUNCH(ctx, k_classify, (unsigned)((M + 255) / 256), 256, 0, counts, M32, levels, split, S.poff.as<uint32_t>(), S.split.as<uint32_t>(), meta);
  LAUNCH(ctx, k_size_scan, 1, 32, 0, meta, split);
  LAUNCH(ctx, k_worklist_fill, (unsigned)((M + 255) / 256), 256, 0, counts, M32, levels, split, S.work.as<uint2>(), meta);
  const unsigned acc_ctas = (unsigned)((max_items + ACC_THREADS - 1) / ACC_THREADS);
  if (levels > 0)
    LAUNCH(ctx, k_accumulate<true>, acc_ctas, ACC_THREADS, 0, (const Affine*)nullptr, cur_x, cur_y, (const uint32_t*)nullptr, counts, starts, levels,
           S.poff.as<uint32_t>(), S.work.as<uint2>(), meta, split, rec_q, buckets, S.partials.as<XYZZ>(), live);
  else
    LAUNCH(ctx, k_accumulate<false>, acc_ctas, ACC_THREADS, 0, d_bases, (const Fq*)nullptr, (const Fq*)nullptr, S.sorted.as<uint32_t>(), counts, starts, 0,
           S.poff.as<uint32_t>(), S.work.as<uint2>(), meta, split, rec_q, buckets, S.partials.as<XYZZ>(), live);
  GM_CUDA(cudaEventRecord(ctx->ev[4], st));
  const size_t red_sh = RED_THREADS * sizeof(XYZZ);
  LAUNCH(ctx, k_split_combine, (unsigned)std::min<size_t>(max_split, (size_t)ctx->sm_count * 4), RED_THREADS, red_sh, S.split.as<uint32_t>(),
         counts, levels, S.poff.as<uint32_t>(), meta, split, S.partials.as<XYZZ>(), buckets, live);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

// Largest pass (a power-of-two number of terms, at most 2^27: W * n references are addressed with 31 bits + sign) whose
// sort buffers and affine-level scratch fit in the HBM that is free now plus what the scratch arena already holds.
static size_t msm_pass_size(gm_ctx* ctx, const MsmBases& B, size_t n) {
  size_t m = (size_t)1 << 27;
  if (n <= ((size_t)1 << 24)) return std::min(n, m);     // up to 2^24 terms always fit next to their table (19 GB)
  const MsmScratch& S = ctx->msm;
  const size_t have = S.digits.cap + S.sorted.cap + S.aff_a.cap + S.aff_b.cap + S.aff_prefix.cap + S.aff_meta.cap;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return std::min(n, m);
  const size_t budget = have + free_b - std::min<size_t>(free_b, (size_t)4 << 30);   // keep 4 GB of head room
  while (m > ((size_t)1 << 22)) {
    const MsmPlan P = B.table != nullptr ? msm_plan_merged(std::min(m, n), B.c) : msm_plan(std::min(m, n));
    const size_t M = (size_t)(P.merged ? 1 : P.W) * P.nb;
    const size_t refs = (size_t)P.W * std::min(m, n);
    const int levels = affine_levels(refs, M);
    const size_t need = refs * 4 + padded_refs_bound(refs, M, levels) * 4 + affine_scratch_bytes(refs, M, levels);
    if (need <= budget) break;
    m >>= 1;
  }
  return std::min(n, m);
}


int msm_accumulate(gm_ctx* ctx, const MsmBases& B, size_t base_offset, const uint32_t* d_scalars, size_t n, bool bigint, XYZZ* d_acc) {
  // very large inputs run as several passes (reference width, HBM for the level scratch); each pass reduces its buckets
  const size_t max_pass = msm_pass_size(ctx, B, n);
  for (size_t off = 0; off < n; off += max_pass) {
    const size_t m = std::min(max_pass, n - off);
    const MsmPlan P = plan_for(B, m);
    const size_t M = (size_t)(P.merged ? 1 : P.W) * P.nb;
    GM_TRY(ctx->msm.buckets.reserve(M * sizeof(XYZZ)));
    GM_TRY(msm_sort_accumulate(ctx, B, base_offset + off, d_scalars + off * 8 * ctx->msm.scalar_stride, m, bigint, P, ctx->msm.buckets.as<XYZZ>(), nullptr));
    GM_TRY(msm_reduce(ctx, P, ctx->msm.buckets.as<XYZZ>(), ctx->msm.counts.as<uint32_t>(), d_acc));
  }
  return GM_OK;
}

// *d_acc += sum_i scalars[i] * bases[i] with the scalars in PINNED host memory (see HostFeed); GM_ERR_ARG when the input
// needs several passes (the caller then falls back to the streamed path)
int msm_accumulate_pinned(gm_ctx* ctx, const MsmBases& B, size_t base_offset, const uint64_t* h_scalars, size_t n, bool bigint, int pieces,
                          XYZZ* d_acc) {
  if (n == 0) return GM_OK;
  if (n > msm_pass_size(ctx, B, n) || ctx->msm.scalar_stride != 1) return GM_ERR_ARG;
  GM_TRY(ctx->msm.scalars.reserve(n * 32));
  const MsmPlan P = plan_for(B, n);
  const size_t M = (size_t)(P.merged ? 1 : P.W) * P.nb;
  GM_TRY(ctx->msm.buckets.reserve(M * sizeof(XYZZ)));
  HostFeed feed;
  feed.host = h_scalars;
  feed.pieces = pieces;
  GM_TRY(msm_sort_accumulate(ctx, B, base_offset, ctx->msm.scalars.as<uint32_t>(), n, bigint, P, ctx->msm.buckets.as<XYZZ>(), nullptr, &feed));
  return msm_reduce(ctx, P, ctx->msm.buckets.as<XYZZ>(), ctx->msm.counts.as<uint32_t>(), d_acc);
}

// ---- streamed MSM with device-resident buckets: chunks only sort + accumulate, the reduction runs once ----
MsmPlan msm_stream_plan(const MsmBases& B, size_t chunk_cap) { return plan_for(B, std::max<size_t>(chunk_cap, 1)); }

size_t msm_plan_buckets(const MsmPlan& P) { return (size_t)(P.merged ? 1 : P.W) * P.nb; }

int msm_stream_push(gm_ctx* ctx, const MsmBases& B, size_t base_offset, const uint32_t* d_scalars, size_t n, bool bigint,
                    const MsmPlan& P, XYZZ* d_buckets, uint32_t* d_live) {
  const size_t max_pass = (size_t)1 << 27;
  for (size_t off = 0; off < n; off += max_pass) {
    const size_t m = std::min(max_pass, n - off);
    GM_TRY(msm_sort_accumulate(ctx, B, base_offset + off, d_scalars + off * 8, m, bigint, P, d_buckets, d_live));
  }
  return GM_OK;
}

int msm_stream_reduce(gm_ctx* ctx, const MsmPlan& P, const XYZZ* d_buckets, const uint32_t* d_live, XYZZ* d_acc) {
  return msm_reduce(ctx, P, d_buckets, d_live, d_acc);
}

// Planning decisions as plain numbers (host only, no device needed): window bits, windows, buckets per set, affine
// levels and the (G, warps) shape of the first level - what tests/test_host_logic.py pins.
void msm_describe_plan(size_t n, bool with_table, int sm_count, int out[8]) {
  const MsmPlan P = with_table ? msm_plan_merged(n, 0) : msm_plan(n);
  const size_t M = (size_t)(P.merged ? 1 : P.W) * P.nb;
  const size_t refs = (size_t)P.W * n;
  const int levels = affine_levels(refs, M);
  gm_ctx fake;
  fake.sm_count = sm_count;
  const size_t s1 = padded_refs_bound(refs, M, levels) >> 1;
  const AffShape sh = aff_shape(&fake, s1);
  out[0] = P.c; out[1] = P.W; out[2] = (int)P.nb; out[3] = P.merged ? 1 : 0; out[4] = levels;
  out[5] = sh.G; out[6] = (int)sh.warps; out[7] = P.L;
}









}  // namespace gm
