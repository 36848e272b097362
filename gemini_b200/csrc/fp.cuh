// Montgomery prime-field arithmetic on 32-bit limbs (BLS12-381 Fq: 12 limbs,
// Fr: 8 limbs).  Elements live in Montgomery form, fully reduced (< p), in the
// same little-endian limb order arkworks' Fp<MontBackend, N> keeps in memory
// (ark-ff 0.4.2, pinned by /root/reference/Cargo.lock:62-64), so host buffers
// are consumed and produced without any conversion.
//
// mont_mul is an operand-scanning (CIOS) product whose partial sums are split
// into an even- and an odd-column accumulator: every 64-bit partial product
// a[j]*w then lands on an aligned (lo,hi) pair of one accumulator, which lets a
// whole row run as one carry chain of mad.lo.cc/madc.hi.cc pairs (IMAD.WIDE.X
// in SASS) with no carry fix-up instructions.
#pragma once
#include "limbs.cuh"

namespace gm {

// ---------------------------------------------------------------------------
// Field parameters: compile-time limb tables (constexpr switch => immediates
// after unrolling; no constant-memory traffic, identical on host and device).
// ---------------------------------------------------------------------------
struct FqParams {
  static constexpr int N = 12;
  static constexpr bool LOW_LIMBS_1_FFFFFFFF = false;
  static constexpr uint32_t INV = 0xfffcfffdu;  // -q^{-1} mod 2^32
  GM_HD static constexpr uint32_t inv() { return INV; }
  GM_HD static constexpr uint32_t mod(int j) {
    constexpr uint32_t t[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
    return t[j];
  }
  // R = 2^384 mod q (Montgomery one)
  GM_HD static constexpr uint32_t one(int j) {
    constexpr uint32_t t[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
    return t[j];
  }
  // R^2 mod q
  GM_HD static constexpr uint32_t r2(int j) {
    constexpr uint32_t t[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
    return t[j];
  }
};

// -r^{-1} mod 2^32 is 0xffffffff, so the Montgomery quotient digit is m = -E[0].  When ptxas sees that negation it
// rewrites the products m * r_j and un-fuses every (mad.lo.cc, madc.hi.cc) pair of the reduction rows into
// IMAD + IMAD.HI (48 + 64 instructions instead of 64 IMAD.WIDE; IMAD.HI runs at 26 lanes/clk/SM).  Reading the
// constant from the constant bank keeps m opaque: one extra IMAD per row, all rows fused (checked with cuobjdump).
#if defined(__CUDACC__)
static __constant__ uint32_t FR_INV_BANK = 0xffffffffu;
#endif

struct FrParams {
  static constexpr int N = 8;
  static constexpr bool LOW_LIMBS_1_FFFFFFFF = true;   // mod(0) = 1, mod(1) = 0xffffffff, inv = -1: see mont_reduce_step
  static constexpr uint32_t INV = 0xffffffffu;  // -r^{-1} mod 2^32
  GM_HD static uint32_t inv() {
#if defined(__CUDA_ARCH__)
    return FR_INV_BANK;
#else
    return INV;
#endif
  }
  GM_HD static constexpr uint32_t mod(int j) {
    constexpr uint32_t t[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                               0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    return t[j];
  }
  GM_HD static constexpr uint32_t one(int j) {
    constexpr uint32_t t[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                               0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
    return t[j];
  }
  GM_HD static constexpr uint32_t r2(int j) {
    constexpr uint32_t t[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                               0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
    return t[j];
  }
};

// ---------------------------------------------------------------------------
// Raw limb routines
// ---------------------------------------------------------------------------
namespace detail {

// First row: (E, O) = a * w, E = even-column accumulator (E[k] at column k),
// O = odd-column accumulator (O[k] at column k+1).
template <int N>
GM_HD void row_first(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t w) {
#pragma unroll
  for (int j = 0; j < N; j += 2) mul_wide(E[j], E[j + 1], a[j], w);
#pragma unroll
  for (int j = 1; j < N; j += 2) mul_wide(O[j - 1], O[j], a[j], w);
}

// Montgomery step: add m*p with m chosen so that column 0 (E[0]) becomes zero.
// P::LOW_LIMBS_1_FFFFFFFF (Fr: r = ... ffffffff 00000001): m * p_0 = m and m * p_1 = m 2^32 - m need no multiplier -
// with m = -E[0] the low word of the latter is the old E[0] itself - so the first product of either chain becomes two
// additions on the ALU pipe: 6 IMAD.WIDE per row instead of 8.
template <class P>
GM_HD void mont_reduce_step(uint32_t* E, uint32_t* O) {
  constexpr int N = P::N;
  const uint32_t m = E[0] * P::inv();
  if constexpr (P::LOW_LIMBS_1_FFFFFFFF) {
    const uint32_t hi = m - (m != 0u ? 1u : 0u);   // (m 2^32 - m) >> 32
    O[0] = add_cc(O[0], E[0]);                      // (m 2^32 - m) mod 2^32 = -m = E[0]
    O[1] = addc_cc(O[1], hi);
  } else {
    mad_wide_cc(O[0], O[1], m, P::mod(1), O[0], O[1]);
  }
#pragma unroll
  for (int j = 3; j < N; j += 2) madc_wide_cc(O[j - 1], O[j], m, P::mod(j), O[j - 1], O[j]);
  // (no carry out of the odd chain: the running value is < 2^(32(N+1)))
  if constexpr (P::LOW_LIMBS_1_FFFFFFFF) {
    E[0] = add_cc(E[0], m);                         // = 0, carry = (m != 0)
    E[1] = addc_cc(E[1], 0);
  } else {
    mad_wide_cc(E[0], E[1], m, P::mod(0), E[0], E[1]);
  }
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], m, P::mod(j), E[j], E[j + 1]);
  O[N - 1] = addc(O[N - 1], 0);  // even chain's carry sits at column N
}

// Divide by 2^32 (drop the now-zero column 0) and add the next row a*w.
// On entry E/O are the even/odd accumulators; on exit the roles are swapped
// (the caller passes them swapped to the next step).
template <int N>
GM_HD void row_next(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t w) {
  O[0] = add_cc(O[0], E[1]);  // stray limb of the old even accumulator
#pragma unroll
  for (int j = 1; j < N - 1; j += 2) madc_wide_cc(E[j - 1], E[j], a[j], w, E[j + 1], E[j + 2]);
  madc_wide_top(E[N - 2], E[N - 1], a[N - 1], w);
  mad_wide_cc(O[0], O[1], a[0], w, O[0], O[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(O[j], O[j + 1], a[j], w, O[j], O[j + 1]);
  E[N - 1] = addc(E[N - 1], 0);
}

// r = (a >= p) ? a - p : a
template <class P>
GM_HD void cond_sub_p(uint32_t* r, const uint32_t* a) {
  constexpr int N = P::N;
  uint32_t t[N];
  t[0] = sub_cc(a[0], P::mod(0));
#pragma unroll
  for (int j = 1; j < N; j++) t[j] = subc_cc(a[j], P::mod(j));
  const uint32_t borrow = subc(0, 0);  // 0xffffffff if a < p
#pragma unroll
  for (int j = 0; j < N; j++) r[j] = borrow ? a[j] : t[j];
}

}  // namespace detail

namespace detail {

// t[0..2N) = a * b, N even: the rows of the CIOS schedule without its reduction steps - the lowest column of the
// running sum is final after every row and leaves through t[i] where the CIOS would have made it zero.
template <int N>
GM_HD void mul_full(uint32_t* t, const uint32_t* a, const uint32_t* b) {
  uint32_t x[N], y[N];
  row_first<N>(x, y, a, b[0]);
  t[0] = x[0];
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    row_next<N>(x, y, a, b[i]);  // even = y, odd = x
    t[i] = y[0];
    if (i + 1 < N) {
      row_next<N>(y, x, a, b[i + 1]);
      t[i + 1] = x[0];
    }
  }
  // even = y (y[0] already emitted), odd = x: column k of what is left = y[k] + x[k-1]
  t[N] = add_cc(x[0], y[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) t[N + k] = addc_cc(x[k], y[k + 1]);
  t[2 * N - 1] = addc(x[N - 1], 0);
}

}  // namespace detail

template <class P>
GM_HD void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  uint32_t x[N], y[N];
  detail::row_first<N>(x, y, a, b[0]);
  detail::mont_reduce_step<P>(x, y);
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    detail::row_next<N>(x, y, a, b[i]);  // roles swap: even = y, odd = x
    detail::mont_reduce_step<P>(y, x);
    if (i + 1 < N) {
      detail::row_next<N>(y, x, a, b[i + 1]);  // roles swap back
      detail::mont_reduce_step<P>(x, y);
    }
  }
  // N is even: after the loop the even accumulator is y, the odd one is x.
  // result = (even + odd * 2^32) / 2^32 -> limbs odd[k] + even[k+1]
  uint32_t t[N];
  t[0] = add_cc(x[0], y[1]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) t[k] = addc_cc(x[k], y[k + 1]);
  t[N - 1] = addc(x[N - 1], 0);
  detail::cond_sub_p<P>(r, t);
}

// Montgomery reduction of a single N-limb value: r = a * R^{-1} mod p
// (used for into_bigint: Montgomery -> canonical).
template <class P>
GM_HD void mont_redc(uint32_t* r, const uint32_t* a) {
  constexpr int N = P::N;
  uint32_t one[N];
#pragma unroll
  for (int j = 0; j < N; j++) one[j] = (j == 0) ? 1u : 0u;
  mont_mul<P>(r, a, one);
}

template <class P>
GM_HD void fp_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  uint32_t t[N];
  t[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int j = 1; j < N - 1; j++) t[j] = addc_cc(a[j], b[j]);
  t[N - 1] = addc(a[N - 1], b[N - 1]);  // 2p < 2^(32N): no carry out
  detail::cond_sub_p<P>(r, t);
}

template <class P>
GM_HD void fp_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  uint32_t t[N];
  t[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int j = 1; j < N; j++) t[j] = subc_cc(a[j], b[j]);
  const uint32_t borrow = subc(0, 0);  // all-ones if a < b
  r[0] = add_cc(t[0], P::mod(0) & borrow);
#pragma unroll
  for (int j = 1; j < N - 1; j++) r[j] = addc_cc(t[j], P::mod(j) & borrow);
  r[N - 1] = addc(t[N - 1], P::mod(N - 1) & borrow);
}

// ---------------------------------------------------------------------------
// Value type
// ---------------------------------------------------------------------------
template <class P>
struct Fp {
  using Params = P;
  static constexpr int N = P::N;
  uint32_t v[N];

  GM_HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int j = 0; j < N; j++) r.v[j] = 0;
    return r;
  }
  GM_HD static Fp one() {
    Fp r;
#pragma unroll
    for (int j = 0; j < N; j++) r.v[j] = P::one(j);
    return r;
  }
  GM_HD bool is_zero() const {
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < N; j++) acc |= v[j];
    return acc == 0;
  }
  GM_HD bool operator==(const Fp& o) const {
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < N; j++) acc |= v[j] ^ o.v[j];
    return acc == 0;
  }
  GM_HD bool operator!=(const Fp& o) const { return !(*this == o); }
  GM_HD Fp operator+(const Fp& o) const { Fp r; fp_add<P>(r.v, v, o.v); return r; }
  GM_HD Fp operator-(const Fp& o) const { Fp r; fp_sub<P>(r.v, v, o.v); return r; }
  GM_HD Fp operator*(const Fp& o) const { Fp r; mont_mul<P>(r.v, v, o.v); return r; }
  GM_HD Fp sqr() const { Fp r; mont_mul<P>(r.v, v, v); return r; }
  GM_HD Fp dbl() const { Fp r; fp_add<P>(r.v, v, v); return r; }
  GM_HD Fp neg() const { Fp z = zero(); Fp r; fp_sub<P>(r.v, z.v, v); return r; }
  GM_HD Fp from_mont() const { Fp r; mont_redc<P>(r.v, v); return r; }
  GM_HD Fp to_mont() const {
    Fp k;
#pragma unroll
    for (int j = 0; j < N; j++) k.v[j] = P::r2(j);
    return *this * k;
  }
};

// ---------------------------------------------------------------------------
// Lazy reduction for sums of products (the inner products of the sumcheck messages in fr.cu, like ark-ff's sum_of_products
// behind misc::ip_unsafe): acc += a*b keeps the full 2N-limb product and only folds the TOP half back below p (one
// N-limb conditional subtraction), so a product costs N^2 IMAD.WIDE instead of the 2 N^2 of a Montgomery product;
// one Montgomery reduction at the very end.  Invariant: acc < p * 2^(32N)  (top half < p).  Needs 2 bits of slack
// in the top limb of p (a*b < p^2 < p 2^(32N) / 4 ... true for Fr: r < 2^255).
// ---------------------------------------------------------------------------
template <class P>
struct FpAcc {
  static constexpr int N = P::N;
  uint32_t v[2 * N];

  GM_HD static FpAcc zero() {
    FpAcc r;
#pragma unroll
    for (int j = 0; j < 2 * N; j++) r.v[j] = 0;
    return r;
  }
  // acc += a * b WITHOUT touching the invariant: the product comes from the row schedule of the Montgomery product
  // (detail::mul_full: N^2 IMAD.WIDE.X on aligned carry chains, ~2N additions) and is added with one 2N-limb carry chain -
  // a third of the ALU instructions of column-wise product scanning (measured, round 2: the message kernels spent
  // 450 of 628 instructions per pair on the three-word column accumulators).  Bound: with acc < p 2^(32N) on entry, k products
  // leave acc < p 2^(32N) + k p^2, which is < 2 p 2^(32N) (no overflow, and ONE conditional subtraction restores the
  // invariant) as long as k p < 2^(32N): k <= 2 for Fr (2^256 / r = 2.2), k <= 9 for Fq.
  static constexpr int UNREDUCED_RUN = (N == 8) ? 2 : 9;
  GM_HD void mul_add_unreduced(const Fp<P>& a, const Fp<P>& b) {
    uint32_t t[2 * N];
    detail::mul_full<N>(t, a.v, b.v);
    v[0] = add_cc(v[0], t[0]);
#pragma unroll
    for (int k = 1; k < 2 * N - 1; k++) v[k] = addc_cc(v[k], t[k]);
    v[2 * N - 1] = addc(v[2 * N - 1], t[2 * N - 1]);
  }
  // after at most UNREDUCED_RUN calls of mul_add_unreduced: top half back below p
  GM_HD void normalize() { detail::cond_sub_p<P>(v + N, v + N); }
  // acc += a * b   (a, b < p), invariant kept
  GM_HD void mul_add(const Fp<P>& a, const Fp<P>& b) { mul_add_unreduced(a, b); normalize(); }
  // (acc / 2^(32N)) mod p as a field element: top half + REDC(bottom half)
  GM_HD Fp<P> reduce() const {
    Fp<P> lo, hi, one;
#pragma unroll
    for (int j = 0; j < N; j++) { lo.v[j] = v[j]; hi.v[j] = v[N + j]; one.v[j] = (j == 0) ? 1u : 0u; }
    Fp<P> t;
    mont_mul<P>(t.v, one.v, lo.v);   // lo may exceed p: it is the row operand (a < p, b < R is all CIOS needs)
    return hi + t;
  }
};

using Fq = Fp<FqParams>;
using Fr = Fp<FrParams>;
using FrAcc = FpAcc<FrParams>;

// ---------------------------------------------------------------------------
// Inversion.  fp_inv_fermat: a^(p-2), ~570 Montgomery products.  fp_inv: Kaliski's "almost Montgomery
// inverse" (binary extended Euclid: shifts, compares and subtractions on N-limb integers, no products)
// followed by the power-of-two correction - about 5x fewer instructions, which matters because the
// inversion sits on the serial tail of every MSM (normalisation of the result to Z = 1).
// ---------------------------------------------------------------------------
template <class P>
GM_HD Fp<P> fp_inv_fermat(const Fp<P>& a) {
  constexpr int N = P::N;
  uint32_t e[N];
  uint32_t borrow = 0;
  for (int j = 0; j < N; j++) {
    uint64_t t = (uint64_t)P::mod(j) - (j == 0 ? 2u : 0u) - borrow;
    e[j] = (uint32_t)t;
    borrow = (uint32_t)(t >> 63);
  }
  Fp<P> acc = Fp<P>::one();
#pragma unroll 1
  for (int bit = 32 * N - 1; bit >= 0; bit--) {
    acc = acc.sqr();
    if ((e[bit >> 5] >> (bit & 31)) & 1u) acc = acc * a;
  }
  return acc;
}

namespace detail {
template <int N> GM_HD bool limbs_is_zero(const uint32_t* x) {
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < N; j++) acc |= x[j];
  return acc == 0;
}
template <int N> GM_HD void limbs_shr1(uint32_t* x) {
#pragma unroll
  for (int j = 0; j < N - 1; j++) x[j] = (x[j] >> 1) | (x[j + 1] << 31);
  x[N - 1] >>= 1;
}
template <int N> GM_HD void limbs_shl1(uint32_t* x) {
#pragma unroll
  for (int j = N - 1; j > 0; j--) x[j] = (x[j] << 1) | (x[j - 1] >> 31);
  x[0] <<= 1;
}
// r = a - b, returns true if a < b (borrow)
template <int N> GM_HD bool limbs_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int j = 1; j < N; j++) r[j] = subc_cc(a[j], b[j]);
  return subc(0, 0) != 0;
}
template <int N> GM_HD void limbs_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  r[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int j = 1; j < N - 1; j++) r[j] = addc_cc(a[j], b[j]);
  r[N - 1] = addc(a[N - 1], b[N - 1]);
}
}  // namespace detail

// a != 0, a in Montgomery form; returns a^{-1} in Montgomery form.
template <class P>
GM_HD Fp<P> fp_inv(const Fp<P>& a) {
  constexpr int N = P::N;
  uint32_t u[N], v[N], r[N], s[N], t[N];
#pragma unroll
  for (int j = 0; j < N; j++) { u[j] = P::mod(j); v[j] = a.v[j]; r[j] = 0; s[j] = (j == 0) ? 1u : 0u; }
  int k = 0;
  // invariant: p = u*s + v*r (so r, s <= p while u, v >= 1; they fit N limbs since 2p < 2^(32N))
#pragma unroll 1
  while (!detail::limbs_is_zero<N>(v) && k < 64 * N) {  // k <= 2 * bits(p) < 64N: the bound only guards against a hang
    if ((u[0] & 1u) == 0) {
      detail::limbs_shr1<N>(u); detail::limbs_shl1<N>(s);
    } else if ((v[0] & 1u) == 0) {
      detail::limbs_shr1<N>(v); detail::limbs_shl1<N>(r);
    } else {
      const bool v_lt_u = detail::limbs_sub<N>(t, v, u);   // t = v - u
      if (v_lt_u) {                                         // u > v
        detail::limbs_sub<N>(u, u, v);
        detail::limbs_shr1<N>(u);
        detail::limbs_add<N>(r, r, s);
        detail::limbs_shl1<N>(s);
      } else {                                              // v >= u (v == u ends the loop: v becomes 0)
#pragma unroll
        for (int j = 0; j < N; j++) v[j] = t[j];
        detail::limbs_shr1<N>(v);
        detail::limbs_add<N>(s, s, r);
        detail::limbs_shl1<N>(r);
      }
    }
    k++;
  }
  // r = -a^{-1} 2^k (mod p), r < 2p
  detail::cond_sub_p<P>(r, r);
  Fp<P> x;
  {
    uint32_t pm[N];
#pragma unroll
    for (int j = 0; j < N; j++) pm[j] = P::mod(j);
    detail::limbs_sub<N>(x.v, pm, r);   // p - r in (0, p]
    detail::cond_sub_p<P>(x.v, x.v);
  }
  // x = (aR)^{-1} 2^k = a'^{-1} 2^(k - 32N); the Montgomery form of the inverse is a'^{-1} 2^(32N): multiply by
  // 2^m, m = 64N - k (k >= bits(p), so m < 32N + a few).  2^sh as a plain integer times R^2 does it in two
  // products, mont(mont(x, 2^sh), R^2) = x 2^sh, instead of up to 32N modular doublings on the serial tail.
  const int m = 64 * N - k;
  const int sh = m < 32 * N - 1 ? m : 32 * N - 1;
  Fp<P> e, r2;
#pragma unroll
  for (int j = 0; j < N; j++) { e.v[j] = (j == (sh >> 5)) ? (1u << (sh & 31)) : 0u; r2.v[j] = P::r2(j); }
  x = (x * e) * r2;   // e is not reduced (it may exceed p) but it is the row operand: a < p, b < R is all CIOS needs
#pragma unroll 1
  for (int i = sh; i < m; i++) x = x.dbl();
  return x;
}

}  // namespace gm
