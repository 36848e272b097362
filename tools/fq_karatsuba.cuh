// EXPERIMENT (measured on B200, not used by the library): Karatsuba / separated-operand-scanning Montgomery product.
//   tools/mul_microbench.cu, round 2:  CIOS 29.60 G Fq-mul/s | one Karatsuba level 30.49 G/s (303 -> 278 multiplier-pipe
//   instructions per product in SASS, 69 -> 223 ALU instructions) | two levels 26.77 G/s (ptxas turns the extra additions
//   into IMAD.X / IMAD.MOV on the same pipe).  Inside the MSM kernels the larger live set costs more than the saved
//   products: k_aff_finish spills 110-140 bytes per thread at its 128-register cap and the whole MSM went from 66.6 to
//   75.7 ms at 2^24 (6.25 -> 7.14 ms at 2^20) with -DGM_KARATSUBA.  Kept here with its host tests
//   (tests/test_host_field.py) as the record of that measurement.
#pragma once
#include "../gemini_b200/csrc/fp.cuh"

namespace gm {

// ---------------------------------------------------------------------------
// Karatsuba variant of the Montgomery product (separated operand scanning): T = a*b from THREE half-size
// products instead of four (3 (N/2)^2 IMAD.WIDE instead of N^2), then the Montgomery reduction of the low half of T.
// The multiplier pipe (IMAD.WIDE at 32 lanes/clk/SM) is the roofline of every MSM kernel; the additions Karatsuba
// adds run on the ALU pipe, which those kernels leave idle.
// ---------------------------------------------------------------------------
namespace detail {

// d = |x - y|; returns all-ones when x < y
template <int N>
GM_HD uint32_t abs_diff(uint32_t* d, const uint32_t* x, const uint32_t* y) {
  d[0] = sub_cc(x[0], y[0]);
#pragma unroll
  for (int j = 1; j < N; j++) d[j] = subc_cc(x[j], y[j]);
  const uint32_t m = subc(0, 0);
  // two's complement negation when the difference wrapped: (d ^ m) + (m & 1)
  (void)add_cc(m, m);  // CF = m & 1
#pragma unroll
  for (int j = 0; j < N - 1; j++) d[j] = addc_cc(d[j] ^ m, 0);
  d[N - 1] = addc(d[N - 1] ^ m, 0);
  return m;
}

// t[0..2N) = a * b for any N: product scanning over a three-word column accumulator (the odd-sized leaves of the
// Karatsuba recursion)
template <int N>
GM_HD void mul_full_ps(uint32_t* t, const uint32_t* a, const uint32_t* b) {
  uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
  for (int k = 0; k < 2 * N - 1; k++) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int j = k - i;
      if (j >= 0 && j < N) mad_acc3(c0, c1, c2, a[i], b[j]);
    }
    t[k] = c0;
    c0 = c1; c1 = c2; c2 = 0;
  }
  t[2 * N - 1] = c0;
}

// T[0..2N) = a * b by LEVELS levels of (subtractive) Karatsuba, N = 2H
template <int N, int LEVELS>
GM_HD void mul_full_karatsuba(uint32_t* T, const uint32_t* a, const uint32_t* b) {
  constexpr int H = N / 2;
  static_assert(N % 2 == 0, "Karatsuba splits an even number of limbs");
  if constexpr (LEVELS == 0) {
    if constexpr (N % 2 == 0) mul_full<N>(T, a, b); else mul_full_ps<N>(T, a, b);
  } else {
    uint32_t z0[N], z2[N], zm[N], da[H], db[H];
    if constexpr (LEVELS > 1 && H % 2 == 0) {
      mul_full_karatsuba<H, LEVELS - 1>(z0, a, b);
      mul_full_karatsuba<H, LEVELS - 1>(z2, a + H, b + H);
    } else if constexpr (H % 2 == 0) {
      mul_full<H>(z0, a, b);
      mul_full<H>(z2, a + H, b + H);
    } else {
      mul_full_ps<H>(z0, a, b);
      mul_full_ps<H>(z2, a + H, b + H);
    }
    const uint32_t sa = abs_diff<H>(da, a, a + H);   // a_lo - a_hi
    const uint32_t sb = abs_diff<H>(db, b + H, b);   // b_hi - b_lo
    if constexpr (LEVELS > 1 && H % 2 == 0) mul_full_karatsuba<H, LEVELS - 1>(zm, da, db);
    else if constexpr (H % 2 == 0) mul_full<H>(zm, da, db);
    else mul_full_ps<H>(zm, da, db);
    const uint32_t neg = sa ^ sb;
    // mid = z0 + z2 + (a_lo - a_hi)(b_hi - b_lo) = a_lo b_hi + a_hi b_lo   (N + 1 limbs)
    uint32_t mid[N + 1];
    mid[0] = add_cc(z0[0], z2[0]);
#pragma unroll
    for (int k = 1; k < N; k++) mid[k] = addc_cc(z0[k], z2[k]);
    mid[N] = addc(0, 0);
    (void)add_cc(neg, neg);  // CF = 1 when the cross term is negative: mid += ~zm + 1 (sign-extended)
#pragma unroll
    for (int k = 0; k < N; k++) mid[k] = addc_cc(mid[k], zm[k] ^ neg);
    mid[N] = addc(mid[N], neg);
    // T = z0 + mid 2^(32H) + z2 2^(32N)
#pragma unroll
    for (int k = 0; k < H; k++) T[k] = z0[k];
    T[H] = add_cc(z0[H], mid[0]);
#pragma unroll
    for (int k = 1; k < H; k++) T[H + k] = addc_cc(z0[H + k], mid[k]);
#pragma unroll
    for (int k = 0; k < H; k++) T[N + k] = addc_cc(z2[k], mid[H + k]);
    if constexpr (H + 1 < N) {
      T[N + H] = addc_cc(z2[H], mid[N]);
#pragma unroll
      for (int k = H + 1; k < N - 1; k++) T[N + k] = addc_cc(z2[k], 0);
      T[2 * N - 1] = addc(z2[N - 1], 0);
    } else {
      T[N + H] = addc(z2[H], mid[N]);
    }
  }
}

// One row of the Montgomery reduction of a value that receives no further products.  On entry column 0 of the window
// is E[0] + S (S = stray limb of the previous odd accumulator, S = 0 on the first row), the odd accumulator is O.
// m*p is added so that column 0 becomes a multiple of 2^32; afterwards E[0] = -S mod 2^32, i.e. column 0 carries
// exactly (S != 0) into column 1 - injected as the carry-in of the odd chain, no carry propagation needed.
template <class P>
GM_HD void redc_row(uint32_t* E, uint32_t* O, uint32_t S) {
  constexpr int N = P::N;
  const uint32_t m = (E[0] + S) * P::inv();
  (void)add_cc(S, 0xffffffffu);  // CF = (S != 0)
  madc_wide_cc(O[0], O[1], m, P::mod(1), O[0], O[1]);
#pragma unroll
  for (int j = 3; j < N; j += 2) madc_wide_cc(O[j - 1], O[j], m, P::mod(j), O[j - 1], O[j]);
  mad_wide_cc(E[0], E[1], m, P::mod(0), E[0], E[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) madc_wide_cc(E[j], E[j + 1], m, P::mod(j), E[j], E[j + 1]);
  O[N - 1] = addc(O[N - 1], 0);
}

// u = lo * 2^(-32N) mod p, u <= p, for an arbitrary N-limb lo.  The window slides by renaming: after a row the old odd
// accumulator is the new even one, and the new odd accumulator is the old even one two limbs up (its two top limbs
// start from zero).
template <class P>
GM_HD void redc_half(uint32_t* u, const uint32_t* lo) {
  constexpr int N = P::N;
  uint32_t x[N], y[N];
#pragma unroll
  for (int k = 0; k < N; k++) { x[k] = lo[k]; y[k] = 0; }
  uint32_t S = 0;
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    redc_row<P>(x, y, S);   // even = x, odd = y
    S = x[1];
#pragma unroll
    for (int k = 0; k < N - 2; k++) x[k] = x[k + 2];
    x[N - 2] = 0; x[N - 1] = 0;
    redc_row<P>(y, x, S);   // even = y, odd = x
    S = y[1];
#pragma unroll
    for (int k = 0; k < N - 2; k++) y[k] = y[k + 2];
    y[N - 2] = 0; y[N - 1] = 0;
  }
  // N even: even = x, odd = y.  Column 0 = x[0] + S, column k = x[k] + y[k-1]; the value is <= p, so y[N-1] = 0.
  u[0] = add_cc(x[0], S);
#pragma unroll
  for (int k = 1; k < N - 1; k++) u[k] = addc_cc(x[k], y[k - 1]);
  u[N - 1] = addc(x[N - 1], y[N - 2]);
}

}  // namespace detail

template <class P, int LEVELS = 1>
GM_HD void mont_mul_karatsuba(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = P::N;
  uint32_t T[2 * N], u[N], t[N];
  detail::mul_full_karatsuba<N, LEVELS>(T, a, b);
  detail::redc_half<P>(u, T);
  // a, b < p: T_hi + u < p^2 / 2^(32N) + p + 1 < 2p
  t[0] = add_cc(T[N], u[0]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) t[k] = addc_cc(T[N + k], u[k]);
  t[N - 1] = addc(T[2 * N - 1], u[N - 1]);
  detail::cond_sub_p<P>(r, t);
}


}  // namespace gm
