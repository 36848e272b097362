// Fq Montgomery multiplication on the FP64 pipe (B200: DFMA at 64 lanes/clk/SM, issued concurrently with
// the IMAD.WIDE integer multiplier - profiles/r01_imad_microbench.txt).
//
// Same function as mont_mul<FqParams> (fp.cuh): inputs and output are 12 x u32 Montgomery limbs
// (R = 2^384), fully reduced, bit-identical results.  Inside, numbers are 16 limbs of 24 bits held in
// doubles (16 x 24 = 384, so the radix-2^24 Montgomery reduction removes exactly R):
//   * every partial product a_j * b_i (< 2^48) and q * p~_j (|.| <= 2^46) is ONE fma whose result is an
//     exactly representable integer: a column receives at most 16 + 16 terms, |column| < 2^53;
//   * the modulus is stored in BALANCED limbs p~_j in [-2^23, 2^23) and the quotient digit q is balanced
//     too, which is what keeps the column sums below 2^53;
//   * rounding to an integer uses the 1.5 * 2^52 magic constant, so no int<->double conversion
//     instructions are needed inside the loop.
// The identical source runs on the host (std::fma), where tests/test_host_field.py compares it with
// Python integers.
#pragma once
#include "../gemini_b200/csrc/fp.cuh"
#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cstring>
#endif

namespace gm {
namespace f64 {

GM_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
GM_HD double add_(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;  // keep (x + M) - M from being folded
  return r;
#endif
}
// exact for 0 <= x < 2^32
GM_HD double u2d(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(__hiloint2double(0x43300000, (int)x), -4503599627370496.0);
#else
  return (double)x;
#endif
}
// low 32 bits of the integer value of x (|x| < 2^51), two's complement
GM_HD int32_t d2i(double x) {
  const double t = add_(x, 6755399441055744.0);
#if defined(__CUDA_ARCH__)
  return __double2loint(t);
#else
  uint64_t bits;
  std::memcpy(&bits, &t, 8);
  return (int32_t)(uint32_t)bits;
#endif
}
// nearest integer of x, |x| < 2^51
GM_HD double rnd(double x) { return add_(add_(x, 6755399441055744.0), -6755399441055744.0); }

constexpr double TWO24 = 16777216.0;
constexpr double INV24 = 1.0 / 16777216.0;
constexpr double PINV_NEG = 16580605.0;  // -q^{-1} mod 2^24

// balanced radix-2^24 limbs of q
GM_HD constexpr double ptil(int j) {
  constexpr double t[16] = {-21845.0, 0.0, -17921.0, -5155840.0, -5505025.0, -646113.0, -6228303.0, 6762707.0,
                            -8056129.0, 4949236.0, -2661257.0, 4410285.0, 1812406.0, -1664437.0, -1427072.0, 1704210.0};
  return t[j];
}

// 12 x u32 -> 16 unsigned 24-bit limbs
GM_HD void split24(const uint32_t* w, uint32_t* l) {
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const uint32_t w0 = w[3 * g], w1 = w[3 * g + 1], w2 = w[3 * g + 2];
    l[4 * g + 0] = w0 & 0xFFFFFFu;
    l[4 * g + 1] = (w0 >> 24) | ((w1 & 0xFFFFu) << 8);
    l[4 * g + 2] = (w1 >> 16) | ((w2 & 0xFFu) << 16);
    l[4 * g + 3] = w2 >> 8;
  }
}

// r = a * b * 2^-384 mod q
GM_HD void fq_mul_f64(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t al[16], bl[16];
  split24(a, al);
  split24(b, bl);
  double ad[16], acc[17];
#pragma unroll
  for (int j = 0; j < 16; j++) { ad[j] = u2d(al[j]); acc[j] = 0.0; }
  acc[16] = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const double bi = u2d(bl[i]);
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = fma_(ad[j], bi, acc[j]);
    // quotient digit: q = bal(-acc[0] * q^{-1} mod 2^24)
    const double h = rnd(acc[0] * INV24);
    const double lo = fma_(-h, TWO24, acc[0]);          // acc[0] mod 2^24, balanced
    const double prod = lo * PINV_NEG;                   // exact, < 2^48
    const double h2 = rnd(prod * INV24);
    const double qd = fma_(-h2, TWO24, prod);            // balanced quotient digit
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = fma_(qd, ptil(j), acc[j]);
    // acc[0] is now a multiple of 2^24: carry into the next column and slide the window
    acc[1] = fma_(acc[0], INV24, acc[1]);
#pragma unroll
    for (int j = 0; j < 16; j++) acc[j] = acc[j + 1];
    acc[16] = 0.0;
  }
  // carry-normalise to balanced 24-bit limbs, then to integers
  int32_t li[16];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const double c = rnd(acc[j] * INV24);
    const double lo = fma_(-c, TWO24, acc[j]);
    if (j < 15) acc[j + 1] = add_(acc[j + 1], c);
    li[j] = d2i(lo);
  }
  // signed limbs -> unsigned with borrow; a final borrow means the value is negative
  uint32_t ul[16];
  int32_t br = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const int32_t v = li[j] + br;
    ul[j] = (uint32_t)v & 0xFFFFFFu;
    br = v >> 24;  // arithmetic: 0 or -1 (v >= -2^23 - 1)
  }
  uint32_t t[12];
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const uint32_t l0 = ul[4 * g], l1 = ul[4 * g + 1], l2 = ul[4 * g + 2], l3 = ul[4 * g + 3];
    t[3 * g + 0] = l0 | (l1 << 24);
    t[3 * g + 1] = (l1 >> 8) | (l2 << 16);
    t[3 * g + 2] = (l2 >> 16) | (l3 << 8);
  }
  // value in (-q, 2q): add q when negative (two's complement wrap-around is exact), then reduce once
  const uint32_t neg = (uint32_t)br;  // 0 or 0xffffffff
  uint32_t u[12];
  u[0] = add_cc(t[0], FqParams::mod(0) & neg);
#pragma unroll
  for (int j = 1; j < 11; j++) u[j] = addc_cc(t[j], FqParams::mod(j) & neg);
  u[11] = addc(t[11], FqParams::mod(11) & neg);
  detail::cond_sub_p<FqParams>(r, u);
}

}  // namespace f64
}  // namespace gm
