// EXPERIMENT: Fq Montgomery product on the FP64 pipe with 48-bit limbs (8 x 48 = 384 = the Montgomery radix of the
// 12 x u32 representation, so inputs and output are the library's own Fq limbs, bit for bit).
//
// B200 issues DFMA (64 lanes/clk/SM) CONCURRENTLY with IMAD.WIDE (32 lanes/clk/SM, profiles/r01_imad_microbench.txt).
// The 24-bit-limb formulation of round 1 (tools/fq_f64.cuh) needed 512 + ~390 FP64 operations per product and lost.  Here a
// 48 x 48-bit limb product a*b < 2^96 is split EXACTLY into its high and low 48 bits by two fused multiply-adds:
//     hn = fma_rz(a, b, H)        H = 2^100 + S, S a multiple of 2^48: the sum stays in [2^100, 2^101), whose ulp is 2^48,
//                                 so rounding toward zero leaves H + floor(a b / 2^48) 2^48 - the high part ACCUMULATES for free
//     t  = hn - H                 = floor(a b / 2^48) 2^48 (exact)
//     lo = fma(a, b, -t)          = a b mod 2^48 (exact)
//     L += lo
// 4 FP64 operations per limb product, no integer instruction, no conversion; the column sums (at most 30 terms below 2^48
// plus a carry) stay below 2^53 and are exact.  One CIOS row = 8 products with b_i, the quotient digit
// q = (column 0) * (-p^-1) mod 2^48 (3 operations to split column 0, 3 for the low product), 8 products with p: 74
// operations, 592 per product + conversions.
//
// The identical source runs on the host (std::fma under FE_TOWARDZERO) for tests/test_host_field.py.
#pragma once
#include "../gemini_b200/csrc/fp.cuh"
#if !defined(__CUDA_ARCH__)
#include <cfenv>
#include <cmath>
#include <cstring>
#endif

namespace gm {
namespace f64v2 {

GM_HD double fma_rn(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
GM_HD double fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rz(a, b, c);
#else
  std::fesetround(FE_TOWARDZERO);
  volatile double r = std::fma(a, b, c);
  std::fesetround(FE_TONEAREST);
  return r;
#endif
}
GM_HD double add_rz(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rz(a, b);
#else
  std::fesetround(FE_TOWARDZERO);
  volatile double x = a, y = b;
  volatile double r = x + y;
  std::fesetround(FE_TONEAREST);
  return r;
#endif
}
GM_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
GM_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;
  return r;
#endif
}

constexpr double C1 = 1267650600228229401496703205376.0;   // 2^100
constexpr double M48 = 1.0 / 281474976710656.0;             // 2^-48
constexpr double T52 = 4503599627370496.0;                  // 2^52
constexpr double PINV = 281462091612157.0;                  // -q^-1 mod 2^48

// 48-bit limbs of q
GM_HD constexpr double plimb(int j) {
  constexpr double t[8] = {281474976688811.0, 194974335351294.0, 270634993844222.0, 113459389855408.0,
                           83034393350847.0,  73992301405303.0,  253550359455670.0, 28591897852287.0};
  return t[j];
}

GM_HD double bits_to_double(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((int)hi, (int)lo);
#else
  const uint64_t b = ((uint64_t)hi << 32) | lo;
  double d;
  std::memcpy(&d, &b, 8);
  return d;
#endif
}
GM_HD void double_to_bits(double d, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  hi = (uint32_t)__double2hiint(d);
  lo = (uint32_t)__double2loint(d);
#else
  uint64_t b;
  std::memcpy(&b, &d, 8);
  hi = (uint32_t)(b >> 32);
  lo = (uint32_t)b;
#endif
}

// 12 x u32 -> 8 limbs of 48 bits as doubles (exact integers)
GM_HD void to_limbs48(double* d, const uint32_t* w) {
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const uint32_t w0 = w[3 * m], w1 = w[3 * m + 1], w2 = w[3 * m + 2];
    d[2 * m] = add_rn(bits_to_double(0x43300000u | (w1 & 0xFFFFu), w0), -T52);
    d[2 * m + 1] = add_rn(bits_to_double(0x43300000u | (w2 >> 16), (w1 >> 16) | (w2 << 16)), -T52);
  }
}

// acc (H[j+1], L[j]) += x * y for one limb product
#define GM_F64_MAC(x, y, Hn, Ln)                    \
  do {                                              \
    const double _hn = fma_rz((x), (y), (Hn));      \
    const double _t = add_rn(_hn, -(Hn));           \
    (Hn) = _hn;                                     \
    (Ln) = add_rn((Ln), fma_rn((x), (y), -_t));     \
  } while (0)

// r = a * b * 2^-384 mod q, fully reduced; a, b < q
GM_HD void fq_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  double ad[8], bd[8];
  to_limbs48(ad, a);
  to_limbs48(bd, b);
  double L[9], H[9];
#pragma unroll
  for (int k = 0; k < 9; k++) { L[k] = 0.0; H[k] = C1; }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const double bi = bd[i];
#pragma unroll
    for (int j = 0; j < 8; j++) GM_F64_MAC(ad[j], bi, H[j + 1], L[j]);
    // column 0 of the window, its low 48 bits, the quotient digit
    double c0 = fma_rn(add_rn(H[0], -C1), M48, L[0]);
    const double u = add_rn(add_rz(c0, C1), -C1);
    const double cl = add_rn(c0, -u);
    const double hq = fma_rz(cl, PINV, C1);
    const double q = fma_rn(cl, PINV, -add_rn(hq, -C1));
    {
      const double hn = fma_rz(q, plimb(0), H[1]);
      const double t = add_rn(hn, -H[1]);
      H[1] = hn;
      c0 = add_rn(c0, fma_rn(q, plimb(0), -t));   // now a multiple of 2^48
    }
#pragma unroll
    for (int j = 1; j < 8; j++) GM_F64_MAC(q, plimb(j), H[j + 1], L[j]);
    L[1] = fma_rn(c0, M48, L[1]);                 // carry out of column 0
#pragma unroll
    for (int k = 0; k < 8; k++) { L[k] = L[k + 1]; H[k] = H[k + 1]; }
    L[8] = 0.0;
    H[8] = C1;
  }
  // carry-normalise the eight result columns and repack into 12 x u32
  uint32_t lo32[8], hi16[8];
  double carry = 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const double v = add_rn(fma_rn(add_rn(H[k], -C1), M48, L[k]), carry);
    const double u = add_rn(add_rz(v, C1), -C1);
    const double limb = add_rn(v, -u);
    carry = mul_rn(u, M48);
    uint32_t hi;
    double_to_bits(add_rn(limb, T52), hi, lo32[k]);
    hi16[k] = hi & 0xFFFFu;
  }
  uint32_t t[12];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    t[3 * m] = lo32[2 * m];
    t[3 * m + 1] = hi16[2 * m] | (lo32[2 * m + 1] << 16);
    t[3 * m + 2] = (lo32[2 * m + 1] >> 16) | (hi16[2 * m + 1] << 16);
  }
  detail::cond_sub_p<FqParams>(r, t);
}

}  // namespace f64v2
}  // namespace gm
