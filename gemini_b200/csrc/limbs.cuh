// 32-bit limb primitives with an explicit carry chain.
//
// On the device every primitive is one PTX instruction of the add.cc / madc
// family (ptxas fuses adjacent mad.lo.cc + madc.hi.cc on the same operands into
// one IMAD.WIDE.U32.X on sm_100a).  On the host the same primitives are emulated
// with a thread-local carry flag so that the *identical* limb schedules in
// fp.cuh can be exercised by the CPU-only test-suite (tests/test_host_field.py)
// before any GPU time is spent.  The host emulation is test scaffolding for the
// limb schedule, it is never used to compute a product result.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GM_HD __host__ __device__ __forceinline__
#define GM_D __device__ __forceinline__
#else
#define GM_HD inline
#define GM_D inline
#endif

namespace gm {

#if defined(__CUDA_ARCH__)

GM_D uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
GM_D uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
GM_D uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
GM_D uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
GM_D uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
GM_D uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
GM_D uint32_t mul_lo(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
GM_D uint32_t mul_hi(uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
GM_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
GM_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
GM_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
GM_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}

// 64-bit partial product + 64-bit addend as ONE asm statement, so that ptxas sees
// the same virtual registers in the lo and hi halves and emits IMAD.WIDE.U32(.X).
GM_D void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) = a*b + (clo,chi); carry out in CF
GM_D void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
               : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// (lo,hi) = a*b + (clo,chi) + CF; carry out in CF
GM_D void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
               : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// (lo,hi) = a*b + CF; no carry out (top of a chain)
GM_D void madc_wide_top(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// three-word column accumulator (c0, c1, c2) += a*b   (product scanning; IMAD.WIDE + IADD3.X in SASS)
GM_D void mad_acc3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
               : "+r"(c0), "+r"(c1), "+r"(c2) : "r"(a), "r"(b));
}

#else  // host emulation of the PTX carry flag ------------------------------

namespace detail { inline uint32_t& cf() { static thread_local uint32_t f = 0; return f; } }

inline uint32_t add_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b; detail::cf() = (uint32_t)(t >> 32); return (uint32_t)t;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b + detail::cf(); detail::cf() = (uint32_t)(t >> 32); return (uint32_t)t;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + detail::cf(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b; detail::cf() = (uint32_t)(t >> 63); return (uint32_t)t;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b - detail::cf(); detail::cf() = (uint32_t)(t >> 63); return (uint32_t)t;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - detail::cf(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (uint64_t)mul_lo(a, b) + c; detail::cf() = (uint32_t)(t >> 32); return (uint32_t)t;
}
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (uint64_t)mul_lo(a, b) + c + detail::cf(); detail::cf() = (uint32_t)(t >> 32); return (uint32_t)t;
}
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (uint64_t)mul_hi(a, b) + c + detail::cf(); detail::cf() = (uint32_t)(t >> 32); return (uint32_t)t;
}
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + detail::cf(); }

inline void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { lo = mul_lo(a, b); hi = mul_hi(a, b); }
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  uint32_t l = mad_lo_cc(a, b, clo); uint32_t h = madc_hi_cc(a, b, chi); lo = l; hi = h;
}
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  uint32_t l = madc_lo_cc(a, b, clo); uint32_t h = madc_hi_cc(a, b, chi); lo = l; hi = h;
}
inline void madc_wide_top(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint32_t l = madc_lo_cc(a, b, 0); uint32_t h = madc_hi(a, b, 0); lo = l; hi = h;
}
inline void mad_acc3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t a, uint32_t b) {
  const uint64_t p = (uint64_t)a * b;
  const uint64_t t0 = (uint64_t)c0 + (uint32_t)p;
  const uint64_t t1 = (uint64_t)c1 + (uint32_t)(p >> 32) + (t0 >> 32);
  c0 = (uint32_t)t0; c1 = (uint32_t)t1; c2 += (uint32_t)(t1 >> 32);
}

#endif

}  // namespace gm
