// Modular inversion by batched division steps (Bernstein-Yang "safegcd", in the 30-bit-limb form popularised by
// Pornin and by libsecp256k1's modinv32).  Validated on the host against Python integers
// (tests/test_host_field.py::test_fast_inverse_divsteps) and on the device (tests/test_gpu_arith.py, selftest op 7);
// it is the inversion of every latency-critical kernel: k_aff_invert (one per warp of every affine level) and
// k_normalize / xyzz_to_jacobian_normalized (the tail of every MSM).
//
// Why: the Kaliski almost-inverse of fp.cuh walks ~540 dependent big-number iterations (0.11-0.16 ms on B200, the
// fixed cost of every affine level of the MSM and of k_normalize).  Here 30 division steps are decided on the low
// 32 bits of (f, g) alone and applied to the big numbers as ONE 2x2 integer matrix: ~25-37 big-number updates.
//
//   divstep(delta, f, g) = (1 - delta, g, (g - f) / 2)            if delta > 0 and g odd
//                          (1 + delta, f, (g + (g mod 2) f) / 2)  otherwise
// f = p, g = a; when g reaches 0, f = +-1 and d * a = f (mod p), where (d, e) started as (0, 1) and went through
// the same matrices (mod p).
#pragma once
#include "fp.cuh"

namespace gm {
namespace fastinv {

static constexpr int32_t M30 = (int32_t)0x3FFFFFFF;

// Signed integers in L limbs of 30 bits (the top limb carries the sign).
template <int L>
struct S30 {
  int32_t v[L];
};

template <int N> struct Limbs30 { static constexpr int L = (32 * N + 2 + 29) / 30; };  // room for (-2p, p)

// N x u32 (non-negative) -> 30-bit limbs
template <int N, int L>
GM_HD void to_s30(S30<L>& r, const uint32_t* a) {
#pragma unroll
  for (int i = 0; i < L; i++) {
    const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
    uint64_t two = 0;
    if (w < N) two = a[w];
    if (w + 1 < N) two |= (uint64_t)a[w + 1] << 32;
    r.v[i] = (int32_t)((two >> sh) & (uint32_t)M30);
  }
}
// 30-bit limbs (value in [0, 2^(32N))) -> N x u32
template <int N, int L>
GM_HD void from_s30(uint32_t* a, const S30<L>& r) {
#pragma unroll
  for (int w = 0; w < N; w++) a[w] = 0;
#pragma unroll
  for (int i = 0; i < L; i++) {
    const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
    const uint64_t val = (uint64_t)(uint32_t)r.v[i] << sh;
    if (w < N) a[w] |= (uint32_t)val;
    if (w + 1 < N) a[w + 1] |= (uint32_t)(val >> 32);
  }
}

struct Trans {  // 2^30 * [f'; g'] = [[u, v], [q, r]] * [f; g]
  int32_t u, v, q, r;
};

// 30 division steps on the low bits of f (odd) and g; returns the new eta = -delta
GM_HD int32_t divsteps_30(int32_t eta, uint32_t f0, uint32_t g0, Trans& t) {
  int32_t u = 1, v = 0, q = 0, r = 1;
  uint32_t f = f0, g = g0;
#pragma unroll 1
  for (int i = 0; i < 30; i++) {
    if (g & 1u) {
      if (eta < 0) {            // delta > 0: (f, g) <- (g, g - f), rows swap
        const uint32_t tf = f; f = g; g = g - tf;
        const int32_t tu = u, tv = v;
        u = q; v = r; q = q - tu; r = r - tv;
        eta = -eta;             // 1 - delta = -(eta') ... eta' = -(1 - delta) = -eta - 1 + ... (the -1 is applied below)
      } else {
        g = g + f; q = q + u; r = r + v;
      }
    }
    // halve g; the f row doubles instead (everything is scaled by 2 per step)
    g >>= 1; u += u; v += v;
    eta -= 1;
  }
  t.u = u; t.v = v; t.q = q; t.r = r;
  return eta;
}

// (f, g) <- t * (f, g) / 2^30   (exact)
template <int L>
GM_HD void update_fg(S30<L>& f, S30<L>& g, const Trans& t) {
  const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
  int64_t cf = u * f.v[0] + v * g.v[0];
  int64_t cg = q * f.v[0] + r * g.v[0];
  cf >>= 30; cg >>= 30;
#pragma unroll
  for (int i = 1; i < L; i++) {
    cf += u * f.v[i] + v * g.v[i];
    cg += q * f.v[i] + r * g.v[i];
    f.v[i - 1] = (int32_t)cf & M30; cf >>= 30;
    g.v[i - 1] = (int32_t)cg & M30; cg >>= 30;
  }
  f.v[L - 1] = (int32_t)cf;
  g.v[L - 1] = (int32_t)cg;
}

// (d, e) <- t * (d, e) / 2^30 mod p, both kept in (-2p, p); pinv30 = p^{-1} mod 2^30
template <int L>
GM_HD void update_de(S30<L>& d, S30<L>& e, const Trans& t, const S30<L>& p, uint32_t pinv30) {
  const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
  const int32_t sd = d.v[L - 1] >> 31, se = e.v[L - 1] >> 31;     // all-ones when negative
  // start with enough multiples of p to keep the results above -2p
  int32_t md = (t.u & sd) + (t.v & se);
  int32_t me = (t.q & sd) + (t.r & se);
  int64_t cd = u * d.v[0] + v * e.v[0];
  int64_t ce = q * d.v[0] + r * e.v[0];
  // make the low 30 bits of (c + m p) vanish
  md -= (int32_t)((pinv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((pinv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd += (int64_t)p.v[0] * md;
  ce += (int64_t)p.v[0] * me;
  cd >>= 30; ce >>= 30;
#pragma unroll
  for (int i = 1; i < L; i++) {
    cd += u * d.v[i] + v * e.v[i] + (int64_t)p.v[i] * md;
    ce += q * d.v[i] + r * e.v[i] + (int64_t)p.v[i] * me;
    d.v[i - 1] = (int32_t)cd & M30; cd >>= 30;
    e.v[i - 1] = (int32_t)ce & M30; ce >>= 30;
  }
  d.v[L - 1] = (int32_t)cd;
  e.v[L - 1] = (int32_t)ce;
}

// r = sign * d mod p, brought to [0, p) : d in (-2p, p), sign = +-1 (f at the end)
template <int L>
GM_HD void normalize(S30<L>& d, int32_t neg, const S30<L>& p) {
  // add p while negative (at most twice), negate if asked, add p again if that made it negative
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    const int32_t s = d.v[L - 1] >> 31;
    int32_t c = 0;
#pragma unroll
    for (int i = 0; i < L; i++) { c += d.v[i] + (p.v[i] & s); d.v[i] = (i < L - 1) ? (c & M30) : c; if (i < L - 1) c >>= 30; }
  }
  if (neg) {
    int32_t c = 0;
#pragma unroll
    for (int i = 0; i < L; i++) { c += -d.v[i]; d.v[i] = (i < L - 1) ? (c & M30) : c; if (i < L - 1) c >>= 30; }
    const int32_t s = d.v[L - 1] >> 31;
    c = 0;
#pragma unroll
    for (int i = 0; i < L; i++) { c += d.v[i] + (p.v[i] & s); d.v[i] = (i < L - 1) ? (c & M30) : c; if (i < L - 1) c >>= 30; }
  }
}

template <int L>
GM_HD bool is_zero(const S30<L>& g) {
  int32_t acc = 0;
#pragma unroll
  for (int i = 0; i < L; i++) acc |= g.v[i];
  return acc == 0;
}

}  // namespace fastinv

// a != 0 in Montgomery form -> a^{-1} in Montgomery form (same contract as fp_inv)
template <class P>
GM_HD Fp<P> fp_inv_divsteps(const Fp<P>& a) {
  using namespace fastinv;
  constexpr int N = P::N;
  constexpr int L = Limbs30<N>::L;
  uint32_t pm[N];
#pragma unroll
  for (int j = 0; j < N; j++) pm[j] = P::mod(j);
  S30<L> p, f, g, d, e;
  to_s30<N, L>(p, pm);
  f = p;
  to_s30<N, L>(g, a.v);
#pragma unroll
  for (int i = 0; i < L; i++) { d.v[i] = 0; e.v[i] = 0; }
  e.v[0] = 1;
  // p^{-1} mod 2^30 from -p^{-1} mod 2^32 (the Montgomery constant)
  const uint32_t pinv30 = (0u - P::INV) & (uint32_t)M30;
  int32_t eta = -1;
#pragma unroll 1
  for (int it = 0; it < (49 * 32 * N + 57) / 17 / 30 + 2; it++) {     // divstep bound of Bernstein-Yang, in batches of 30
    Trans t;
    eta = divsteps_30(eta, (uint32_t)f.v[0] | ((uint32_t)f.v[1] << 30), (uint32_t)g.v[0] | ((uint32_t)g.v[1] << 30), t);
    update_de<L>(d, e, t, p, pinv30);
    update_fg<L>(f, g, t);
    if (is_zero<L>(g)) break;
  }
  // f = +-1:  d * (aR) = f (mod p)
  normalize<L>(d, f.v[L - 1] >> 31 ? 1 : 0, p);
  Fp<P> x;
  from_s30<N, L>(x.v, d);
  // x = (aR)^{-1} = a^{-1} R^{-1}; the Montgomery form of a^{-1} is a^{-1} R = x R^2: two products by R^2
  Fp<P> r2;
#pragma unroll
  for (int j = 0; j < N; j++) r2.v[j] = P::r2(j);
  return (x * r2) * r2;
}

// The inversion the latency-critical kernels call (k_aff_invert, k_normalize).  -DGM_KALISKI_INV switches back to
// the binary almost-inverse of fp.cuh (A/B measurements).
template <class P>
GM_HD Fp<P> fp_inv_serial(const Fp<P>& a) {
#ifdef GM_KALISKI_INV
  return fp_inv(a);
#else
  return fp_inv_divsteps(a);
#endif
}

}  // namespace gm
