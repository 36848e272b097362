// CPU-side harness that compiles the *device* limb schedules (fp.cuh / g1.cuh)
// with the host emulation of the PTX carry flag, so tests/test_host_field.py can
// compare them with Python big integers without a GPU.  Test scaffolding only.
#include "../../gemini_b200/csrc/fp.cuh"
#include "../../gemini_b200/csrc/g1.cuh"
#include "../../gemini_b200/csrc/g1_affine.cuh"
#include "../../tools/fq_f64.cuh"
#include "../../tools/fq_karatsuba.cuh"
#include "../../tools/fq_f64v2.cuh"
#include "../../gemini_b200/csrc/fp_inv_fast.cuh"
#include <string.h>
using namespace gm;

template <class F> static void bin(void (*op)(F&, const F&, const F&), const uint32_t* a, const uint32_t* b, uint32_t* r, int n) {
  for (int i = 0; i < n; i++) {
    F x, y, z;
    memcpy(x.v, a + i * F::N, 4 * F::N); memcpy(y.v, b + i * F::N, 4 * F::N);
    op(z, x, y);
    memcpy(r + i * F::N, z.v, 4 * F::N);
  }
}
template <class F> static void f_mul(F& z, const F& x, const F& y) { z = x * y; }
template <class F> static void f_add(F& z, const F& x, const F& y) { z = x + y; }
template <class F> static void f_sub(F& z, const F& x, const F& y) { z = x - y; }
template <class F> static void f_inv(F& z, const F& x, const F&) { z = fp_inv(x); }
template <class F> static void f_inv_fast(F& z, const F& x, const F&) { z = fp_inv_divsteps(x); }
template <class F> static void f_redc(F& z, const F& x, const F&) { z = x.from_mont(); }
template <class F> static void f_tom(F& z, const F& x, const F&) { z = x.to_mont(); }
template <class F> static void f_sqr(F& z, const F& x, const F&) { z = x.sqr(); }
static void f_mul_f64(Fq& z, const Fq& x, const Fq& y) { f64::fq_mul_f64(z.v, x.v, y.v); }
static void f_mul_f64v2(Fq& z, const Fq& x, const Fq& y) { f64v2::fq_mul(z.v, x.v, y.v); }
template <class F> static void f_mul_k(F& z, const F& x, const F& y) { mont_mul_karatsuba<typename F::Params>(z.v, x.v, y.v); }
template <class F> static void f_mul_k2(F& z, const F& x, const F& y) { mont_mul_karatsuba<typename F::Params, 2>(z.v, x.v, y.v); }

extern "C" {
void hc_fq(int op, const uint32_t* a, const uint32_t* b, uint32_t* r, int n) {
  void (*ops[])(Fq&, const Fq&, const Fq&) = {f_mul<Fq>, f_add<Fq>, f_sub<Fq>, f_inv<Fq>, f_redc<Fq>, f_tom<Fq>, f_sqr<Fq>, f_mul_f64, f_inv_fast<Fq>, f_mul_k<Fq>, f_mul_k2<Fq>, f_mul_f64v2};
  bin<Fq>(ops[op], a, b, r, n);
}
void hc_fr(int op, const uint32_t* a, const uint32_t* b, uint32_t* r, int n) {
  void (*ops[])(Fr&, const Fr&, const Fr&) = {f_mul<Fr>, f_add<Fr>, f_sub<Fr>, f_inv<Fr>, f_redc<Fr>, f_tom<Fr>, f_sqr<Fr>, f_sqr<Fr>, f_inv_fast<Fr>, f_mul_k<Fr>, f_mul_k2<Fr>};
  bin<Fr>(ops[op], a, b, r, n);
}
// plain 2N-limb products of ARBITRARY N-limb operands (the building blocks of mont_mul_karatsuba); which: 0 = schoolbook
// rows (mul_full), 1 = one Karatsuba level; nl = 4, 6, 8 or 12 limbs
void hc_mul_full(int which, int nl, const uint32_t* a, const uint32_t* b, uint32_t* t, int n) {
  for (int i = 0; i < n; i++) {
    const uint32_t* x = a + i * nl; const uint32_t* y = b + i * nl; uint32_t* z = t + 2 * i * nl;
    if (which == 0) {
      if (nl == 4) detail::mul_full<4>(z, x, y); else if (nl == 6) detail::mul_full<6>(z, x, y);
      else if (nl == 8) detail::mul_full<8>(z, x, y); else detail::mul_full<12>(z, x, y);
    } else if (which == 1) {
      if (nl == 8) detail::mul_full_karatsuba<8, 1>(z, x, y); else detail::mul_full_karatsuba<12, 1>(z, x, y);
    } else if (which == 2) {
      if (nl == 8) detail::mul_full_karatsuba<8, 2>(z, x, y); else detail::mul_full_karatsuba<12, 2>(z, x, y);
    } else {
      if (nl == 3) detail::mul_full_ps<3>(z, x, y); else if (nl == 6) detail::mul_full_karatsuba<6, 1>(z, x, y);
      else detail::mul_full_ps<5>(z, x, y);
    }
  }
}
// u = lo * 2^(-32 N) mod p (<= p) for arbitrary N-limb lo
void hc_redc_half(int is_fr, const uint32_t* lo, uint32_t* u, int n) {
  for (int i = 0; i < n; i++) {
    if (is_fr) detail::redc_half<FrParams>(u + 8 * i, lo + 8 * i); else detail::redc_half<FqParams>(u + 12 * i, lo + 12 * i);
  }
}
// XYZZ accumulator (48 u32) (+)= affine point (24 u32, (0,0) = identity), optionally negated
void hc_xyzz_madd(uint32_t* acc, const uint32_t* aff, int neg, int n) {
  for (int i = 0; i < n; i++) {
    XYZZ A; memcpy(&A, acc + 48 * i, 192);
    Affine P; memcpy(&P, aff + 24 * i, 96);
    if (neg) P.y = P.y.neg();
    xyzz_madd(A, P);
    memcpy(acc + 48 * i, &A, 192);
  }
}
void hc_xyzz_add(uint32_t* acc, const uint32_t* other, int n) {
  for (int i = 0; i < n; i++) {
    XYZZ A, B; memcpy(&A, acc + 48 * i, 192); memcpy(&B, other + 48 * i, 192);
    xyzz_add(A, B);
    memcpy(acc + 48 * i, &A, 192);
  }
}
void hc_xyzz_dbl(uint32_t* acc, int n) {
  for (int i = 0; i < n; i++) {
    XYZZ A; memcpy(&A, acc + 48 * i, 192);
    xyzz_dbl(A);
    memcpy(acc + 48 * i, &A, 192);
  }
}
// XYZZ -> normalised Jacobian (x, y, 1) in Montgomery form, (0,1,0)-style identity => all zero z
void hc_xyzz_to_jacobian(const uint32_t* acc, uint32_t* out, int n) {
  for (int i = 0; i < n; i++) {
    XYZZ A; memcpy(&A, acc + 48 * i, 192);
    Jacobian J = xyzz_to_jacobian_normalized(A);
    memcpy(out + 36 * i, &J, 144);
  }
}
// One batch of affine pair additions with ONE shared inversion, walked exactly like a thread of
// k_aff_prepare / k_aff_finish (msm.cu): exclusive prefix products forward, back-substitution backward.
// p1, p2: n x 24 u32; has2: n flags; out: n x 24 u32; kinds: n
void hc_aff_batch(const uint32_t* p1, const uint32_t* p2, const int* has2, int n, uint32_t* out, uint32_t* kinds) {
  Fq* prefix = new Fq[n];
  Fq run = Fq::one();
  for (int i = 0; i < n; i++) {
    Affine a, b; memcpy(&a, p1 + 24 * i, 96); memcpy(&b, p2 + 24 * i, 96);
    Fq den;
    kinds[i] = aff_pair_kind(a, b, has2[i] != 0, den);
    if (aff_kind_needs_inverse(kinds[i])) { prefix[i] = run; run = run * den; }
  }
  Fq inv = fp_inv(run);
  for (int i = n - 1; i >= 0; i--) {
    Affine a, b; memcpy(&a, p1 + 24 * i, 96); memcpy(&b, p2 + 24 * i, 96);
    Fq inv_den = Fq::one();
    if (aff_kind_needs_inverse(kinds[i])) {
      Fq den;
      aff_pair_kind(a, b, has2[i] != 0, den);
      inv_den = inv * prefix[i];
      inv = inv * den;
    }
    Affine r = aff_pair_finish(kinds[i], a, b, inv_den);
    memcpy(out + 24 * i, &r, 96);
  }
  delete[] prefix;
}
// lazy-reduction accumulator: out = sum_i a_i * b_i (Montgomery form), n pairs of 8 u32
void hc_fr_sum_of_products(const uint32_t* a, const uint32_t* b, int n, uint32_t* out) {
  FrAcc acc = FrAcc::zero();
  for (int i = 0; i < n; i++) {
    Fr x, y; memcpy(x.v, a + 8 * i, 32); memcpy(y.v, b + 8 * i, 32);
    // the schedule of pair_contrib (fr.cu): UNREDUCED_RUN = 2 products, then one normalisation
    acc.mul_add_unreduced(x, y);
    if (i & 1) acc.normalize();
  }
  acc.normalize();
  Fr r = acc.reduce();
  memcpy(out, r.v, 32);
}
}
