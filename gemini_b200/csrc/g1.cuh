// BLS12-381 G1 group law (y^2 = x^3 + 4 over Fq) in extended Jacobian (XYZZ)
// coordinates: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, identity <=> ZZ == 0.
//
// Replaces the `G::Group += &G` / `-= ` bucket updates and the running-sum adds
// of arkworks' Pippenger (spec: /root/reference/src/kzg/msm/variable_base.rs:
// 135-166).  Every routine is complete for the degenerate inputs the reference's
// default workloads produce: identity bases (kzg/time.rs:87 index_by), P + P
// (all-identical bases, examples/snark.rs:63) and P + (-P).
//
// Memory formats
//   Affine   : x | y, 96 B, Montgomery limbs; the pair (0, 0) - not on the curve -
//              encodes the identity (arkworks Affine::identity() is x=y=0,infinity=1)
//   Jacobian : X | Y | Z, 144 B = arkworks' Projective<Config>; identity is (1,1,0)
#pragma once
#include "fp.cuh"
#include "fp_inv_fast.cuh"

namespace gm {

struct Affine {
  Fq x, y;
  GM_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

struct Jacobian {
  Fq x, y, z;
};

struct XYZZ {
  Fq x, y, zz, zzz;
  GM_HD bool is_identity() const { return zz.is_zero(); }
  GM_HD static XYZZ identity() {
    XYZZ r;
    r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero();
    return r;
  }
};

// acc = 2 * (affine p), p != identity.  (BLS12-381 G1 has odd order: y != 0.)
GM_HD void xyzz_set_double_affine(XYZZ& acc, const Affine& p) {
  Fq u = p.y.dbl();
  Fq v = u.sqr();
  Fq w = u * v;
  Fq s = p.x * v;
  Fq xx = p.x.sqr();
  Fq m = xx.dbl() + xx;
  acc.x = m.sqr() - s.dbl();
  acc.y = m * (s - acc.x) - w * p.y;
  acc.zz = v;
  acc.zzz = w;
}

// acc = 2 * acc
GM_HD void xyzz_dbl(XYZZ& acc) {
  if (acc.is_identity()) return;
  Fq u = acc.y.dbl();
  Fq v = u.sqr();
  Fq w = u * v;
  Fq s = acc.x * v;
  Fq xx = acc.x.sqr();
  Fq m = xx.dbl() + xx;
  Fq x3 = m.sqr() - s.dbl();
  acc.y = m * (s - x3) - w * acc.y;
  acc.x = x3;
  acc.zz = v * acc.zz;
  acc.zzz = w * acc.zzz;
}

// acc += p (mixed addition, 8M + 2S on the generic path)
GM_HD void xyzz_madd(XYZZ& acc, const Affine& p) {
  if (p.is_identity()) return;
  if (acc.is_identity()) {
    acc.x = p.x; acc.y = p.y; acc.zz = Fq::one(); acc.zzz = Fq::one();
    return;
  }
  Fq h = p.x * acc.zz - acc.x;    // U2 - X1
  Fq r = p.y * acc.zzz - acc.y;   // S2 - Y1
  if (h.is_zero()) {
    if (r.is_zero()) xyzz_set_double_affine(acc, p);
    else acc = XYZZ::identity();
    return;
  }
  Fq hh = h.sqr();
  Fq hhh = h * hh;
  Fq q = acc.x * hh;
  Fq x3 = r.sqr() - hhh - q.dbl();
  acc.y = r * (q - x3) - acc.y * hhh;
  acc.x = x3;
  acc.zz = acc.zz * hh;
  acc.zzz = acc.zzz * hhh;
}

// acc += b (12M + 2S on the generic path)
GM_HD void xyzz_add(XYZZ& acc, const XYZZ& b) {
  if (b.is_identity()) return;
  if (acc.is_identity()) { acc = b; return; }
  Fq u1 = acc.x * b.zz;
  Fq u2 = b.x * acc.zz;
  Fq s1 = acc.y * b.zzz;
  Fq s2 = b.y * acc.zzz;
  Fq h = u2 - u1;
  Fq r = s2 - s1;
  if (h.is_zero()) {
    if (r.is_zero()) xyzz_dbl(acc);
    else acc = XYZZ::identity();
    return;
  }
  Fq hh = h.sqr();
  Fq hhh = h * hh;
  Fq q = u1 * hh;
  Fq x3 = r.sqr() - hhh - q.dbl();
  acc.y = r * (q - x3) - s1 * hhh;
  acc.x = x3;
  acc.zz = acc.zz * b.zz * hh;
  acc.zzz = acc.zzz * b.zzz * hhh;
}

GM_HD XYZZ xyzz_from_affine(const Affine& p) {
  XYZZ r = XYZZ::identity();
  if (!p.is_identity()) { r.x = p.x; r.y = p.y; r.zz = Fq::one(); r.zzz = Fq::one(); }
  return r;
}

GM_HD XYZZ xyzz_from_jacobian(const Jacobian& j) {
  XYZZ r;
  if (j.z.is_zero()) return XYZZ::identity();
  r.x = j.x; r.y = j.y; r.zz = j.z.sqr(); r.zzz = r.zz * j.z;
  return r;
}

// Canonical output: affine coordinates with Z = 1 (Montgomery one); identity as
// arkworks' Projective::zero() = (1, 1, 0).  Normalising on the device makes the
// 144-byte result independent of the order in which bucket sums were formed.
GM_HD Jacobian xyzz_to_jacobian_normalized(const XYZZ& a) {
  Jacobian j;
  if (a.is_identity()) { j.x = Fq::one(); j.y = Fq::one(); j.z = Fq::zero(); return j; }
  Fq izzz = fp_inv_serial(a.zzz);
  Fq t = a.zz * izzz;       // = ZZ / ZZZ = 1 / Z
  j.x = a.x * t.sqr();      // X / ZZ
  j.y = a.y * izzz;         // Y / ZZZ
  j.z = Fq::one();
  return j;
}

}  // namespace gm
