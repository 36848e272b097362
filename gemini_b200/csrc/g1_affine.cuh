// Affine point additions with a shared inversion (Montgomery's trick) for the bucket accumulation.
//
// The bucket sums of Pippenger (`buckets[i] += base`, /root/reference/src/kzg/msm/variable_base.rs:135-147) do
// not need projective coordinates when many independent additions are in flight: P3 = P1 + P2 in affine
// coordinates costs one inversion of (x2 - x1), and k inversions cost one inversion plus 3(k - 1) products.
// That is 6 products per addition (1 prefix + 2 back-substitution + lambda, lambda^2, y3) instead of the 10 of a
// mixed XYZZ addition.  The references of every bucket are reduced pairwise, level by level (a binary tree per
// bucket), and each level shares its inversions across all buckets.
//
// Everything here is complete for the degenerate inputs of the reference's default workloads: identity points
// ((0,0), kzg/time.rs:87), P + P (all-identical bases, examples/snark.rs:63) and P + (-P).
#pragma once
#include "g1.cuh"

namespace gm {

enum : uint32_t {
  PK_PASS1 = 0,  // result = first point  (no partner, or the partner is the identity)
  PK_PASS2 = 1,  // result = second point (the first one is the identity)
  PK_ADD = 2,    // generic chord: denominator x2 - x1
  PK_DBL = 3,    // tangent: denominator 2 y1
  PK_ZERO = 4    // P + (-P) = identity
};

// Classify the pair and return the denominator whose inverse the addition needs (ADD / DBL only).
GM_HD uint32_t aff_pair_kind(const Affine& p1, const Affine& p2, bool has2, Fq& den) {
  if (!has2 || p2.is_identity()) return PK_PASS1;
  if (p1.is_identity()) return PK_PASS2;
  den = p2.x - p1.x;
  if (!den.is_zero()) return PK_ADD;
  if (p1.y != p2.y) return PK_ZERO;
  den = p1.y.dbl();
  return den.is_zero() ? PK_ZERO : PK_DBL;  // y = 0 cannot happen on G1 (odd order); a 2-torsion point doubles to O
}

GM_HD bool aff_kind_needs_inverse(uint32_t kind) { return kind == PK_ADD || kind == PK_DBL; }

// inv_den = 1 / den(kind, p1, p2)
GM_HD Affine aff_pair_finish(uint32_t kind, const Affine& p1, const Affine& p2, const Fq& inv_den) {
  if (kind == PK_PASS1) return p1;
  if (kind == PK_PASS2) return p2;
  Affine r;
  if (kind == PK_ZERO) { r.x = Fq::zero(); r.y = Fq::zero(); return r; }
  Fq num;
  if (kind == PK_ADD) num = p2.y - p1.y;
  else { Fq xx = p1.x.sqr(); num = xx.dbl() + xx; }
  const Fq lam = num * inv_den;
  r.x = lam.sqr() - p1.x - p2.x;
  r.y = lam * (p1.x - r.x) - p1.y;
  return r;
}

}  // namespace gm
