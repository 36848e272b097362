"""Independent pin of the oracle's G1 group law.

The reference holds no golden G1 vector (its own MSM tests, src/kzg/msm/variable_base.rs:179-215, live in orphaned
files that never compile) and cannot be built here, so oracle/pyref.py's `g1_add / g1_double / g1_mul / naive_msm`
were until now only self-checked ([r]G = O, on-curve).  sympy ships an unrelated implementation of elliptic-curve
arithmetic over prime fields (sympy.ntheory.elliptic_curve.EllipticCurve: projective formulas written by other
people for other purposes); here BLS12-381 G1 (y^2 = x^3 + 4 over F_q) is instantiated in it and every oracle
primitive is compared with it on random points, the edge cases included.  With this the chain
    CUDA path == oracle (tests/test_gpu_*) == independent implementation
is closed for the group law; field arithmetic is Python's own integers on both sides."""
import random

import pytest

import pyref as o

sympy_ec = pytest.importorskip("sympy.ntheory.elliptic_curve")

E = sympy_ec.EllipticCurve(0, 4, modulus=o.Q)
ZERO = sympy_ec.EllipticCurvePoint.point_at_infinity(E)


def to_s(p):
    return ZERO if p is None else E(p[0], p[1])


def from_s(p):
    if int(p.z) == 0:
        return None
    zi = pow(int(p.z), -1, o.Q)
    return (int(p.x) * zi % o.Q, int(p.y) * zi % o.Q)


@pytest.fixture(scope="module")
def points():
    rng = random.Random(20261017)
    ks = [rng.randrange(1, o.R) for _ in range(50)]
    return ks, [o.g1_mul(o.G1_GEN, k) for k in ks]


def test_generator_and_order():
    assert o.g1_is_on_curve(o.G1_GEN)
    g = to_s(o.G1_GEN)
    assert from_s(o.R * g) is None and o.g1_mul(o.G1_GEN, o.R) is None
    assert from_s((o.R - 1) * g) == o.g1_neg(o.G1_GEN) == o.g1_mul(o.G1_GEN, o.R - 1)


def test_scalar_multiplication(points):
    ks, pts = points
    g = to_s(o.G1_GEN)
    for k, p in zip(ks[:25], pts[:25]):
        assert from_s(k * g) == p
    for k in (0, 1, 2, 3, (1 << 254), o.R - 2, (1 << 255) % o.R):
        assert from_s(k * g) == o.g1_mul(o.G1_GEN, k)
    # multiples of a non-generator point
    base = pts[7]
    for k in ks[:5]:
        assert from_s(k * to_s(base)) == o.g1_mul(base, k)


def test_addition_doubling_negation(points):
    _, pts = points
    for a, b in zip(pts[:-1], pts[1:]):
        assert from_s(to_s(a) + to_s(b)) == o.g1_add(a, b)
    for a in pts[:30]:
        assert from_s(to_s(a) + to_s(a)) == o.g1_double(a) == o.g1_add(a, a)
        assert o.g1_add(a, o.g1_neg(a)) is None and from_s(to_s(a) + to_s(o.g1_neg(a))) is None
        assert o.g1_add(a, None) == a and o.g1_add(None, a) == a
    assert o.g1_add(None, None) is None and o.g1_double(None) is None


def test_naive_and_pippenger_msm(points):
    ks, pts = points
    rng = random.Random(5)
    for n in (1, 2, 7, 24):
        sc = [rng.randrange(o.R) for _ in range(n)]
        acc = ZERO
        for s, p in zip(sc, pts[:n]):
            acc = acc + s * to_s(p)
        want = from_s(acc)
        assert o.naive_msm(pts[:n], sc) == want
        assert o.pippenger_msm(pts[:n], sc) == want      # arkworks' signed-digit windows (variable_base.rs:95-177)
    # degenerate inputs of the reference's default workloads: equal scalars, equal bases, identity bases, cancellation
    s = rng.randrange(o.R)
    acc = ZERO
    for p in pts[:20]:
        acc = acc + to_s(p)
    assert o.naive_msm(pts[:20], [s] * 20) == from_s(s * acc)
    assert o.naive_msm([pts[3]] * 9, list(range(1, 10))) == from_s(45 * to_s(pts[3]))
    assert o.naive_msm([None, pts[1], None], [5, 6, 7]) == from_s(6 * to_s(pts[1]))
    assert o.naive_msm([pts[2], o.g1_neg(pts[2])], [11, 11]) is None


def test_golden_msm_vectors_against_independent_curve():
    """the frozen MSM vectors of tests/golden/hotpath_v1.json (what the CUDA path is held to) recomputed with sympy"""
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_v1.json")
    with open(path) as fh:
        gold = json.load(fh)
    cases = gold["msm"] if isinstance(gold.get("msm"), list) else [v for k, v in gold.items() if k.startswith("msm")]
    checked = 0
    for case in cases:
        if not isinstance(case, dict) or "bases" not in case:
            continue
        bases = [None if b is None else (int(b[0], 16) if isinstance(b[0], str) else b[0], int(b[1], 16) if isinstance(b[1], str) else b[1])
                 for b in case["bases"]]
        scalars = [int(s, 16) if isinstance(s, str) else s for s in case["scalars"]]
        n = min(len(bases), len(scalars))
        if n > 40:
            continue
        acc = ZERO
        for s, p in zip(scalars[:n], bases[:n]):
            acc = acc + (s % o.R) * to_s(p)
        want = case["result"]
        want = None if want is None else (int(want[0], 16) if isinstance(want[0], str) else want[0],
                                          int(want[1], 16) if isinstance(want[1], str) else want[1])
        assert from_s(acc) == want, case.get("name")
        checked += 1
    assert checked >= 3
