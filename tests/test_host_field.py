"""The DEVICE limb schedules (gemini_b200/csrc/fp.cuh, g1.cuh) compiled for the host with an emulated
PTX carry flag (tests/csrc/host_check.cpp) and compared with Python big integers.  This is how the
Montgomery even/odd carry-chain product and the XYZZ formulas are validated before any GPU time."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import pyref as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hc():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    so = os.path.join(ROOT, "build", "libhostcheck.so")
    src = os.path.join(ROOT, "tests", "csrc", "host_check.cpp")
    subprocess.run(["g++", "-O2", "-frounding-math", "-x", "c++", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    return C.CDLL(so)


def pack(vals, n32):
    a = np.zeros((len(vals), n32), dtype=np.uint32)
    for i, v in enumerate(vals):
        for j in range(n32):
            a[i, j] = (v >> (32 * j)) & 0xFFFFFFFF
    return a


def unpack(a):
    return [sum(int(a[i, j]) << (32 * j) for j in range(a.shape[1])) for i in range(a.shape[0])]


def P(x):
    return C.c_void_p(x.ctypes.data)


@pytest.mark.parametrize("name,p,n", [("hc_fq", o.Q, 12), ("hc_fr", o.R, 8)])
def test_field_limb_schedules(hc, name, p, n):
    fn = getattr(hc, name)
    rng = random.Random(1)
    Rm = (1 << (32 * n)) % p
    Ri = pow(Rm, -1, p)
    edge = [0, 1, p - 1, p - 2, Rm, (p - 1) // 2, (1 << (32 * n - 3)) % p]
    A = edge + [rng.randrange(p) for _ in range(800)]
    B = [rng.randrange(p) for _ in range(len(A) - len(edge))] + edge
    a, b = pack(A, n), pack(B, n)
    r = np.zeros_like(a)
    exp = {0: lambda x, y: x * y * Ri % p, 1: lambda x, y: (x + y) % p, 2: lambda x, y: (x - y) % p,
           4: lambda x, y: x * Ri % p, 5: lambda x, y: x * Rm % p, 6: lambda x, y: x * x * Ri % p}
    for op, f in exp.items():
        fn(op, P(a), P(b), P(r), len(A))
        assert unpack(r) == [f(x, y) for x, y in zip(A, B)], (name, op)
    A2 = [x for x in A if x][:30]
    a2 = pack(A2, n)
    r2 = np.zeros_like(a2)
    fn(3, P(a2), P(a2), P(r2), len(A2))
    assert unpack(r2) == [pow(x * Ri % p, -1, p) * Rm % p for x in A2]


def fq_l(x):
    x = x * o.FQ_MONT_R % o.Q
    return [(x >> (32 * j)) & 0xFFFFFFFF for j in range(12)]


def l_fq(l):
    return sum(int(v) << (32 * j) for j, v in enumerate(l)) * pow(o.FQ_MONT_R, -1, o.Q) % o.Q


def xyzz(p, z=1):
    if p is None:
        return [0] * 48
    zz = z * z % o.Q
    zzz = zz * z % o.Q
    return fq_l(p[0] * zz % o.Q) + fq_l(p[1] * zzz % o.Q) + fq_l(zz) + fq_l(zzz)


def from_jac(row):
    x, y, z = l_fq(row[:12]), l_fq(row[12:24]), l_fq(row[24:])
    if z == 0:
        assert (x, y) == (1, 1)
        return None
    assert z == 1
    return (x, y)


def test_xyzz_formulas_complete(hc):
    rng = random.Random(2)
    pts = [o.g1_mul(o.G1_GEN, rng.randrange(1, o.R)) for _ in range(10)]
    cases = [(a, b) for a in pts[:5] for b in pts[5:]]
    cases += [(None, pts[0]), (pts[0], None), (None, None), (pts[1], pts[1]), (pts[2], o.g1_neg(pts[2]))]
    n = len(cases)
    aff = np.array([[0] * 24 if b is None else fq_l(b[0]) + fq_l(b[1]) for _, b in cases], dtype=np.uint32)
    for neg in (0, 1):
        acc = np.array([xyzz(a, rng.randrange(1, o.Q)) for a, _ in cases], dtype=np.uint32)
        hc.hc_xyzz_madd(P(acc), P(aff), neg, n)
        out = np.zeros((n, 36), dtype=np.uint32)
        hc.hc_xyzz_to_jacobian(P(acc), P(out), n)
        for i, (a, b) in enumerate(cases):
            assert from_jac(out[i]) == o.g1_add(a, o.g1_neg(b) if neg else b), (neg, i)
    acc = np.array([xyzz(a, rng.randrange(1, o.Q)) for a, _ in cases], dtype=np.uint32)
    oth = np.array([xyzz(b, rng.randrange(1, o.Q)) for _, b in cases], dtype=np.uint32)
    hc.hc_xyzz_add(P(acc), P(oth), n)
    out = np.zeros((n, 36), dtype=np.uint32)
    hc.hc_xyzz_to_jacobian(P(acc), P(out), n)
    for i, (a, b) in enumerate(cases):
        assert from_jac(out[i]) == o.g1_add(a, b), i
    acc = np.array([xyzz(a, rng.randrange(1, o.Q)) for a, _ in cases], dtype=np.uint32)
    hc.hc_xyzz_dbl(P(acc), n)
    hc.hc_xyzz_to_jacobian(P(acc), P(out), n)
    for i, (a, _) in enumerate(cases):
        assert from_jac(out[i]) == o.g1_double(a), i


def test_affine_pair_batch_shared_inversion(hc):
    """g1_affine.cuh: pair classification, chord / tangent formulas and the prefix-product walk that shares one
    inversion across a whole batch (the per-thread schedule of k_aff_prepare / k_aff_finish)."""
    rng = random.Random(7)
    pts = [o.g1_mul(o.G1_GEN, rng.randrange(1, o.R)) for _ in range(12)]
    cases = [(a, b, 1) for a in pts[:6] for b in pts[6:]]
    cases += [(None, pts[0], 1), (pts[0], None, 1), (None, None, 1), (pts[1], pts[1], 1), (pts[2], o.g1_neg(pts[2]), 1),
              (pts[3], pts[4], 0), (None, pts[4], 0), (pts[5], pts[5], 1)]
    rng.shuffle(cases)
    n = len(cases)
    enc = lambda p: [0] * 24 if p is None else fq_l(p[0]) + fq_l(p[1])
    p1 = np.array([enc(a) for a, _, _ in cases], dtype=np.uint32)
    p2 = np.array([enc(b) for _, b, _ in cases], dtype=np.uint32)
    has2 = np.array([h for _, _, h in cases], dtype=np.int32)
    out = np.zeros((n, 24), dtype=np.uint32)
    kinds = np.zeros(n, dtype=np.uint32)
    hc.hc_aff_batch(P(p1), P(p2), P(has2), n, P(out), P(kinds))
    for i, (a, b, h) in enumerate(cases):
        want = o.g1_add(a, b) if h else a
        x, y = l_fq(out[i, :12]), l_fq(out[i, 12:])
        got = None if (x, y) == (0, 0) else (x, y)
        assert got == want, (i, int(kinds[i]))
    assert set(int(k) for k in kinds) == {0, 1, 2, 3, 4}


def test_lazy_sum_of_products(hc):
    """FpAcc (fp.cuh): full 512-bit products accumulated with a conditional subtraction of the top half and ONE
    Montgomery reduction at the end - the schedule behind the sumcheck message sums."""
    rng = random.Random(11)
    p, n32 = o.R, 8
    Rm = (1 << 256) % p
    Ri = pow(Rm, -1, p)
    for n in (1, 2, 7, 300):
        A = [rng.randrange(p) for _ in range(n)]
        B = [rng.randrange(p) for _ in range(n)]
        if n == 7:
            A[:4] = [p - 1, p - 1, 0, 1]
            B[:4] = [p - 1, p - 2, p - 1, p - 1]
        a, b = pack(A, n32), pack(B, n32)
        out = np.zeros((1, n32), dtype=np.uint32)
        hc.hc_fr_sum_of_products(P(a), P(b), n, P(out))
        assert unpack(out)[0] == sum(x * y for x, y in zip(A, B)) * Ri % p
    # worst case for the invariant: many maximal products
    A = [p - 1] * 64
    a = pack(A, n32)
    out = np.zeros((1, n32), dtype=np.uint32)
    hc.hc_fr_sum_of_products(P(a), P(a), 64, P(out))
    assert unpack(out)[0] == 64 * (p - 1) * (p - 1) * Ri % p


@pytest.mark.parametrize("name,p,n", [("hc_fq", o.Q, 12), ("hc_fr", o.R, 8)])
def test_fast_inverse_divsteps(hc, name, p, n):
    """fp_inv_fast.cuh (staged for round 2): batched division steps, 30-bit limbs, against Python's pow(x, -1, p) and
    against the Kaliski inverse that the kernels use today."""
    fn = getattr(hc, name)
    rng = random.Random(21)
    Rm = (1 << (32 * n)) % p
    Ri = pow(Rm, -1, p)
    A = [1, 2, 3, p - 1, p - 2, Rm, (p - 1) // 2, (p + 1) // 2, (1 << (32 * n - 3)) % p, 1 << 29, 1 << 30, (1 << 30) - 1, 1 << 31]
    A += [rng.randrange(1, p) for _ in range(600)] + [rng.randrange(1, 1 << 64) for _ in range(50)] + [p - rng.randrange(1, 1 << 40) for _ in range(50)]
    a = pack(A, n)
    fast, slow = np.zeros_like(a), np.zeros_like(a)
    fn(8, P(a), P(a), P(fast), len(A))
    fn(3, P(a), P(a), P(slow), len(A))
    want = [pow(x * Ri % p, -1, p) * Rm % p for x in A]
    assert unpack(slow) == want
    assert unpack(fast) == want


# ---- Karatsuba / separated-operand-scanning Montgomery product (tools/fq_karatsuba.cuh, an experiment; detail::mul_full is the product of FpAcc) ------------------------
def _edge_ints(nl, rng, k):
    full = (1 << (32 * nl)) - 1
    e = [0, 1, full, full - 1, 1 << (32 * nl - 1), (1 << (16 * nl)) - 1, full ^ ((1 << (16 * nl)) - 1), 0xFFFFFFFF,
         full ^ 0xFFFFFFFF, int("0000ffff" * nl, 16), int("ffff0000" * nl, 16)]
    return e + [rng.randrange(full + 1) for _ in range(k)]


@pytest.mark.parametrize("which,nl", [(0, 4), (0, 6), (0, 8), (0, 12), (1, 8), (1, 12), (2, 8), (2, 12), (3, 3), (3, 5), (3, 6)])
def test_plain_products_on_arbitrary_limbs(hc, which, nl):
    rng = random.Random(100 * which + nl)
    E = _edge_ints(nl, rng, 0)
    A = [x for x in E for _ in E] + _edge_ints(nl, rng, 600)
    B = [y for _ in E for y in E] + _edge_ints(nl, rng, 600)[::-1]
    a, b = pack(A, nl), pack(B, nl)
    t = np.zeros((len(A), 2 * nl), dtype=np.uint32)
    hc.hc_mul_full(which, nl, P(a), P(b), P(t), len(A))
    assert unpack(t) == [x * y for x, y in zip(A, B)]


@pytest.mark.parametrize("is_fr,p,n", [(0, o.Q, 12), (1, o.R, 8)])
def test_redc_half_on_arbitrary_limbs(hc, is_fr, p, n):
    rng = random.Random(7 + is_fr)
    L = _edge_ints(n, rng, 1500) + [p, p - 1, p + 1, 2 * p, (1 << (32 * n)) - p]
    lo = pack(L, n)
    u = np.zeros_like(lo)
    hc.hc_redc_half(is_fr, P(lo), P(u), len(L))
    Ri = pow(1 << (32 * n), -1, p)
    got = unpack(u)
    assert all(g <= p for g in got)
    assert [g % p for g in got] == [x * Ri % p for x in L]


@pytest.mark.parametrize("name,p,n", [("hc_fq", o.Q, 12), ("hc_fr", o.R, 8)])
def test_karatsuba_montgomery_product(hc, name, p, n):
    fn = getattr(hc, name)
    rng = random.Random(11)
    Rm = (1 << (32 * n)) % p
    Ri = pow(Rm, -1, p)
    edge = [0, 1, p - 1, p - 2, Rm, (p - 1) // 2, (1 << (32 * n - 3)) % p, (1 << (16 * n)) - 1, (1 << (16 * n))]
    A = [x for x in edge for _ in edge] + [rng.randrange(p) for _ in range(3000)]
    B = [y for _ in edge for y in edge] + [rng.randrange(p) for _ in range(3000)]
    a, b = pack(A, n), pack(B, n)
    r = np.zeros_like(a)
    for op in (9, 10):   # one and two Karatsuba levels
        fn(op, P(a), P(b), P(r), len(A))
        assert unpack(r) == [x * y * Ri % p for x, y in zip(A, B)], op


def test_fq_product_on_the_fp64_pipe_48_bit_limbs(hc):
    """tools/fq_f64v2.cuh (experiment): the Fq Montgomery product from 48-bit limbs in doubles, every partial product split
    exactly by two fused multiply-adds (one rounding toward zero).  Same limbs in, same limbs out as mont_mul."""
    p, n = o.Q, 12
    rng = random.Random(48)
    Rm = (1 << 384) % p
    Ri = pow(Rm, -1, p)
    lim = (1 << 48) - 1
    edge = [0, 1, p - 1, p - 2, Rm, (p - 1) // 2, (1 << 381) % p, (1 << 380), lim, lim << 48, (lim << 336) % p,
            sum(lim << (48 * k) for k in range(7)), sum(1 << (48 * k) for k in range(8)) % p, (1 << 192) - 1, 1 << 192]
    A = [x for x in edge for _ in edge] + [rng.randrange(p) for _ in range(4000)]
    B = [y for _ in edge for y in edge] + [rng.randrange(p) for _ in range(4000)]
    a, b = pack(A, n), pack(B, n)
    r = np.zeros_like(a)
    hc.hc_fq(11, P(a), P(b), P(r), len(A))
    assert unpack(r) == [x * y * Ri % p for x, y in zip(A, B)]


def test_unreduced_run_bound_of_the_lazy_accumulator():
    """FpAcc::mul_add_unreduced (fp.cuh): with the top half below p on entry, k maximal products keep the accumulator below
    2 p 2^(32N) - no overflow of the 2N limbs, and ONE conditional subtraction of the top half restores the invariant -
    exactly while k p < 2^(32N).  The constants in the header (2 for Fr, 9 for Fq) must be the largest such k."""
    for p, n, run in ((o.R, 8, 2), (o.Q, 12, 9)):
        Rr = 1 << (32 * n)
        worst = lambda k: (p - 1) * Rr + (Rr - 1) + k * (p - 1) ** 2        # top half p - 1, bottom half all ones
        assert run * p < Rr <= (run + 1) * p
        assert worst(run) < 2 * p * Rr < Rr * Rr
        assert worst(run + 1) >= 2 * p * Rr or (run + 1) * p >= Rr
