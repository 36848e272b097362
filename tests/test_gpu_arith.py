"""Device field / curve primitives vs Python big integers (through the C ABI self-test entry points)."""
import ctypes as C
import random

import numpy as np
import pytest

import pyref as o
from gemini_b200._lib import check, lib

pytestmark = pytest.mark.gpu


def pack32(vals, n32):
    a = np.zeros((len(vals), n32), dtype=np.uint32)
    for i, v in enumerate(vals):
        for j in range(n32):
            a[i, j] = (v >> (32 * j)) & 0xFFFFFFFF
    return a


def unpack32(a):
    return [sum(int(a[i, j]) << (32 * j) for j in range(a.shape[1])) for i in range(a.shape[0])]


def P(x):
    return C.c_void_p(x.ctypes.data)


@pytest.mark.parametrize("fid,p,n32", [(0, o.Q, 12), (1, o.R, 8)])
def test_field_ops(ctx, fid, p, n32):
    rng = random.Random(11 + fid)
    Rm = (1 << (32 * n32)) % p
    Ri = pow(Rm, -1, p)
    edge = [0, 1, p - 1, p - 2, Rm, (p - 1) // 2, (1 << (32 * n32 - 3)) % p]
    A = edge + [rng.randrange(p) for _ in range(3000)]
    B = [rng.randrange(p) for _ in range(len(A) - len(edge))] + edge
    a, b = pack32(A, n32), pack32(B, n32)
    r = np.zeros_like(a)
    expect = {
        0: lambda x, y: x * y * Ri % p, 1: lambda x, y: (x + y) % p, 2: lambda x, y: (x - y) % p,
        4: lambda x, y: x * Ri % p, 5: lambda x, y: x * Rm % p, 6: lambda x, y: x * x * Ri % p,
    }
    for op, f in expect.items():
        check(lib.gm_selftest_field(ctx._h, fid, op, P(a), P(b), P(r), len(A)))
        assert unpack32(r) == [f(x, y) for x, y in zip(A, B)], f"field {fid} op {op}"
    A2 = [x for x in A if x][:64]
    a2 = pack32(A2, n32)
    r2 = np.zeros_like(a2)
    check(lib.gm_selftest_field(ctx._h, fid, 3, P(a2), P(a2), P(r2), len(A2)))
    assert unpack32(r2) == [pow(x * Ri % p, -1, p) * Rm % p for x in A2]
    # op 7: the division-step inverse used by k_aff_invert / k_normalize, on many more inputs
    A3 = [x for x in A if x]
    a3 = pack32(A3, n32)
    r3 = np.zeros_like(a3)
    check(lib.gm_selftest_field(ctx._h, fid, 7, P(a3), P(a3), P(r3), len(A3)))
    assert unpack32(r3) == [pow(x * Ri % p, -1, p) * Rm % p for x in A3]


def _fq_l(x):
    x = x * o.FQ_MONT_R % o.Q
    return [(x >> (32 * j)) & 0xFFFFFFFF for j in range(12)]


def _l_fq(l):
    return sum(int(v) << (32 * j) for j, v in enumerate(l)) * pow(o.FQ_MONT_R, -1, o.Q) % o.Q


def _aff(p):
    return [0] * 24 if p is None else _fq_l(p[0]) + _fq_l(p[1])


def _xyzz(p, z=1):
    if p is None:
        return [0] * 48
    zz = z * z % o.Q
    zzz = zz * z % o.Q
    return _fq_l(p[0] * zz % o.Q) + _fq_l(p[1] * zzz % o.Q) + _fq_l(zz) + _fq_l(zzz)


def _from_jac(row):
    x, y, z = _l_fq(row[:12]), _l_fq(row[12:24]), _l_fq(row[24:])
    if z == 0:
        assert (x, y) == (1, 1)
        return None
    assert z == 1
    return (x, y)


def test_curve_ops(ctx):
    rng = random.Random(5)
    pts = [o.g1_mul(o.G1_GEN, rng.randrange(1, o.R)) for _ in range(16)]
    cases = [(a, b) for a in pts[:8] for b in pts[8:]]
    cases += [(None, pts[0]), (pts[0], None), (None, None), (pts[1], pts[1]), (pts[2], o.g1_neg(pts[2]))]
    n = len(cases)
    other_aff = np.array([_aff(b) for _, b in cases], dtype=np.uint32)
    for op in (0, 1, 2, 3):
        acc = np.array([_xyzz(a, rng.randrange(1, o.Q)) for a, _ in cases], dtype=np.uint32)
        if op == 2:
            other = np.array([_xyzz(b, rng.randrange(1, o.Q)) for _, b in cases], dtype=np.uint32)
        else:
            other = other_aff
        out = np.zeros((n, 36), dtype=np.uint32)
        check(lib.gm_selftest_curve(ctx._h, op, P(acc), P(other), P(out), n))
        for i, (a, b) in enumerate(cases):
            want = {0: o.g1_add(a, b), 1: o.g1_add(a, o.g1_neg(b)), 2: o.g1_add(a, b), 3: o.g1_double(a)}[op]
            assert _from_jac(out[i]) == want, (op, i)
