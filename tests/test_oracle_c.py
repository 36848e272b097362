"""The C restatement (oracle/gemini_oracle.c) against the big-integer oracle (oracle/pyref.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import pyref as o
from util import R, rand_points, rand_scalars

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "libgemini_oracle.so")


@pytest.fixture(scope="module")
def clib():
    if not os.path.exists(SO):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    lib = C.CDLL(SO)
    lib.go_msm_g1.restype = C.c_int
    lib.go_msm_g1.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    lib.go_fr_fold.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.go_sumcheck_time.restype = C.c_size_t
    lib.go_sumcheck_time.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    return lib


def fr_l(vals):
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        a[i] = o.int_to_limbs(o.fr_to_mont(v % R), 4)
    return a


def fr_i(a):
    return [o.fr_from_mont(o.limbs_to_int(r)) for r in np.asarray(a).reshape(-1, 4)]


def g1_l(pts):
    a = np.zeros((len(pts), 12), dtype=np.uint64)
    for i, p in enumerate(pts):
        if p is not None:
            a[i, :6] = o.int_to_limbs(o.fq_to_mont(p[0]), 6)
            a[i, 6:] = o.int_to_limbs(o.fq_to_mont(p[1]), 6)
    return a


def c_msm(lib, bases, scalars, threads=4):
    b, s = g1_l(bases), fr_l(scalars)
    out = np.zeros(12, dtype=np.uint64)
    c = lib.go_msm_g1(b.ctypes.data, s.ctypes.data, min(len(bases), len(scalars)), 0, threads, out.ctypes.data)
    x, y = o.fq_from_mont(o.limbs_to_int(out[:6])), o.fq_from_mont(o.limbs_to_int(out[6:]))
    return (None if x == 0 and y == 0 else (x, y)), c


@pytest.mark.parametrize("n", [1, 5, 31, 32, 200, 1 << 10])
def test_c_msm_vs_pyref(clib, n):
    bases, scalars = rand_points(n, 40), rand_scalars(n, 41)
    got, c = c_msm(clib, bases, scalars)
    assert c == o.msm_window_size(n)
    assert got == o.msm_unchecked(bases, scalars)
    if n <= 200:
        assert got == o.naive_msm(bases, scalars)


def test_c_msm_degenerate(clib):
    pts = rand_points(20, 42)
    bases = pts + [None, pts[0], pts[0], o.g1_neg(pts[1])]
    scalars = rand_scalars(len(bases), 43)
    scalars[21] = scalars[0]
    scalars[23] = scalars[1]
    assert c_msm(clib, bases, scalars)[0] == o.naive_msm(bases, scalars)
    s = rand_scalars(1, 44)[0]
    assert c_msm(clib, pts * 5, [s] * 100)[0] == o.naive_msm(pts * 5, [s] * 100)
    assert c_msm(clib, [o.G1_GEN] * 64, [0, 1, R - 1, 1 << 254] * 16)[0] == o.g1_mul(o.G1_GEN, 16 * (R + (1 << 254)) % R)


@pytest.mark.parametrize("n", [1, 2, 3, 16, 19, 1000])
def test_c_fold(clib, n):
    f, r = rand_scalars(n, n), rand_scalars(1, 45)[0]
    fa, ra = fr_l(f), fr_l([r])
    out = np.zeros(((n + 1) // 2, 4), dtype=np.uint64)
    clib.go_fr_fold(fa.ctypes.data, n, ra.ctypes.data, out.ctypes.data)
    assert fr_i(out) == o.fold_polynomial(f, r)


@pytest.mark.parametrize("nf,ng", [(1, 1), (16, 16), (29, 29), (93, 16), (16, 93), (1000, 1000)])
def test_c_sumcheck(clib, nf, ng):
    f, g, tw = rand_scalars(nf, nf), rand_scalars(ng, ng + 1), rand_scalars(1, 46)[0]
    chals = rand_scalars(16, 47)
    it = iter(chals)
    msgs, used, final = o.sumcheck_prove(o.TimeProver(f, g, tw), lambda m: next(it))
    fa, ga, ta, ca = fr_l(f), fr_l(g), fr_l([tw]), fr_l(chals)
    out = np.zeros((16, 8), dtype=np.uint64)
    fin = np.zeros(8, dtype=np.uint64)
    rounds = clib.go_sumcheck_time(fa.ctypes.data, nf, ga.ctypes.data, ng, ta.ctypes.data, ca.ctypes.data, 16, out.ctypes.data, fin.ctypes.data)
    assert rounds == len(msgs)
    got = fr_i(out[:rounds])
    assert [(got[2 * i], got[2 * i + 1]) for i in range(rounds)] == msgs
    assert tuple(fr_i(fin)) == final


def test_native_adx_build_agrees_with_portable(clib):
    """`make -C oracle native` (mulx / adcx / adox products when the host has BMI2 + ADX - the build bench.py times as the
    CPU baseline) must return exactly what the portable build returns."""
    import subprocess

    odir = os.path.join(ROOT, "oracle")
    subprocess.run(["make", "-C", odir, "native"], check=True, stdout=subprocess.DEVNULL)
    nat = C.CDLL(os.path.join(odir, "_build", "libgemini_oracle_native.so"))
    nat.go_build_kind.restype = C.c_char_p
    nat.go_msm_g1.restype = C.c_int
    nat.go_msm_g1.argtypes = clib.go_msm_g1.argtypes
    nat.go_sumcheck_time.restype = C.c_size_t
    nat.go_sumcheck_time.argtypes = clib.go_sumcheck_time.argtypes
    n = 3000
    bases, scalars = g1_l(rand_points(n, 70)), fr_l(rand_scalars(n, 71))
    a, b = np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
    clib.go_msm_g1(bases.ctypes.data, scalars.ctypes.data, n, 0, 4, a.ctypes.data)
    nat.go_msm_g1(bases.ctypes.data, scalars.ctypes.data, n, 0, 4, b.ctypes.data)
    assert np.array_equal(a, b), nat.go_build_kind()
    f, g = fr_l(rand_scalars(1000, 72)), fr_l(rand_scalars(777, 73))
    tw, ch = fr_l(rand_scalars(1, 74)), fr_l(rand_scalars(12, 75))
    outs = []
    for lib in (clib, nat):
        msgs, fin = np.zeros((12, 8), dtype=np.uint64), np.zeros(8, dtype=np.uint64)
        k = lib.go_sumcheck_time(f.ctypes.data, 1000, g.ctypes.data, 777, tw.ctypes.data, ch.ctypes.data, 12, msgs.ctypes.data, fin.ctypes.data)
        outs.append((k, msgs.copy(), fin.copy()))
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
