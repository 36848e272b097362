"""Pins the big-integer oracle (oracle/pyref.py) to every known-answer value and differential
property the reference's own tests hold for this path."""
import random

import pytest

import pyref as o

R = o.R


def rs(n, seed):
    rng = random.Random(seed)
    return [rng.randrange(R) for _ in range(n)]


def test_curve_self_check():
    assert o.g1_is_on_curve(o.G1_GEN)
    assert o.g1_mul(o.G1_GEN, R) is None          # [r]G = O
    assert o.g1_add(o.G1_GEN, o.g1_neg(o.G1_GEN)) is None
    assert o.g1_mul(o.G1_GEN, 5) == o.g1_add(o.g1_double(o.g1_double(o.G1_GEN)), o.G1_GEN)
    assert o.FQ_INV64 == 0x89F3FFFCFFFCFFFD and o.FR_INV64 == 0xFFFFFFFEFFFFFFFF  # SURVEY 8c constants
    assert o.FR_MONT_R == 0x1824B159ACC5056F998C4FEFECBC4FF55884B7FA0003480200000001FFFFFFFE


def test_window_sizes_match_survey():
    # n=2^12->10 (26 win), 2^16->13 (20), 2^20->15 (17), 2^24->18 (15), 2^28->21 (13)
    for logn, c, w in [(12, 10, 26), (16, 13, 20), (20, 15, 17), (24, 18, 15), (28, 21, 13)]:
        assert o.msm_window_size(1 << logn) == c
        assert (255 + c - 1) // c == w
    assert o.msm_window_size(31) == 3
    assert o.ark_log2(17) == 5  # time_prover.rs:154-158


def test_radix_digits_recompose():
    """variable_base.rs:63-93 test_radix."""
    for w in (3, 10, 15, 21):
        for a in rs(20, w) + [0, 1, R - 1]:
            d = o.make_digits(a, w, 0 if a else 255)
            assert sum(x << (i * w) for i, x in enumerate(d)) == a
            d = o.make_digits(a, w, 255)
            assert sum(x << (i * w) for i, x in enumerate(d)) == a
            assert all(-(1 << (w - 1)) <= x for x in d[:-1]) and all(x < (1 << (w - 1)) for x in d[:-1])


def test_pippenger_vs_naive():
    """variable_base.rs:179-215."""
    rng = random.Random(1)
    bases = [o.g1_mul(o.G1_GEN, rng.randrange(1, R)) for _ in range(40)]
    for n in (1, 7, 31, 40):
        sc = rs(n, n)
        assert o.pippenger_msm(bases[:n], sc) == o.naive_msm(bases[:n], sc)
    # chunked / hash-map variants (stream_pippenger.rs:356-425)
    sc = rs(40, 99)
    cp = o.ChunkedPippenger(7)
    hp = o.HashMapPippenger(5)
    for b, s in zip(bases, sc):
        cp.add(b, s)
        hp.add(b, s)
    assert cp.finalize() == hp.finalize() == o.naive_msm(bases, sc)


def test_fold_stream_kats():
    """sumcheck/streams.rs:232-288."""
    assert list(o.folded_polynomial_stream([1, 2, 1, 1], [1, 2])) == [8]
    out = list(o.folded_polynomial_stream([1] * 12, [1, 1, 1, 1]))
    assert out[-1] == 12
    tree = list(o.folded_polynomial_tree([1, 1, 2], [1, 1])) if False else None
    assert list(o.folded_polynomial_tree([1, 2, 1, 1], [1, 2])) == [(1, 3), (1, 2), (2, 8)]
    # tensorcheck/mod.rs:388-398
    assert o.foldings_polynomial([100, 101, 102, 103], [1, 1]) == [[201, 205]]


@pytest.mark.parametrize("n,levels", [(16, 1), (16, 3), (19, 1), (19, 2), (19, 3)])
def test_fold_polynomial_vs_stream(n, levels):
    """sumcheck/tests.rs:140-200 (sizes 16 and 19, 1-3 levels, zero padding of non powers of two)."""
    f, ch = rs(n, n), rs(levels, 7)
    cur = f
    for c in ch:
        cur = o.fold_polynomial(cur, c)
    assert list(o.folded_polynomial_stream(f[::-1], ch))[::-1] == cur


def test_open_multi_points_kat():
    """kzg/space.rs:321-387: remainder(53) = 1807299544171."""
    rng = random.Random(3)
    srs = [o.g1_mul(o.G1_GEN, rng.randrange(1, R)) for _ in range(10)]
    rem, _ = o.kzg_stream_open_multi_points(srs[::-1], [80, 80, 88, 3, 73, 7, 24], [53 * 53, 53, R - 53], 4)
    assert o.evaluate_be(rem, 53) == 1807299544171


@pytest.mark.parametrize("nf,ng", [(30, 30), (93, 16), (16, 16), (29, 29)])
def test_time_vs_space_vs_elastic(nf, ng):
    """sumcheck/tests.rs:41-138."""
    f, g, tw = rs(nf, 1), rs(ng, 2), rs(1, 3)[0]
    ch = rs(12, 4)
    runs = []
    for mk in (lambda: o.TimeProver(f, g, tw), lambda: o.SpaceProver(f[::-1], g[::-1], tw), lambda: o.ElasticProver(f[::-1], g[::-1], tw)):
        it = iter(ch)
        runs.append(o.sumcheck_prove(mk(), lambda m: next(it)))
    nmin = min(len(r[0]) for r in runs)
    if nf == ng:
        assert runs[0][0] == runs[1][0] == runs[2][0]
        if nf & (nf - 1) == 0:
            assert runs[0][2] == runs[1][2]
    else:
        assert runs[1][0] == runs[2][0]
        assert runs[0][0][:3] == runs[1][0][:3]  # first three rounds, tests.rs:113-138
    assert nmin >= 3


def test_sumcheck_completeness():
    """sumcheck/tests.rs:202-224: n = 2^10+1, random twist; messages satisfy subclaim.rs:77-97."""
    n = (1 << 10) + 1
    f, g, tw = rs(n, 5), rs(n, 6), rs(1, 7)[0]
    it = iter(rs(16, 8))
    msgs, ch, final = o.sumcheck_prove(o.TimeProver(f, g, tw), lambda m: next(it))
    asserted = sum(a * b * pow(tw, i, R) for i, (a, b) in enumerate(zip(f, g))) % R
    assert o.subclaim_reduce(msgs, ch, asserted) == final[0] * final[1] % R


def test_kzg_time_equals_space():
    """kzg/tests.rs:15-59."""
    rng = random.Random(9)
    d = 15
    srs = [o.g1_mul(o.G1_GEN, rng.randrange(1, R)) for _ in range(d + 4)]
    poly = rs(d + 1, 10)
    assert o.kzg_commit(srs, poly) == o.kzg_stream_commit(srs[::-1], poly[::-1])
    alpha = rs(1, 11)[0]
    ev_t, pf_t = o.kzg_open(srs, poly, alpha)
    ev_s, pf_s = o.kzg_stream_open(srs[::-1], poly[::-1], alpha, 4)
    assert (ev_t, pf_t) == (ev_s, pf_s) and ev_t == o.evaluate_le(poly, alpha)
    ch = rs(3, 12)
    folds = o.foldings_polynomial(poly, ch + [0])
    assert o.kzg_commit_folding(srs[::-1], poly[::-1], ch, 20) == [o.kzg_commit(srs, f) for f in folds]


def test_open_folding_restatement_matches_time_side_algebra():
    """kzg/space.rs:229-285 (streaming, HashMapPippenger) against the textbook statement: per fold level the quotient and
    remainder of f^(i) by the vanishing polynomial; proof = sum_i eta_i * commit(quotient_i)."""
    import random

    from util import rand_points

    rng = random.Random(99)
    for n, depth, npts in ((37, 4, 3), (16, 4, 3), (5, 3, 3), (9, 2, 1)):
        coeffs_le = [rng.randrange(o.R) for _ in range(n)]
        challenges = [rng.randrange(o.R) for _ in range(depth)]
        points = [rng.randrange(o.R) for _ in range(npts)]
        etas = [rng.randrange(o.R) for _ in range(depth)]
        srs_le = rand_points(n + 3, 123)
        rem, proof = o.kzg_open_folding(srs_le[::-1], coeffs_le[::-1], challenges, points, etas, 1 << 10)
        z = o.vanishing_polynomial(points)
        want_proof = None
        f = coeffs_le
        for i in range(depth):
            f = o.fold_polynomial(f, challenges[i])
            q = o.poly_div(f, z)
            # remainder = f - q z
            qz = [0] * max(len(f), len(q) + len(z) - 1 if q else 0)
            for a, qa in enumerate(q):
                for b, zb in enumerate(z):
                    qz[a + b] = (qz[a + b] + qa * zb) % o.R
            r = [(f[d] - (qz[d] if d < len(qz) else 0)) % o.R if d < len(f) else 0 for d in range(npts)]
            assert rem[i] == r[::-1], (n, i)
            if q:
                want_proof = o.g1_add(want_proof, o.g1_mul(o.naive_msm(srs_le, q), etas[i]))
        assert proof == want_proof


def test_linear_combination_kat():
    """src/misc.rs:402-421"""
    polys = [[100, 101, 102, 103], [100, 100, 100, 100]]
    assert o.linear_combination(polys, [1, 10]) == [1100, 1101, 1102, 1103]
    assert o.linear_combination([], []) == []


def test_matrix_tensor_stream_kats():
    """src/snark/streams.rs:106-133 and :177-222 (random field elements replaced by seeded ones)"""
    import random

    rng = random.Random(12)
    r = rng.randrange(o.R)
    assert list(o.matrix_tensor_stream(o.diagonal_matrix_stream(r, 4), [1, 1])) == [r, r, r, r]
    t0, t1 = rng.randrange(o.R), rng.randrange(o.R)
    assert list(o.matrix_tensor_stream(o.diagonal_matrix_stream(r, 4), [t0, t1])) == [r * t0 * t1 % o.R, r * t1 % o.R, r * t0 % o.R, r]
    E = o.EOL
    matrix = [(1, 0), (1, 1), (1, 2), (1, 3), E, (1, 1), E, (1, 2), E, E]
    got = list(o.matrix_tensor_stream(matrix, [r, r * r % o.R]))
    assert got == [(r * r * r + r * r + r + 1) % o.R, r, r * r % o.R, 0]
    ch = [rng.randrange(o.R) for _ in range(4)]
    got = list(o.matrix_tensor_stream(o.diagonal_matrix_stream(1, 16), ch))
    assert got[::-1] == o.tensor(ch)


def test_lincomb_stream():
    """src/subprotocols/tensorcheck/streams.rs:249-290 plus the alignment rule of LinCombStream::iter (pads)"""
    ones = [[1] * 100 for _ in range(20)]
    assert o.lincomb_stream(ones, [0] * 20)[0] == 0
    # unequal lengths: big-endian streams line up at the low-degree end = linear_combination of the reversed vectors
    a, b, c = [5, 6, 7, 8], [1, 2], [9]
    got = o.lincomb_stream([a, b, c], [2, 3, 4])
    assert got[::-1] == o.linear_combination([a[::-1], b[::-1], c[::-1]], [2, 3, 4])
    assert got == [10, 12, 14 + 3, 16 + 6 + 36]


class _HashTranscript:
    """Deterministic Fiat-Shamir stand-in (the Merlin restatement lives in the product and has its own vectors)."""

    def __init__(self):
        import hashlib

        self._hashlib = hashlib
        self.h = hashlib.sha256(b"test-transcript")

    def append_serializable(self, label, obj):
        self.h.update(label + repr(obj).encode())

    def append_g1(self, label, point):
        self.h.update(label + b"G1" + repr(point).encode())

    def get_challenge(self, label):
        self.h.update(b"challenge" + label)
        return int.from_bytes(self._hashlib.sha512(self.h.digest()).digest(), "little") % o.R


@pytest.mark.parametrize("rows,cols,seed", [(8, 16, 1), (8, 8, 2), (5, 7, 3), (16, 16, 4)])
def test_snark_time_proof_equals_elastic_proof(rows, cols, seed):
    """The reference's strongest test, snark/tests.rs:13-58 (`assert_eq!(time_proof, space_proof)`), on random sparse
    matrices: it ties the streaming restatements (ElasticProver, stream commit, MatrixTensor, LinCombStream,
    commit_folding, evaluate_folding, open_multi_points, open_folding) to the time-side ones."""
    import random

    from util import rand_points

    rng = random.Random(seed)

    def matrix():
        m = []
        for _ in range(rows):
            cs = sorted(rng.sample(range(cols), rng.randrange(1, min(4, cols) + 1)))
            m.append([(rng.randrange(1, o.R), c) for c in cs])
        return m

    z = [rng.randrange(o.R) for _ in range(cols)]
    r1cs = {"a": matrix(), "b": matrix(), "c": matrix(), "z": z, "w": z[cols // 2:], "x": z[:cols // 2]}
    srs = rand_points(rows + cols + 1, 500 + seed)
    time_proof = o.snark_new_time(r1cs, srs, _HashTranscript())
    elastic_proof = o.snark_new_elastic(r1cs, srs, _HashTranscript(), 20)
    assert elastic_proof == time_proof
