"""Device-resident side of the KZG openings, the space / elastic provers and the stream adaptors (SURVEY.md 8 rows
a2, a8, f3): same results as the host-fed paths and the oracle, with the polynomials never leaving HBM.

Reference patterns: time open == space open (src/kzg/tests.rs:42-59), MatrixTensor against the materialised product
(src/snark/streams.rs:106-222), LinCombStream alignment (src/subprotocols/tensorcheck/streams.rs:249-290), elastic
prover == time prover (src/subprotocols/sumcheck/tests.rs:41-224), and - at sizes the oracle cannot reach - the
pairing-free consequence of the opening equation for a key with known tau: proof = [(f(tau) - f(x)) / (tau - x)] g."""
import random

import numpy as np
import pytest

import gemini_b200 as gm
import pyref as o
from gemini_b200 import field, kzg
from gemini_b200.devvec import DeviceCsr, DeviceFr
from gemini_b200.streams import LinCombStream, MatrixTensor, ReverseStream
from util import R, fr_random_limbs, limbs_to_ints, rand_points, rand_scalars

pytestmark = pytest.mark.gpu


def test_reverse_on_device(ctx):
    for n in (0, 1, 2, 7, 1000, 4097):
        vals = rand_scalars(n, n)
        v = DeviceFr.from_host(ctx, vals)
        assert v.reversed().to_ints() == vals[::-1]
        assert v.reverse_().to_ints() == vals[::-1]


def test_open_variants_agree_with_oracle(ctx):
    srs = rand_points(40, 70)
    ck = gm.CommitterKey(ctx, srs)
    cks = gm.CommitterKeyStream(ctx, srs[::-1])
    pts = rand_scalars(3, 71)
    for n in (1, 2, 3, 4, 17, 33):
        poly = rand_scalars(n, 72 + n)
        dv = DeviceFr.from_host(ctx, poly)
        alpha = rand_scalars(1, 73)[0]
        want = o.kzg_open(srs, poly, alpha)
        assert ck.open(poly, alpha) == want and ck.open(dv, alpha) == want
        assert cks.open(poly[::-1], alpha, 4) == want and cks.open(ReverseStream(dv), alpha, 4) == want
        want_mp = o.kzg_open_multi_points(srs, poly, pts)
        assert ck.open_multi_points(poly, pts) == want_mp == ck.open_multi_points(dv, pts)
        if n >= 3:
            rem_o, proof_o = o.kzg_stream_open_multi_points(srs[::-1], poly[::-1], pts, 5)
            assert cks.open_multi_points(poly[::-1], pts, 5) == (rem_o, proof_o)
            assert cks.open_multi_points(ReverseStream(dv), pts, 5) == (rem_o, proof_o)
        assert ck.commit(dv) == o.kzg_commit(srs, poly) == cks.commit(ReverseStream(dv))
    assert ck.open([], 5) == (0, None)


def test_device_chunking_does_not_change_results(ctx, monkeypatch):
    """the device pushes walk real chunk boundaries when the floor on the chunk size is lowered"""
    srs = rand_points(300, 74)
    cks = gm.CommitterKeyStream(ctx, srs[::-1])
    poly = rand_scalars(257, 75)
    dv = DeviceFr.from_host(ctx, poly)
    want = o.kzg_commit(srs, poly)
    chals = rand_scalars(5, 76)
    want_fold = o.kzg_commit_folding(srs[::-1], poly[::-1], chals, 20)
    for floor in (1, 7, 64, 1 << 27):
        monkeypatch.setattr(kzg, "MIN_DEVICE_CHUNK", floor)
        assert cks.commit(ReverseStream(dv), 3) == want
        assert cks.commit_folding(ReverseStream(dv), chals, 20) == want_fold
        assert cks.commit_folding(poly[::-1], chals, 20) == want_fold


def test_opening_equation_at_2_18(ctx):
    """size-independent property: with powers_of_g[i] = tau^i g the proof of open(f, x) is [(f(tau) - f(x)) / (tau - x)] g
    and the proof of open_multi_points is [q(tau)] g with f = q * Z + rem"""
    n = 1 << 18
    ck = gm.CommitterKey.new(ctx, n - 1, 3, random.Random(11), precompute=True)
    f = DeviceFr.random(ctx, n, 1234)
    x = rand_scalars(1, 77)[0]
    f_tau, f_x = f.evaluate(ck.tau), f.evaluate(x)
    ev, proof = ck.open(f, x)
    assert ev == f_x
    assert proof == o.g1_mul(ck.g, (f_tau - f_x) * pow(ck.tau - x, -1, R) % R)
    cks = gm.CommitterKeyStream.from_committer_key(ck)
    ev_s, proof_s = cks.open(ReverseStream(f), x, 1 << 16)
    assert (ev_s, proof_s) == (ev, proof)
    pts = rand_scalars(3, 78)
    rem, proof_mp = cks.open_multi_points(ReverseStream(f), pts, 1 << 16)
    z_tau = 1
    for p in pts:
        z_tau = z_tau * (ck.tau - p) % R
    rem_tau = o.evaluate_be(rem, ck.tau)
    assert proof_mp == o.g1_mul(ck.g, (f_tau - rem_tau) * pow(z_tau, -1, R) % R) == ck.open_multi_points(f, pts)
    for p in pts:   # the remainder interpolates f on the points
        assert o.evaluate_be(rem, p) == f.evaluate(p)
    # batch opening of polynomials of different lengths
    g = DeviceFr.random(ctx, n // 2 + 5, 99)
    eta = rand_scalars(1, 79)[0]
    comb_tau = (f_tau + eta * g.evaluate(ck.tau)) % R
    got = ck.batch_open_multi_points([f, g], pts, eta)
    comb = DeviceFr.zeros(ctx, n)
    comb.axpy(1, f)
    comb.axpy(eta, g)
    rem_c, _ = cks.open_multi_points(ReverseStream(comb), pts, 1 << 16)
    assert got == o.g1_mul(ck.g, (comb_tau - o.evaluate_be(rem_c, ck.tau)) * pow(z_tau, -1, R) % R)


def test_matrix_tensor_and_lincomb_streams(ctx):
    rng = random.Random(5)
    rows, cols = 13, 16
    m = []
    for _ in range(rows):
        cs = sorted(rng.sample(range(cols), rng.randrange(1, 5)))
        m.append([(rng.randrange(1, R), c) for c in cs])
    v = rand_scalars(4, 80)                      # 2^4 >= rows
    mt = MatrixTensor(ctx, DeviceCsr.from_rows(ctx, m, cols, transpose=True), v)
    want = list(o.matrix_tensor_stream(o.matrix_into_colmaj(m, cols), v))
    assert len(mt) == cols and mt.to_ints_be() == want
    s1, s2, s3 = rand_scalars(16, 81), rand_scalars(9, 82), rand_scalars(12, 83)
    coeffs = rand_scalars(3, 84)
    streams = [ReverseStream(DeviceFr.from_host(ctx, s[::-1])) for s in (s1, s2, s3)]
    lc = LinCombStream(ctx, streams, coeffs)
    assert lc.to_ints_be() == o.lincomb_stream([s1, s2, s3], coeffs)
    lc2 = LinCombStream(ctx, [mt, s2], coeffs[:2])       # nested adaptor + host stream
    assert lc2.to_ints_be() == o.lincomb_stream([want, s2], coeffs[:2])


@pytest.mark.parametrize("nf,ng", [(64, 64), (33, 32), (65, 64), (1 << 12, 1 << 12)])
def test_space_and_elastic_provers_from_device_streams(ctx, nf, ng):
    f_be, g_be = rand_scalars(nf, 90 + nf), rand_scalars(ng, 91 + ng)
    tw = rand_scalars(1, 92)[0]
    ch = rand_scalars(20, 93)
    it = iter(ch)
    want = o.sumcheck_prove(o.SpaceProver(f_be, g_be, tw), lambda m: next(it))
    f_dev = ReverseStream(DeviceFr.from_host(ctx, f_be[::-1]))
    g_dev = ReverseStream(DeviceFr.from_host(ctx, g_be[::-1]))
    for fa, ga in ((f_be, g_be), (f_dev, g_dev), (field.fr_to_limbs(f_be), field.fr_to_limbs(g_be))):
        it = iter(ch)
        got = gm.Sumcheck.prove(gm.SpaceProver(ctx, fa, ga, tw), lambda m: next(it))
        assert got.messages == want[0] and tuple(got.final_foldings[0]) == tuple(want[2])
        for threshold in (22, 3, 0):
            it = iter(ch)
            it2 = iter(ch)
            el = gm.Sumcheck.prove(gm.ElasticProver(ctx, fa, ga, tw, threshold=threshold), lambda m: next(it))
            el_o = o.sumcheck_prove(o.ElasticProver(f_be, g_be, tw, threshold), lambda m: next(it2))
            assert el.messages == el_o[0] and tuple(el.final_foldings[0]) == tuple(el_o[2])


def test_elastic_equals_time_proof_dummy_2_14(ctx):
    """the reference's strongest test (snark/tests.rs:13-58) at a size the oracle cannot reach, on the default dummy
    R1CS of the examples (all-equal scalars): both device compositions must produce the same proof"""
    from gemini_b200 import snark
    from test_gpu_snark import HashTranscript

    n = 1 << 14
    ck = gm.CommitterKey.new(ctx, n + 3, 3, random.Random(3))
    cks = gm.CommitterKeyStream.from_committer_key(ck)
    r1cs = snark.R1cs.dummy(ctx, n, 0x1234567 % R)
    a = snark.new_time(ctx, r1cs, ck, HashTranscript())
    b = snark.new_elastic(ctx, r1cs, cks, HashTranscript(), 1 << 20)
    assert a == b


def test_batch_commit_two_lanes_equals_sequential(ctx, monkeypatch):
    """CommitterKey.batch_commit of resident polynomials runs on two contexts from two host threads (the reduction
    tail of one MSM overlaps the accumulation of the next); order and values are those of the sequential map"""
    srs = rand_points(600, 95)
    ck = gm.CommitterKey(ctx, srs)
    sizes = [600, 1, 300, 0, 17, 512, 2, 129, 64]
    polys = [rand_scalars(n, 96 + n) for n in sizes]
    dev = [DeviceFr.from_host(ctx, p) for p in polys]
    want = [o.kzg_commit(srs, p) for p in polys]
    assert ck.batch_commit(dev) == want
    monkeypatch.setattr(kzg, "CONCURRENT_COMMITS", False)
    assert ck.batch_commit(dev) == want
    assert ck.batch_commit(polys) == want          # host inputs: the plain sequential map
    # against a precomputed key at a size where both lanes carry real work
    ck2 = gm.CommitterKey.new(ctx, (1 << 16) - 1, 3, random.Random(4), precompute=True)
    monkeypatch.setattr(kzg, "CONCURRENT_COMMITS", True)
    f = DeviceFr.random(ctx, 1 << 16, 55)
    levels = f.fold_chain(rand_scalars(12, 97))
    got = ck2.batch_commit(levels)
    monkeypatch.setattr(kzg, "CONCURRENT_COMMITS", False)
    assert got == ck2.batch_commit(levels)
