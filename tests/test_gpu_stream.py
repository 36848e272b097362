"""Streamed MSM parity: msm_chunks / ChunkedPippenger / HashMapPippenger / KZG committer keys.

Pattern of the reference's tests: time commit == space commit (src/kzg/tests.rs:15-29), time open ==
space open (:42-59), the multi-point KAT (src/kzg/space.rs:321-387)."""
import pytest

import gemini_b200 as gm
import pyref as o
from util import R, rand_points, rand_scalars

pytestmark = pytest.mark.gpu


def test_msm_chunks_equals_commit(ctx):
    n = 700
    srs_le = rand_points(n + 30, 20)
    poly = rand_scalars(n, 21)
    want = o.kzg_commit(srs_le, poly)
    ck = gm.CommitterKey(ctx, srs_le)
    assert ck.commit(poly) == want
    cks = gm.CommitterKeyStream.from_committer_key(ck)
    for step in (1 << 20, 256, 97, 1):
        if step == 1 and n > 64:
            assert cks.commit(poly[::-1][-40:], step=step) == o.kzg_commit(srs_le, poly[:40])
            continue
        assert cks.commit(poly[::-1], step=step) == want
    assert gm.msm_chunks(ctx, srs_le[::-1], poly[::-1], step=128) == o.msm_chunks(srs_le[::-1], poly[::-1], 128) == want


def test_batch_commit_and_truncation(ctx):
    srs = rand_points(64, 22)
    ck = gm.CommitterKey(ctx, srs)
    polys = [rand_scalars(k, 23 + k) for k in (1, 2, 15, 64, 80)]
    assert ck.batch_commit(polys) == [o.kzg_commit(srs, p) for p in polys]


@pytest.mark.parametrize("buf", [0, 1, 7, 50, 1000])
def test_chunked_pippenger(ctx, buf):
    bases, scalars = rand_points(120, 24), rand_scalars(120, 25)
    cp = gm.ChunkedPippenger(ctx, buf)
    ref = o.ChunkedPippenger(buf)
    for b, s in zip(bases, scalars):
        cp.add(b, s)
        ref.add(b, s)
    assert cp.finalize() == ref.finalize() == o.naive_msm(bases, scalars)


@pytest.mark.parametrize("cap", [1, 5, 64])
def test_hashmap_pippenger(ctx, cap):
    pts = rand_points(10, 26)
    bases = [pts[i % 10] for i in range(90)] + [None] * 3
    scalars = rand_scalars(93, 27)
    hp = gm.HashMapPippenger(ctx, cap)
    ref = o.HashMapPippenger(cap)
    for b, s in zip(bases, scalars):
        hp.add(b, s)
        ref.add(b, s)
    assert hp.finalize() == ref.finalize() == o.naive_msm(bases, scalars)


def test_open_time_equals_space(ctx):
    d = 15
    srs = rand_points(d + 5, 28)
    poly = rand_scalars(d + 1, 29)
    alpha = rand_scalars(1, 30)[0]
    ck = gm.CommitterKey(ctx, srs)
    cks = gm.CommitterKeyStream.from_committer_key(ck)
    ev_t, pf_t = ck.open(poly, alpha)
    assert (ev_t, pf_t) == o.kzg_open(srs, poly, alpha)
    ev_s, pf_s = cks.open(poly[::-1], alpha, max_msm_buffer=4)
    assert (ev_s, pf_s) == o.kzg_stream_open(srs[::-1], poly[::-1], alpha, 4)
    assert ev_s == ev_t == o.evaluate_le(poly, alpha) and pf_s == pf_t


def test_open_multi_points_kat(ctx):
    """src/kzg/space.rs:321-387: f = 80x^6+80x^5+88x^4+3x^3+73x^2+7x+24 at {53^2, 53, -53}."""
    poly_be = [80, 80, 88, 3, 73, 7, 24]
    pts = [53 * 53, 53, R - 53]
    srs = rand_points(20, 31)
    ck = gm.CommitterKey(ctx, srs)
    cks = gm.CommitterKeyStream.from_committer_key(ck)
    rem, proof = cks.open_multi_points(poly_be, pts, max_msm_buffer=3)
    assert o.evaluate_be(rem, 53) == 1807299544171
    rem_o, proof_o = o.kzg_stream_open_multi_points(srs[::-1], poly_be, pts, 3)
    assert rem == rem_o and proof == proof_o
    assert ck.open_multi_points(poly_be[::-1], pts) == proof_o == o.kzg_open_multi_points(srs, poly_be[::-1], pts)
    polys = [rand_scalars(9, 32), rand_scalars(12, 33), rand_scalars(5, 34)]
    eta = rand_scalars(1, 35)[0]
    comb = [sum(pow(eta, k, R) * (p[i] if i < len(p) else 0) for k, p in enumerate(polys)) % R for i in range(12)]
    assert ck.batch_open_multi_points(polys, pts, eta) == o.kzg_open_multi_points(srs, comb, pts)


@pytest.mark.parametrize("n,k", [(16, 3), (19, 4), (33, 5)])
def test_commit_folding(ctx, n, k):
    srs = rand_points(n + 3, 36)
    poly_be = rand_scalars(n, 37 + n)
    chals = rand_scalars(k, 38)
    cks = gm.CommitterKeyStream(ctx, srs[::-1])
    assert cks.commit_folding(poly_be, chals, 20) == o.kzg_commit_folding(srs[::-1], poly_be, chals, 20)


def test_index_by(ctx):
    """kzg/time.rs:86-95 (psnark's row/column keys): merged and permuted SRS, identity where unused."""
    srs = rand_points(12, 60)
    idx = [3, 0, 3, 7, 7, 7, 1, 0, 11, 5, 5, 2]
    ck = gm.CommitterKey(ctx, srs).index_by(idx)
    want = o.kzg_index_by(srs, idx)
    assert ck.srs.points() == want
    poly = rand_scalars(12, 61)
    assert ck.commit(poly) == o.kzg_commit(want, poly)
