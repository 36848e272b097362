"""Sumcheck prover parity (device vs oracle), mirroring src/subprotocols/sumcheck/tests.rs."""
import random

import pytest

import gemini_b200 as gm
import pyref as o
from util import R, rand_scalars

pytestmark = pytest.mark.gpu


def challenge_fn_factory(seed):
    rng = random.Random(seed)
    return lambda msg: rng.randrange(R)


def drive(prover, seed):
    sc = o.sumcheck_prove(prover, challenge_fn_factory(seed))
    return sc


def drive_dev(prover, seed):
    sc = gm.Sumcheck.prove(prover, challenge_fn_factory(seed))
    return sc.messages, sc.challenges, sc.final_foldings[0]


CASES = [(1, 1), (2, 2), (16, 16), (30, 30), (17, 17), (93, 16), (16, 93), (1025, 1025), (5, 1000), (4096, 4096), (9000, 8191)]


@pytest.mark.parametrize("nf,ng", CASES)
@pytest.mark.parametrize("twist_one", [True, False])
def test_time_prover_messages(ctx, nf, ng, twist_one):
    f, g = rand_scalars(nf, nf), rand_scalars(ng, ng + 7)
    twist = 1 if twist_one else rand_scalars(1, 3)[0]
    want = drive(o.TimeProver(f, g, twist), 42)
    got = drive_dev(gm.TimeProver(ctx, f, g, twist), 42)
    assert got[0] == want[0]
    assert got[1] == want[1]
    assert got[2] == want[2]
    # completeness: messages satisfy the verifier recurrence (subclaim.rs:77-97)
    if nf == ng:
        asserted = sum(a * b * pow(twist, i, R) for i, (a, b) in enumerate(zip(f, g))) % R
        ff, gf = got[2]
        assert o.subclaim_reduce(got[0], got[1], asserted) == ff * gf % R


def test_time_prover_step_by_step_state(ctx):
    """fold() / next_message() interleaving and the folded state itself."""
    f, g, tw = rand_scalars(29, 1), rand_scalars(29, 2), rand_scalars(1, 3)[0]
    ref, dev = o.TimeProver(f, g, tw), gm.TimeProver(ctx, f, g, tw)
    assert dev.rounds() == ref.tot_rounds == 5
    assert dev.next_message(None) == ref.next_message(None)
    for r in rand_scalars(3, 4):
        assert dev.next_message(r) == ref.next_message(r)
        sf, sg, stw = dev.state()
        assert (sf, sg, stw) == (ref.f, ref.g, ref.twist)
    ref.fold(12345)
    dev.fold(12345)
    sf, sg, stw = dev.state()
    assert (sf, sg, stw) == (ref.f, ref.g, ref.twist)
    assert dev.final_foldings() is None and ref.final_foldings() is None
    assert dev.round() == ref.round


def test_next_message_past_last_round_is_an_error(ctx):
    """time_prover.rs:84 asserts round <= tot_rounds; the C ABI reports GM_ERR_STATE instead of unwinding."""
    from gemini_b200._lib import lib

    dev = gm.TimeProver(ctx, [1, 2], [3, 4], 1)
    assert dev.next_message(None) is not None
    assert dev.next_message(7) is None
    ref = o.TimeProver([1, 2], [3, 4], 1)
    ref.next_message(None)
    ref.next_message(7)
    assert dev.final_foldings() == ref.final_foldings()
    lib.gm_sumcheck_set_rounds(dev._h, 5, 2)
    with pytest.raises(gm.GeminiError):
        dev.next_message(None)


@pytest.mark.parametrize("nf,ng", [(16, 16), (31, 31), (64, 20), (20, 64), (1000, 1000)])
def test_herring_prover(ctx, nf, ng):
    f, g = rand_scalars(nf, nf + 1), rand_scalars(ng, ng + 2)
    for twist in (1, rand_scalars(1, 5)[0]):
        want = drive(o.HerringTimeProver(f, g, twist), 9)
        got = drive_dev(gm.HerringTimeProver(ctx, f, g, twist), 9)
        assert got == (want[0], want[1], want[2])


@pytest.mark.parametrize("nf,ng", [(16, 16), (30, 30), (29, 29), (93, 16), (16, 93), (1024, 1024)])
def test_space_and_elastic_prover(ctx, nf, ng):
    """Time vs Space vs Elastic on big-endian streams (sumcheck/tests.rs:41-138)."""
    f, g = rand_scalars(nf, nf + 3), rand_scalars(ng, ng + 4)
    twist = rand_scalars(1, 6)[0]
    want = drive(o.SpaceProver(f[::-1], g[::-1], twist), 13)
    got = drive_dev(gm.SpaceProver(ctx, f[::-1], g[::-1], twist), 13)
    assert got == (want[0], want[1], want[2])
    got_e = drive_dev(gm.ElasticProver(ctx, f[::-1], g[::-1], twist), 13)
    want_e = drive(o.ElasticProver(f[::-1], g[::-1], twist), 13)
    assert got_e == (want_e[0], want_e[1], want_e[2])
    if nf == ng:
        t = drive(o.TimeProver(f, g, twist), 13)
        assert got[0] == t[0]


def test_elastic_explicit_fold_switches_to_time(ctx):
    f, g, tw = rand_scalars(64, 1), rand_scalars(64, 2), rand_scalars(1, 3)[0]
    ref, dev = o.ElasticProver(f[::-1], g[::-1], tw, threshold=4), gm.ElasticProver(ctx, f[::-1], g[::-1], tw, threshold=4)
    assert dev.next_message(None) == ref.next_message(None)
    for r in rand_scalars(6, 4):
        ref.fold(r)
        dev.fold(r)
        assert dev.is_space == ref.is_space
        assert dev.next_message(None) == ref.next_message(None)
    assert dev.final_foldings() == ref.final_foldings()


def test_prove_batch(ctx):
    """sumcheck/tests.rs:226-269: several provers of different sizes under one random linear combination."""
    specs = [(64, 64, 5), (16, 16, 1), (100, 100, 7), (1, 1, 1)]
    mk_ref, mk_dev = [], []
    for nf, ng, tw in specs:
        f, g = rand_scalars(nf, nf + 11), rand_scalars(ng, ng + 12)
        mk_ref.append(o.TimeProver(f, g, tw))
        mk_dev.append(gm.TimeProver(ctx, f, g, tw))
    coeffs = rand_scalars(len(specs), 77)
    c1, c2 = iter(coeffs), iter(coeffs)
    want = o.sumcheck_prove_batch(mk_ref, lambda: next(c1), challenge_fn_factory(5))
    got = gm.Sumcheck.prove_batch(mk_dev, lambda: next(c2), challenge_fn_factory(5))
    assert (got.messages, got.challenges, got.final_foldings) == want
    assert got.rounds == 8


def test_native_prove_loop_equals_python_loop(ctx):
    """gm_sumcheck_prove (Sumcheck::prove, proof.rs:36-66, as one native call with the native Merlin transcript) against
    the same loop driven round by round from Python over the oracle's pure-Python Merlin."""
    from gemini_b200.transcript import MerlinTranscript

    for nf, ng, tw in ((1000, 1000, 7), (93, 16, 1), (4097, 4096, 12345)):
        f, g = rand_scalars(nf, 300 + nf), rand_scalars(ng, 301 + ng)
        t_native, t_oracle = MerlinTranscript(), o.MerlinTranscript()
        t_native.append_serializable(b"zc(alpha)", 99)
        t_oracle.append_serializable(b"zc(alpha)", 99)
        got = gm.Sumcheck.prove_transcript(gm.TimeProver(ctx, f, g, tw), t_native)
        want = o.sumcheck_prove_transcript(o.TimeProver(f, g, tw), t_oracle)
        assert got.messages == want["messages"] and got.challenges == want["challenges"]
        assert [tuple(x) for x in got.final_foldings] == [tuple(x) for x in want["final_foldings"]]
        # the transcripts are in the same state afterwards (final foldings were appended on both sides)
        assert t_native.get_challenge(b"eta") == t_oracle.get_challenge(b"eta")


@pytest.mark.parametrize("nf,ng,tw", [(1 << 15, 1 << 15, 3), ((1 << 14) + 1, 1 << 14, 1), (5000, 4097, 77), (8192, 8192, 1), (64, 64, 9), (2, 2, 5), (1, 1, 5)])
def test_native_prove_loop_sizes_and_flavours(ctx, nf, ng, tw):
    """gm_sumcheck_prove (every round's message arrives through the prover's pinned mailbox) against the oracle driven by
    the oracle's own Merlin, for ragged lengths, twist = 1 and degenerate sizes; the herring flavour against its
    round-by-round Python-driven twin."""
    from gemini_b200.transcript import MerlinTranscript

    f, g = rand_scalars(nf, 400 + nf), rand_scalars(ng, 401 + ng)
    want = o.sumcheck_prove_transcript(o.TimeProver(f, g, tw), o.MerlinTranscript())
    got = gm.Sumcheck.prove_transcript(gm.TimeProver(ctx, f, g, tw), MerlinTranscript())
    assert got.messages == want["messages"] and got.challenges == want["challenges"]
    assert [tuple(x) for x in got.final_foldings] == [tuple(x) for x in want["final_foldings"]]
    if min(nf, ng) >= 2:
        h = gm.Sumcheck.prove_transcript(gm.HerringTimeProver(ctx, f, g, tw), MerlinTranscript())
        it = iter(h.challenges)
        h0 = gm.Sumcheck.prove(gm.HerringTimeProver(ctx, f, g, tw), lambda m: next(it))
        assert h.messages == h0.messages and [tuple(x) for x in h.final_foldings] == [tuple(x) for x in h0.final_foldings]
