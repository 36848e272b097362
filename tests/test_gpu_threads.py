"""Distinct sumcheck handles driven concurrently from different host threads - the `prove_batch` pattern of the
reference (provers are `Send + Sync` and driven from rayon workers, /root/reference/src/subprotocols/sumcheck/proof.rs:85).
Each handle owns its stream, events, vectors and pinned message slot (include/gemini_b200.h "threads"); ctypes releases
the GIL during the calls, so the four provers really overlap.  Messages are checked against the oracle."""
import random
import threading

import pytest

import pyref as o
from util import R


@pytest.mark.gpu
def test_four_provers_from_four_threads(ctx):
    import gemini_b200 as gm

    rng = random.Random(99)
    cases = []
    for k, n in enumerate((4096 + 3, 2500, 8192, 1777)):
        f = [rng.randrange(R) for _ in range(n)]
        g = [rng.randrange(R) for _ in range(n - 5 * k)]
        tw = 1 if k == 2 else rng.randrange(R)
        ch = [rng.randrange(R) for _ in range(16)]
        cases.append((f, g, tw, ch))
    want = []
    for f, g, tw, ch in cases:
        it = iter(ch)
        want.append(o.sumcheck_prove(o.TimeProver(f, g, tw), lambda m: next(it)))
    got = [None] * len(cases)
    errors = []
    start = threading.Barrier(len(cases))

    def work(i):
        try:
            f, g, tw, ch = cases[i]
            start.wait()
            for _ in range(3):   # several provers per thread, created and freed while the others run
                p = gm.TimeProver(ctx, f, g, tw)
                it = iter(ch)
                got[i] = gm.Sumcheck.prove(p, lambda m: next(it))
                p.free()
        except Exception as exc:  # pragma: no cover
            errors.append((i, repr(exc)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(len(cases)):
        assert got[i].messages == want[i][0], f"prover {i}: messages differ"
        assert tuple(got[i].final_foldings[0]) == tuple(want[i][2]), f"prover {i}: final foldings differ"


@pytest.mark.gpu
def test_prover_threads_next_to_msm_thread(ctx):
    """context-level calls (MSM) are serialised by the context lock; provers on their own streams keep running"""
    import gemini_b200 as gm
    from util import rand_points, rand_scalars

    pts, sc = rand_points(300, 5), rand_scalars(300, 6)
    want_msm = o.naive_msm(pts, sc)
    ck = gm.CommitterKey(ctx, pts)
    f, g = rand_scalars(3000, 7), rand_scalars(3000, 8)
    ch = rand_scalars(16, 9)
    it0 = iter(ch)
    want_sc = o.sumcheck_prove(o.TimeProver(f, g, 5), lambda m: next(it0))
    res, errors = {}, []

    def msm_loop():
        try:
            for _ in range(5):
                assert ck.commit(sc) == want_msm
        except Exception as exc:  # pragma: no cover
            errors.append(repr(exc))

    def sc_loop(tag):
        try:
            for _ in range(5):
                it = iter(ch)
                res[tag] = gm.Sumcheck.prove(gm.TimeProver(ctx, f, g, 5), lambda m: next(it))
        except Exception as exc:  # pragma: no cover
            errors.append(repr(exc))

    ts = [threading.Thread(target=msm_loop), threading.Thread(target=msm_loop), threading.Thread(target=sc_loop, args=("a",)),
          threading.Thread(target=sc_loop, args=("b",))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    for tag in ("a", "b"):
        assert res[tag].messages == want_sc[0]
