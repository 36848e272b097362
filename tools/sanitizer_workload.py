#!/usr/bin/env python
"""A small end-to-end pass over every kernel family, checked against the oracle, sized to finish under
compute-sanitizer (memcheck / racecheck / synccheck slow kernels down 10-100x):

    compute-sanitizer --tool memcheck  python tools/sanitizer_workload.py
    compute-sanitizer --tool racecheck python tools/sanitizer_workload.py

MSM with and without the precomputed table and with forced affine levels (all-equal scalars, identity and duplicate
bases included), streamed MSM, folds, TimeProver / SpaceProver from two host threads, the Fr vector helpers, the
device-side KZG opening, and a context shut down before its handles are freed."""
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]


def main():
    import gemini_b200 as gm
    import pyref as o
    from gemini_b200.devvec import DeviceFr
    from gemini_b200.streams import ReverseStream
    from util import rand_points, rand_scalars

    ctx = gm.Context(0)
    n = 200
    pts = rand_points(n, 1)
    bases = pts[:50] + [None] + pts[50:120] + [pts[3], pts[3], o.g1_neg(pts[4])] + pts[120:]
    sc = rand_scalars(len(bases), 2)
    sc[121], sc[123] = sc[3], sc[4]
    want = o.naive_msm(bases, sc)
    for levels in ("0", "2"):
        os.environ["GM_MSM_AFFINE"] = levels
        srs = ctx.srs_load(bases)
        assert gm.field.jacobian_to_affine(ctx.msm(srs, sc)) == want
        assert gm.field.jacobian_to_affine(ctx.msm(srs, [sc[0]] * len(bases))) == o.g1_mul(o.naive_msm(bases, [1] * len(bases)), sc[0])
        srs.precompute()
        assert gm.field.jacobian_to_affine(ctx.msm(srs, sc)) == want
        st = gm.msm._DeviceStream(ctx, srs, 64)
        for s0 in range(0, len(bases), 64):
            st.push_range(s0, sc[s0:s0 + 64])
        assert st.finalize() == want
        st.free()
        srs.free()
    del os.environ["GM_MSM_AFFINE"]
    f, g = rand_scalars(1000, 3), rand_scalars(777, 4)
    r = rand_scalars(1, 5)[0]
    assert gm.fold_polynomial(ctx, f, r) == o.fold_polynomial(f, r)
    ch = rand_scalars(16, 6)
    res, errs = {}, []

    def prove(tag, make, make_o):
        try:
            it, it2 = iter(ch), iter(ch)
            got = gm.Sumcheck.prove(make(), lambda m: next(it))
            ref = o.sumcheck_prove(make_o(), lambda m: next(it2))
            res[tag] = got.messages == ref[0] and tuple(got.final_foldings[0]) == tuple(ref[2])
        except Exception as exc:  # pragma: no cover
            errs.append(repr(exc))

    ts = [threading.Thread(target=prove, args=("time", lambda: gm.TimeProver(ctx, f, g, r), lambda: o.TimeProver(f, g, r))),
          threading.Thread(target=prove, args=("space", lambda: gm.SpaceProver(ctx, f[:65], g[:64], 1), lambda: o.SpaceProver(f[:65], g[:64], 1)))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs and res == {"time": True, "space": True}, (errs, res)
    srs_pts = rand_points(40, 7)
    ck = gm.CommitterKey(ctx, srs_pts)
    cks = gm.CommitterKeyStream(ctx, srs_pts[::-1])
    poly = rand_scalars(33, 8)
    dv = DeviceFr.from_host(ctx, poly)
    assert ck.open(dv, r) == o.kzg_open(srs_pts, poly, r) == cks.open(ReverseStream(dv), r, 4)
    pts3 = rand_scalars(3, 9)
    assert ck.open_multi_points(dv, pts3) == o.kzg_open_multi_points(srs_pts, poly, pts3)
    assert cks.commit_folding(ReverseStream(dv), ch[:4], 20) == o.kzg_commit_folding(srs_pts[::-1], poly[::-1], ch[:4], 20)
    prover = gm.TimeProver(ctx, f[:10], g[:10], 1)
    ctx.close()            # handles outlive the context
    prover.free()
    ck.srs.free()
    cks.srs_be.free()
    print("sanitizer workload ok")


if __name__ == "__main__":
    main()
