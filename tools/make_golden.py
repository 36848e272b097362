"""Freeze golden vectors for the hot path under tests/golden/.

The reference is Rust and cannot be run in this image (no cargo/rustc; ark-ec / ark-ff are not vendored), so the
vectors are produced by the oracle of record of DESIGN.md section 2: the big-integer restatement oracle/pyref.py,
with every MSM value computed by NAIVE double-and-add (not by Pippenger).  They pin the C restatement, the Pippenger
restatement and the CUDA path to one frozen answer that no later edit of the oracle can silently move.

    python tools/make_golden.py            # rewrites tests/golden/hotpath_v1.json
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref as o  # noqa: E402

R = o.R


def hx(v):
    return None if v is None else hex(v)


def pt(p):
    return None if p is None else [hex(p[0]), hex(p[1])]


def chain(rng, n):
    p = o.g1_mul(o.G1_GEN, rng.randrange(1, R))
    d = o.g1_mul(o.G1_GEN, rng.randrange(1, R))
    out = []
    for _ in range(n):
        out.append(p)
        p = o.g1_add(p, d)
    return out


def main():
    rng = random.Random(0x47454D494E49)
    doc = {"curve": "BLS12-381 G1", "generator": "tools/make_golden.py", "msm": [], "fold": [], "sumcheck": [], "herring": []}

    # --- MSM: uniform, edge scalars, identity / duplicate / negated bases, all-equal scalars, all-equal bases ---
    b = chain(rng, 48)
    cases = [("uniform", b, [rng.randrange(R) for _ in b])]
    edge = [0, 1, R - 1, 1 << 254, R - 2, 2, (R - 1) // 2, (R + 1) // 2, 0, 1 << 128, (1 << 255) % R, 12345]
    cases.append(("edge_scalars", b[:12], edge))
    mixed = b[:10] + [None, None] + b[10:20] + [b[3], b[3], o.g1_neg(b[4])]
    cases.append(("identity_dup_neg_bases", mixed, [rng.randrange(R) for _ in mixed]))
    s = rng.randrange(R)
    cases.append(("all_equal_scalars", b, [s] * len(b)))
    cases.append(("all_equal_bases", [b[7]] * 40, [rng.randrange(R) for _ in range(40)]))
    cases.append(("all_equal_both", [b[9]] * 33, [s] * 33))
    cases.append(("cancel_to_identity", [b[0], o.g1_neg(b[0]), b[1], b[1]], [5, 5, R - 3, 3]))
    cases.append(("scalars_longer_than_bases", b[:5], [rng.randrange(R) for _ in range(9)]))
    cases.append(("small_sparse", b, [rng.choice([0, 0, 1, 1, rng.randrange(1 << 16)]) for _ in b]))
    for name, bases, scalars in cases:
        doc["msm"].append({"name": name, "bases": [pt(p) for p in bases], "scalars": [hx(v) for v in scalars],
                           "result": pt(o.naive_msm(bases, scalars))})

    # --- fold_polynomial (misc.rs:52-56) ---
    for n in (1, 2, 7, 16, 33):
        f = [rng.randrange(R) for _ in range(n)]
        r = rng.randrange(R)
        doc["fold"].append({"f": [hx(v) for v in f], "r": hx(r), "out": [hx(v) for v in o.fold_polynomial(f, r)]})

    # --- TimeProver transcripts (time_prover.rs:83-137) with fixed challenges ---
    for nf, ng, tw in ((16, 16, 1), (17, 17, None), (29, 8, None), (5, 40, 1), (1, 1, None)):
        f = [rng.randrange(R) for _ in range(nf)]
        g = [rng.randrange(R) for _ in range(ng)]
        twist = 1 if tw == 1 else rng.randrange(R)
        chal = [rng.randrange(R) for _ in range(8)]
        for kind, cls in (("sumcheck", o.TimeProver), ("herring", o.HerringTimeProver)):
            it = iter(chal)
            msgs, used, ff = o.sumcheck_prove(cls(f, g, twist), lambda m: next(it))
            doc[kind].append({"f": [hx(v) for v in f], "g": [hx(v) for v in g], "twist": hx(twist),
                              "challenges": [hx(v) for v in used], "messages": [[hx(a), hx(b)] for a, b in msgs],
                              "final_foldings": [hx(ff[0]), hx(ff[1])]})

    path = os.path.join(ROOT, "tests", "golden", "hotpath_v1.json")
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=0, sort_keys=True)
        fh.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
