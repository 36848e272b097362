"""Loader of tests/golden/hotpath_v1.json (written by tools/make_golden.py)."""
import json
import os

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_v1.json")


def _i(v):
    return None if v is None else int(v, 16)


def _p(v):
    return None if v is None else (int(v[0], 16), int(v[1], 16))


def load():
    with open(PATH) as fh:
        doc = json.load(fh)
    msm = [dict(name=c["name"], bases=[_p(b) for b in c["bases"]], scalars=[_i(s) for s in c["scalars"]], result=_p(c["result"]))
           for c in doc["msm"]]
    fold = [dict(f=[_i(v) for v in c["f"]], r=_i(c["r"]), out=[_i(v) for v in c["out"]]) for c in doc["fold"]]

    def sc(c):
        return dict(f=[_i(v) for v in c["f"]], g=[_i(v) for v in c["g"]], twist=_i(c["twist"]),
                    challenges=[_i(v) for v in c["challenges"]], messages=[(_i(a), _i(b)) for a, b in c["messages"]],
                    final_foldings=tuple(_i(v) for v in c["final_foldings"]))

    return dict(msm=msm, fold=fold, sumcheck=[sc(c) for c in doc["sumcheck"]], herring=[sc(c) for c in doc["herring"]])
