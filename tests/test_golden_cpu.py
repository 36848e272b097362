"""The oracles (big-integer Pippenger restatement, C restatement) against the frozen vectors of tests/golden/."""
import numpy as np
import pytest

import golden_util
import pyref as o
from test_oracle_c import c_msm, clib, fr_i, fr_l  # noqa: F401  (clib is a fixture)

G = golden_util.load()


@pytest.mark.parametrize("case", G["msm"], ids=lambda c: c["name"])
def test_golden_msm_oracles(clib, case):
    n = min(len(case["bases"]), len(case["scalars"]))
    assert o.msm_unchecked(case["bases"], case["scalars"]) == case["result"]
    assert o.pippenger_msm(case["bases"][:n], case["scalars"][:n]) == case["result"]
    assert c_msm(clib, case["bases"], case["scalars"])[0] == case["result"]
    for p in case["bases"]:
        assert o.g1_is_on_curve(p)


@pytest.mark.parametrize("case", G["fold"], ids=lambda c: str(len(c["f"])))
def test_golden_fold_oracles(clib, case):
    assert o.fold_polynomial(case["f"], case["r"]) == case["out"]
    fa, ra = fr_l(case["f"]), fr_l([case["r"]])
    out = np.zeros((len(case["out"]), 4), dtype=np.uint64)
    clib.go_fr_fold(fa.ctypes.data, len(case["f"]), ra.ctypes.data, out.ctypes.data)
    assert fr_i(out) == case["out"]


@pytest.mark.parametrize("kind", ["sumcheck", "herring"])
def test_golden_sumcheck_oracle(kind):
    cls = o.TimeProver if kind == "sumcheck" else o.HerringTimeProver
    for case in G[kind]:
        it = iter(case["challenges"] + [0])
        msgs, used, ff = o.sumcheck_prove(cls(case["f"], case["g"], case["twist"]), lambda m: next(it))
        assert msgs == case["messages"] and used == case["challenges"] and tuple(ff) == case["final_foldings"]


def test_golden_sumcheck_c(clib):
    for case in G["sumcheck"]:
        nf, ng, k = len(case["f"]), len(case["g"]), len(case["challenges"])
        fa, ga, ta, ca = fr_l(case["f"]), fr_l(case["g"]), fr_l([case["twist"]]), fr_l(case["challenges"] + [0])
        out = np.zeros((k + 1, 8), dtype=np.uint64)
        fin = np.zeros(8, dtype=np.uint64)
        rounds = clib.go_sumcheck_time(fa.ctypes.data, nf, ga.ctypes.data, ng, ta.ctypes.data, ca.ctypes.data, k + 1,
                                       out.ctypes.data, fin.ctypes.data)
        got = fr_i(out[:rounds])
        assert [(got[2 * i], got[2 * i + 1]) for i in range(rounds)] == case["messages"]
        assert tuple(fr_i(fin)) == case["final_foldings"]
