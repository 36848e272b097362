"""The CUDA path (through the C ABI) against the frozen vectors of tests/golden/."""
import pytest

import gemini_b200 as gm
import golden_util

pytestmark = pytest.mark.gpu
G = golden_util.load()


@pytest.mark.parametrize("case", G["msm"], ids=lambda c: c["name"])
def test_golden_msm_device(ctx, case):
    msm = gm.VariableBaseMSM(ctx)
    assert msm.msm_unchecked(case["bases"], case["scalars"]) == case["result"]
    n = min(len(case["bases"]), len(case["scalars"]))
    srs = ctx.srs_load(case["bases"][:n])
    assert gm.field.jacobian_to_affine(ctx.msm(srs, case["scalars"][:n])) == case["result"]


def test_golden_fold_device(ctx):
    for case in G["fold"]:
        assert gm.fold_polynomial(ctx, case["f"], case["r"]) == case["out"]


@pytest.mark.parametrize("kind", ["sumcheck", "herring"])
def test_golden_sumcheck_device(ctx, kind):
    cls = gm.TimeProver if kind == "sumcheck" else gm.HerringTimeProver
    for case in G[kind]:
        it = iter(case["challenges"] + [0])
        sc = gm.Sumcheck.prove(cls(ctx, case["f"], case["g"], case["twist"]), lambda m: next(it))
        assert sc.messages == case["messages"] and sc.challenges == case["challenges"]
        assert tuple(sc.final_foldings[0]) == case["final_foldings"]
