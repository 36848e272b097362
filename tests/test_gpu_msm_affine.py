"""The affine pre-reduction levels of the bucket accumulation (msm.cu 3b, g1_affine.cuh): pairwise bucket sums with
shared inversions must give the same group element as the XYZZ-only path on every input shape the MSM tests cover.
GM_MSM_AFFINE forces the number of levels (small inputs would not use them on their own)."""
import numpy as np
import pytest

import gemini_b200 as gm
import golden_util
import pyref as o
from gemini_b200 import field
from util import R, fr_random_limbs, limbs_to_ints, rand_points, rand_scalars

pytestmark = pytest.mark.gpu
G = golden_util.load()


@pytest.fixture(params=[1, 2, 4, 5])
def levels(request, monkeypatch):
    monkeypatch.setenv("GM_MSM_AFFINE", str(request.param))
    return request.param


def run(ctx, bases, scalars):
    return gm.VariableBaseMSM(ctx).msm_unchecked(bases, scalars)


@pytest.mark.parametrize("n", [1, 2, 3, 33, 257, 1000])
def test_affine_small_vs_naive(ctx, levels, n):
    bases, scalars = rand_points(n, 1), rand_scalars(n, n + 1)
    assert run(ctx, bases, scalars) == o.naive_msm(bases, scalars)


def test_affine_golden(ctx, levels):
    for case in G["msm"]:
        assert run(ctx, case["bases"], case["scalars"]) == case["result"], case["name"]


def test_affine_degenerate_inputs(ctx, levels):
    pts = rand_points(40, 5)
    bases = pts[:10] + [None, None] + pts[10:20] + [pts[3], pts[3], o.g1_neg(pts[4])] + pts[20:]
    scalars = rand_scalars(len(bases), 6)
    scalars[22] = scalars[3]        # P + P inside a bucket
    scalars[24] = scalars[4]        # P + (-P) inside a bucket
    assert run(ctx, bases, scalars) == o.naive_msm(bases, scalars)
    # all-equal scalars: one bucket per window holds every point (src/circuit.rs:349-365)
    n = 3000
    bases = rand_points(n, 7)
    s = rand_scalars(1, 8)[0]
    assert run(ctx, bases, [s] * n) == o.g1_mul(o.naive_msm(bases, [1] * n), s)
    # all-identical bases (examples/snark.rs:62-65): every pair of a bucket is a doubling
    n = 2000
    scalars = rand_scalars(n, 9)
    srs = ctx.srs_fill(o.G1_GEN, n)
    assert field.jacobian_to_affine(ctx.msm(srs, scalars)) == o.g1_mul(o.G1_GEN, sum(scalars) % R)
    assert field.jacobian_to_affine(ctx.msm(srs, [5] * n)) == o.g1_mul(o.G1_GEN, 5 * n)
    # P, -P alternating with equal scalars: whole buckets cancel to the identity
    pts = rand_points(64, 11)
    bases = [q for p in pts for q in (p, o.g1_neg(p))]
    assert run(ctx, bases, [s] * len(bases)) is None


def test_affine_precomputed_table_and_stream(ctx, levels):
    n = 1 << 14
    srs = ctx.srs_generate(n, first_multiple=1).precompute()
    limbs = fr_random_limbs(n, seed=300 + levels)
    rinv = pow(1 << 256, -1, R)
    vals = limbs_to_ints(limbs)
    tot = sum(v * (i + 1) for i, v in enumerate(vals)) % R * rinv % R
    raw = ctx.msm(srs, limbs)
    assert field.jacobian_to_affine(raw) == o.g1_mul(o.G1_GEN, tot)
    st = gm.msm._DeviceStream(ctx, srs, 1 << 12)
    for s0 in range(0, n, 1 << 12):
        st.push_range(s0, limbs[s0:s0 + (1 << 12)])
    assert st.finalize() == o.g1_mul(o.G1_GEN, tot)


def test_affine_levels_do_not_change_the_bytes(ctx, monkeypatch):
    """The normalised 144-byte output is the same with 0, 3 and 8 levels (plain bases, 2^16 points)."""
    n = 1 << 16
    srs = ctx.srs_generate(n, first_multiple=3)
    limbs = fr_random_limbs(n, seed=77)
    outs = []
    for lv in (0, 3, 8):
        monkeypatch.setenv("GM_MSM_AFFINE", str(lv))
        outs.append(ctx.msm(srs, limbs))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    rinv = pow(1 << 256, -1, R)
    tot = sum(v * (i + 3) for i, v in enumerate(limbs_to_ints(limbs))) % R * rinv % R
    assert field.jacobian_to_affine(outs[0]) == o.g1_mul(o.G1_GEN, tot)
