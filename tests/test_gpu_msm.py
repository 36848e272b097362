"""MSM parity: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Mirrors the reference's differential pattern (src/kzg/msm/variable_base.rs:179-215: Pippenger vs
naive sum of scalar multiples) and the degenerate inputs its default workloads produce.
"""
import numpy as np
import pytest

import gemini_b200 as gm
import pyref as o
from gemini_b200 import field
from util import R, fr_random_limbs, limbs_to_ints, rand_points, rand_scalars

pytestmark = pytest.mark.gpu


def run(ctx, bases, scalars):
    return gm.VariableBaseMSM(ctx).msm_unchecked(bases, scalars)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 7, 31, 32, 33, 100, 257, 1000])
def test_msm_small_vs_naive(ctx, n):
    bases, scalars = rand_points(n, 1), rand_scalars(n, n + 1)
    assert run(ctx, bases, scalars) == o.naive_msm(bases, scalars)


@pytest.mark.parametrize("n", [1 << 12, 5000])
def test_msm_config1_vs_pippenger_oracle(ctx, n):
    """BASELINE config 1 (msm_bench 2^12) - bases resident on the device."""
    bases, scalars = rand_points(n, 2), rand_scalars(n, 3)
    srs = ctx.srs_load(bases)
    got = field.jacobian_to_affine(ctx.msm(srs, scalars))
    assert got == o.msm_unchecked(bases, scalars)
    # raw output is normalised: Z == 1 in Montgomery form
    raw = ctx.msm(srs, scalars)
    assert list(raw[12:]) == field._limbs((1 << 384) % field.Q, 6)


def test_msm_edge_scalars(ctx):
    bases = rand_points(12, 4)
    scalars = [0, 1, R - 1, 1 << 254, R - 2, 2, (R - 1) // 2, (R + 1) // 2, 0, 1 << 128, (1 << 255) % R, 12345]
    assert run(ctx, bases, scalars) == o.naive_msm(bases, scalars)
    assert run(ctx, bases, [0] * 12) is None


def test_msm_identity_and_duplicate_bases(ctx):
    """index_by injects identity points (kzg/time.rs:87); duplicates force P+P in a bucket."""
    pts = rand_points(40, 5)
    bases = pts[:10] + [None, None] + pts[10:20] + [pts[3], pts[3], o.g1_neg(pts[4])] + pts[20:]
    scalars = rand_scalars(len(bases), 6)
    scalars[22] = scalars[3]        # same base, same scalar -> doubling inside a bucket
    scalars[24] = scalars[4]        # P + (-P) inside a bucket
    assert run(ctx, bases, scalars) == o.naive_msm(bases, scalars)
    # arkworks 104-byte records with the infinity flag
    ark = field.g1_to_ark104(bases)
    got = field.jacobian_to_affine(ctx.msm_hostbases(ark, scalars))
    assert got == o.naive_msm(bases, scalars)
    srs = ctx.srs_load(ark)
    assert field.jacobian_to_affine(ctx.msm(srs, scalars)) == o.naive_msm(bases, scalars)


def test_msm_all_equal_scalars_splits_buckets(ctx):
    """dummy_r1cs (src/circuit.rs:349-365) makes every scalar identical: one bucket per window holds all
    points and is cut into many work items + a CTA combine."""
    n = 3000
    bases = rand_points(n, 7)
    s = rand_scalars(1, 8)[0]
    want = o.g1_mul(o.naive_msm(bases, [1] * n), s)
    assert run(ctx, bases, [s] * n) == want


def test_msm_all_identical_bases(ctx):
    """elastic example SRS: DummyStreamer(G1::generator(), n) (examples/snark.rs:62-65)."""
    n = 2000
    scalars = rand_scalars(n, 9)
    srs = ctx.srs_fill(o.G1_GEN, n)
    want = o.g1_mul(o.G1_GEN, sum(scalars) % R)
    assert field.jacobian_to_affine(ctx.msm(srs, scalars)) == want
    assert field.jacobian_to_affine(ctx.msm(srs, [5] * n)) == o.g1_mul(o.G1_GEN, 5 * n)


def test_msm_truncation_and_checked(ctx):
    bases, scalars = rand_points(50, 10), rand_scalars(64, 11)
    srs = ctx.srs_load(bases)
    v = gm.VariableBaseMSM(ctx)
    # msm_unchecked truncates to the shorter input (commit relies on it, kzg/time.rs:82)
    assert v.msm_unchecked(srs, scalars) == o.naive_msm(bases, scalars[:50])
    assert v.msm_unchecked(srs, scalars[:20]) == o.naive_msm(bases[:20], scalars[:20])
    assert v.msm(srs, scalars) == o.msm_checked(bases, scalars) == ("err", 50)
    assert v.msm(bases, scalars[:50]) == ("ok", o.naive_msm(bases, scalars[:50]))
    assert v.msm_bigint(srs, scalars[:50]) == o.naive_msm(bases, scalars[:50])
    assert field.jacobian_to_affine(ctx.msm(srs, scalars[:10], base_offset=30)) == o.naive_msm(bases[30:40], scalars[:10])


def test_srs_generate_and_g1_sum(ctx):
    srs = ctx.srs_generate(100, first_multiple=0)
    pts = srs.points()
    assert pts[0] is None and pts[1] == o.G1_GEN
    assert pts[37] == o.g1_mul(o.G1_GEN, 37) and pts[99] == o.g1_mul(o.G1_GEN, 99)
    srs2 = ctx.srs_generate(33, first_multiple=(1 << 40) + 5)
    assert srs2.points()[32] == o.g1_mul(o.G1_GEN, (1 << 40) + 37)
    parts = [o.g1_mul(o.G1_GEN, k) for k in (3, 5, 9)] + [None]
    jac = np.stack([field.affine_to_jacobian_limbs(p) for p in parts])
    assert field.jacobian_to_affine(ctx.g1_sum(jac)) == o.g1_mul(o.G1_GEN, 17)


@pytest.mark.parametrize("logn", [16, 20])
def test_msm_closed_form_large(ctx, logn):
    """Size-independent property at BASELINE config 2 size: bases P_i = [i+1]G generated on the device,
    so sum_i s_i P_i = [sum_i s_i (i+1) mod r] G, which the oracle evaluates with one scalar mul."""
    n = 1 << logn
    srs = ctx.srs_generate(n, first_multiple=1)
    limbs = fr_random_limbs(n, seed=logn)
    rinv = pow(1 << 256, -1, R)
    tot = sum(v * (i + 1) for i, v in enumerate(limbs_to_ints(limbs))) % R * rinv % R
    assert field.jacobian_to_affine(ctx.msm(srs, limbs)) == o.g1_mul(o.G1_GEN, tot)
    # the device generator agrees with its numpy restatement (same scalars, resident in HBM)
    d = ctx.dev_alloc(n * 32)
    ctx.fr_random_dev(d, n, logn)
    assert np.array_equal(ctx.dev_download(d, n * 32).reshape(n, 4), limbs)
    assert field.jacobian_to_affine(ctx.msm_dev(srs, d, n)) == o.g1_mul(o.G1_GEN, tot)
    ctx.dev_free(d)
    # linearity: msm(2s) == 2 msm(s) through the bigint entry point
    small = limbs_to_ints(limbs[:4096])
    a = field.jacobian_to_affine(ctx.msm(srs, [2 * (v * rinv % R) % R for v in small], bigint=True))
    b = field.jacobian_to_affine(ctx.msm(srs, limbs[:4096]))
    assert a == o.g1_double(b)


# ---- precomputed 2^(c*w) tables (gm_srs_precompute): same results through the merged-window path ----
def test_precompute_matches_plain_path(ctx):
    pts = rand_points(300, 50)
    bases = pts[:100] + [None] + pts[100:200] + [pts[7], pts[7], o.g1_neg(pts[8])] + pts[200:]
    scalars = rand_scalars(len(bases), 51)
    scalars[201] = scalars[7]
    scalars[203] = scalars[8]
    scalars[5:9] = [0, 1, R - 1, 1 << 254]
    want = o.naive_msm(bases, scalars)
    srs = ctx.srs_load(bases)
    assert field.jacobian_to_affine(ctx.msm(srs, scalars)) == want
    srs.precompute()
    c, levels = srs.precompute_info()
    assert levels == (256 + c - 1) // c and c >= 10
    raw = ctx.msm(srs, scalars)
    assert field.jacobian_to_affine(raw) == want
    assert np.array_equal(raw, ctx.msm(ctx.srs_load(bases), scalars))  # byte-identical normalised output
    # prefix, offset and bigint entry points against the table
    assert field.jacobian_to_affine(ctx.msm(srs, scalars[:50])) == o.naive_msm(bases[:50], scalars[:50])
    assert field.jacobian_to_affine(ctx.msm(srs, scalars[:40], base_offset=150)) == o.naive_msm(bases[150:190], scalars[:40])
    assert gm.VariableBaseMSM(ctx).msm_bigint(srs, scalars) == want
    # all-equal scalars: the single hot bucket per window is split into work items
    s = rand_scalars(1, 52)[0]
    assert field.jacobian_to_affine(ctx.msm(srs, [s] * len(bases))) == o.g1_mul(o.naive_msm(bases, [1] * len(bases)), s)
    # the cost model picks c = 20 (13 table levels) for a 2^20-point key
    assert ctx.srs_generate(1 << 20).precompute().precompute_info() == (20, 13)


@pytest.mark.parametrize("logn", [14, 20])
def test_precompute_closed_form_large(ctx, logn):
    n = 1 << logn
    srs = ctx.srs_generate(n, first_multiple=1).precompute()
    limbs = fr_random_limbs(n, seed=100 + logn)
    rinv = pow(1 << 256, -1, R)
    tot = sum(v * (i + 1) for i, v in enumerate(limbs_to_ints(limbs))) % R * rinv % R
    assert field.jacobian_to_affine(ctx.msm(srs, limbs)) == o.g1_mul(o.G1_GEN, tot)
    # short commitments against the long key use the nested short-prefix tables (tensorcheck fold levels)
    for m in (1, 100, n >> 7, (n >> 3) + 1):
        tot_m = sum(v * (i + 1) for i, v in enumerate(limbs_to_ints(limbs[:m]))) % R * rinv % R
        assert field.jacobian_to_affine(ctx.msm(srs, limbs[:m])) == o.g1_mul(o.G1_GEN, tot_m)
    # streamed chunks against the same table (msm_chunks, chunk 2^12) give the same point
    st = gm.msm._DeviceStream(ctx, srs, 1 << 12)
    for s0 in range(0, min(n, 1 << 15), 1 << 12):
        st.push_range(s0, limbs[s0:s0 + (1 << 12)])
    m = min(n, 1 << 15)
    tot_m = sum(v * (i + 1) for i, v in enumerate(limbs_to_ints(limbs[:m]))) % R * rinv % R
    assert st.finalize() == o.g1_mul(o.G1_GEN, tot_m)


def test_committer_key_new_fixed_base_setup(ctx):
    """CommitterKey::new (kzg/time.rs:49-72): powers_of_g[i] = tau^i g by the device fixed-base MSM."""
    import random

    rng = random.Random(31)
    g = o.g1_mul(o.G1_GEN, rng.randrange(1, R))
    for tau in (rng.randrange(1, R), 1, R - 1, 2):
        n = 37
        pts = ctx.srs_setup(g, tau, n).points()
        assert pts == [o.g1_mul(g, pow(tau, i, R)) for i in range(n)]
    assert ctx.srs_setup(g, 0, 3).points() == [g, None, None]
    # size-independent property at 2^18: commit(f) against the generated key is [f(tau)] g
    ck = gm.CommitterKey.new(ctx, (1 << 18) - 1, 3, random.Random(5))
    assert ck.max_degree() == (1 << 18) - 1
    limbs = fr_random_limbs(1 << 18, seed=9)
    rinv = pow(1 << 256, -1, R)
    acc = 0
    for v in reversed(limbs_to_ints(limbs)):
        acc = (acc * ck.tau + v * rinv) % R
    assert field.jacobian_to_affine(ctx.msm(ck.srs, limbs)) == o.g1_mul(ck.g, acc)
    sample = ck.srs.points(12345, 2)
    assert sample == [o.g1_mul(ck.g, pow(ck.tau, 12345, R)), o.g1_mul(ck.g, pow(ck.tau, 12346, R))]
