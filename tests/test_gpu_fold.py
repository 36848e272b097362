"""Fr fold parity: fold_polynomial / foldings_polynomial / fold chain vs the oracle, incl. the reference KATs."""
import numpy as np
import pytest

import gemini_b200 as gm
import pyref as o
from gemini_b200 import field
from util import R, fr_random_limbs, limbs_to_ints, rand_scalars

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2, 3, 16, 19, 255, 256, 1025, 40001])
def test_fold_polynomial(ctx, n):
    f, r = rand_scalars(n, n), rand_scalars(1, 99)[0]
    assert gm.fold_polynomial(ctx, f, r) == o.fold_polynomial(f, r)


def test_fold_kats(ctx):
    # tensorcheck/mod.rs:388-398
    assert gm.foldings_polynomial(ctx, [100, 101, 102, 103], [1, 1]) == [[201, 205]]
    # sumcheck/streams.rs:232-288: big-endian [1,2,1,1] folded by [1,2] -> 8
    f_le = [1, 2, 1, 1][::-1]
    lv = ctx.fr_fold_chain(f_le, [1, 2])
    assert field.fr_from_limbs(lv[-1]) == [8]
    lv = ctx.fr_fold_chain([1] * 12, [1, 1, 1, 1])
    assert field.fr_from_limbs(lv[-1]) == [12]


@pytest.mark.parametrize("n,k", [(16, 4), (19, 3), (1000, 10), (4097, 13)])
def test_foldings_polynomial(ctx, n, k):
    f, ch = rand_scalars(n, n + 1), rand_scalars(k, 5)
    assert gm.foldings_polynomial(ctx, f, ch) == o.foldings_polynomial(f, ch)
    # the streaming tree (FoldedPolynomialTree, big-endian) yields the same levels
    tree = {}
    for lvl, c in o.folded_polynomial_tree(f[::-1], ch[:-1]):
        tree.setdefault(lvl, []).append(c)
    got = gm.foldings_polynomial(ctx, f, ch)
    for lvl, coeffs in tree.items():
        want = coeffs[::-1]
        assert got[lvl - 1] == want[:len(got[lvl - 1])] and not any(want[len(got[lvl - 1]):])


def test_fold_large_property(ctx):
    """2^22 elements (device-resident): fold is linear in f and agrees with the oracle on a sampled slice."""
    n = 1 << 22
    limbs = fr_random_limbs(n, 77)
    r = rand_scalars(1, 78)[0]
    out = ctx.fr_fold(limbs, r)
    rinv = pow(1 << 256, -1, R)
    idx = [0, 1, 2, 12345, n // 2 - 1, n // 4]
    f = limbs_to_ints(limbs)
    for i in idx:
        want = (f[2 * i] * rinv + r * f[2 * i + 1] * rinv) % R
        assert field.fr_from_limbs(out[i]) == [want]
    # evaluation identity: sum_i out[i] x^i == f_even(x) + r f_odd(x) at x = 1
    tot = sum(limbs_to_ints(out)) * rinv % R
    assert tot == (sum(f[0::2]) + r * sum(f[1::2])) * rinv % R


def test_folded_polynomial_tree_consumers(ctx):
    """Elastic tensorcheck pieces: the device-resident FoldedPolynomialTree against the stack-machine
    restatement - stream order, evaluate_folding, transcribe_foldings / partially_foldtree, open_folding."""
    import random

    from gemini_b200 import kzg, tensorcheck
    from util import rand_points

    rng = random.Random(5)
    for n, depth in ((37, 4), (64, 6), (5, 3), (1000, 7)):
        f_be = [rng.randrange(o.R) for _ in range(n)]
        ch = [rng.randrange(o.R) for _ in range(depth)]
        tree = tensorcheck.FoldedPolynomialTree(ctx, f_be, ch)
        assert tree.depth() == depth and len(tree) == n
        assert list(tree.iter()) == list(o.folded_polynomial_tree(f_be, ch))
        x = rng.randrange(o.R)
        assert tensorcheck.evaluate_folding(tree, x) == o.evaluate_folding(f_be, ch, x)
        want_levels = o.foldings_polynomial(f_be[::-1], ch + [0])
        assert tensorcheck.transcribe_foldings(tree, 1) == want_levels[1:]
        partial, transcribed = tensorcheck.partially_foldtree(ctx, f_be, ch)
        assert partial.depth() == depth and transcribed == []      # depth <= SPACE_TIME_THRESHOLD: nothing transcribed
        # open_folding: remainders and the batched evaluation proof
        srs_le = rand_points(n + 2, 77)
        cks = kzg.CommitterKeyStream(ctx, srs_le[::-1])
        points = [rng.randrange(o.R) for _ in range(3)]
        etas = [rng.randrange(o.R) for _ in range(depth)]
        rem, proof = cks.open_folding(tree, points, etas, 1 << 10)
        want_rem, want_proof = o.kzg_open_folding(srs_le[::-1], f_be, ch, points, etas, 1 << 10)
        assert rem == want_rem
        assert proof == want_proof
        # and the commitments of the same tree (commit_folding) for good measure
        assert cks.commit_folding(f_be, ch, 1 << 10) == o.kzg_commit_folding(srs_le[::-1], f_be, ch, 1 << 10)
