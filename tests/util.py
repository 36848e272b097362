"""Shared helpers of the parity tests: seeded inputs and numpy restatements of device generators."""
from __future__ import annotations

import random
from functools import lru_cache

import numpy as np

import pyref as o

R, Q = o.R, o.Q
M64 = (1 << 64) - 1


def rand_scalars(n, seed=0):
    rng = random.Random(seed)
    return [rng.randrange(R) for _ in range(n)]


@lru_cache(maxsize=None)
def _points(n, seed):
    rng = random.Random(seed)
    p = o.g1_mul(o.G1_GEN, rng.randrange(1, R))
    d = o.g1_mul(o.G1_GEN, rng.randrange(1, R))
    out = []
    for _ in range(n):
        out.append(p)
        p = o.g1_add(p, d)
    return tuple(out)


def rand_points(n, seed=0):
    """n distinct points P0 + i*D (cheap affine chain)."""
    return list(_points(n, seed))


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def fr_random_limbs(n: int, seed: int) -> np.ndarray:
    """numpy restatement of k_fr_random (gemini_b200/csrc/fr.cu): (n,4) uint64 Montgomery limbs."""
    with np.errstate(over="ignore"):
        idx = np.arange(4 * n, dtype=np.uint64) + np.uint64(seed)
        w = splitmix64(idx).reshape(n, 4)
    w[:, 3] &= np.uint64(0x7FFFFFFFFFFFFFFF)
    w = np.ascontiguousarray(w)
    r_l = [np.uint64((R >> (64 * j)) & M64) for j in range(4)]
    # lexicographic compare (most significant limb first) to find values >= r
    ge = np.zeros(n, dtype=bool)
    eq = np.ones(n, dtype=bool)
    for j in (3, 2, 1, 0):
        ge |= eq & (w[:, j] > r_l[j])
        eq &= w[:, j] == r_l[j]
    ge |= eq
    # values >= r: subtract r limb by limb (borrow chain) on the selected rows
    sel = np.nonzero(ge)[0]
    if sel.size:
        sub = w[sel]
        borrow = np.zeros(sel.size, dtype=np.uint64)
        for j in range(4):
            a = sub[:, j]
            t = a - r_l[j]
            b1 = (a < r_l[j]).astype(np.uint64)
            t2 = t - borrow
            b2 = (t < borrow).astype(np.uint64)
            sub[:, j] = t2
            borrow = b1 | b2
        w[sel] = sub
    return w


def limbs_to_ints(a: np.ndarray):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in a]
