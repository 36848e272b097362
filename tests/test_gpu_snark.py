"""Device vector helpers and the end-to-end time prover (snark::Proof::new_time) against the oracle."""
import hashlib
import random

import pytest

import gemini_b200 as gm
import pyref as o
from gemini_b200 import devvec, snark
from util import R, rand_points, rand_scalars

pytestmark = pytest.mark.gpu


class HashTranscript:
    """Deterministic stand-in for Merlin shared by the oracle and the device prover in these tests."""

    def __init__(self):
        self.h = hashlib.sha256(b"test-transcript")

    def append_serializable(self, label, obj):
        self.h.update(label + repr(obj).encode())

    def append_g1(self, label, point):
        self.h.update(label + b"G1" + repr(point).encode())

    def get_challenge(self, label):
        self.h.update(b"challenge" + label)
        return int.from_bytes(hashlib.sha512(self.h.digest()).digest(), "little") % R


@pytest.mark.parametrize("n", [1, 2, 5, 64, 1000, 5000])
def test_powers_eval_hadamard_axpy(ctx, n):
    f, g = rand_scalars(n, n), rand_scalars(n, n + 1)
    x = rand_scalars(1, 3)[0]
    df, dg = devvec.DeviceFr.from_host(ctx, f), devvec.DeviceFr.from_host(ctx, g)
    assert devvec.powers(ctx, x, n).to_ints() == o.powers(x, n)
    assert df.evaluate_pm(x) == (o.evaluate_le(f, x), o.evaluate_le(f, (-x) % R))
    assert df.hadamard(dg).to_ints() == o.hadamard(f, g)
    acc = df.clone().axpy(x, dg, n // 2)
    assert acc.to_ints() == [(a + x * b) % R if i < n // 2 else a for i, (a, b) in enumerate(zip(f, g))]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 9, 13])
def test_tensor(ctx, k):
    rho = rand_scalars(k, k)
    assert devvec.tensor(ctx, rho).to_ints() == o.tensor(rho)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 1024, 1025, 40000])
def test_div_linear(ctx, n):
    f, a = rand_scalars(n, n + 5), rand_scalars(1, 8)[0]
    q, rem = devvec.DeviceFr.from_host(ctx, f).div_linear(a)
    want_q = o.poly_div(f, [(-a) % R, 1])
    assert q.to_ints() == want_q and rem == o.evaluate_le(f, a)


def test_spmv(ctx):
    rng = random.Random(4)
    n = 300
    rows = [[(rng.randrange(R), rng.randrange(n)) for _ in range(rng.randrange(0, 4))] for _ in range(n)]
    z = rand_scalars(n, 5)
    dz = devvec.DeviceFr.from_host(ctx, z)
    assert devvec.DeviceCsr.from_rows(ctx, rows, n).matvec(dz).to_ints() == o.product_matrix_vector(rows, z)
    want_t = [0] * n
    for i, row in enumerate(rows):
        for v, c in row:
            want_t[c] = (want_t[c] + v * z[i]) % R
    assert devvec.DeviceCsr.from_rows(ctx, rows, n, transpose=True).matvec(dz).to_ints() == want_t


def _random_r1cs(n, seed):
    """A satisfiable random sparse R1CS: rows of A, B random; C = diag so that (Az)*(Bz) = Cz."""
    rng = random.Random(seed)
    z = [rng.randrange(1, R) for _ in range(n)]
    a = [[(rng.randrange(R), rng.randrange(n)) for _ in range(2)] for _ in range(n)]
    b = [[(rng.randrange(R), rng.randrange(n)) for _ in range(2)] for _ in range(n)]
    za, zb = o.product_matrix_vector(a, z), o.product_matrix_vector(b, z)
    c = [[(za[i] * zb[i] % R * pow(z[i], -1, R) % R, i)] for i in range(n)]
    return {"a": a, "b": b, "c": c, "z": z, "w": z[1:], "x": z[:1]}


@pytest.mark.parametrize("n,kind", [(8, "random"), (16, "dummy"), (64, "random"), (256, "dummy")])
def test_new_time_matches_oracle(ctx, n, kind):
    """snark/tests.rs pattern: the whole Proof must be equal, element by element."""
    srs = rand_points(2 * n + 1, 70)
    r1 = o.dummy_r1cs(rand_scalars(1, n)[0], n) if kind == "dummy" else _random_r1cs(n, n)
    want = o.snark_new_time(r1, srs, HashTranscript())
    ck = gm.CommitterKey(ctx, srs)
    dev_r1 = snark.R1cs.from_rows(ctx, r1["a"], r1["b"], r1["c"], r1["z"], r1["w"])
    timers = {}
    got = snark.new_time(ctx, dev_r1, ck, HashTranscript(), timers)
    assert got["witness_commitment"] == want["witness_commitment"]
    assert got["zc_alpha"] == want["zc_alpha"]
    assert got["first_sumcheck_msgs"] == want["first_sumcheck_msgs"]
    assert got["second_sumcheck_msgs"] == want["second_sumcheck_msgs"]
    assert got["tensorcheck_proof"] == want["tensorcheck_proof"]
    assert {"Commitment to w", "First sumcheck", "Second sumcheck", "Tensorcheck"} <= set(timers)
    if kind == "dummy":
        dd = snark.R1cs.dummy(ctx, n, r1["z"][0])
        assert snark.new_time(ctx, dd, ck, HashTranscript()) == got


def test_new_time_with_merlin_runs(ctx):
    from gemini_b200.transcript import MerlinTranscript

    n = 32
    srs = rand_points(2 * n + 1, 71)
    r1 = o.dummy_r1cs(12345, n)
    ck = gm.CommitterKey(ctx, srs)
    got = snark.new_time(ctx, snark.R1cs.dummy(ctx, n, 12345), ck, MerlinTranscript())
    want = o.snark_new_time(r1, srs, o.MerlinTranscript())
    assert got == want
