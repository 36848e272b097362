"""Parity of the device elastic prover composition (gemini_b200.snark.new_elastic, BASELINE config 5's prover) against
the oracle.  Mirrors the reference's strongest test (src/snark/tests.rs:13-58): elastic proof == time proof."""
import random

import pytest

import gemini_b200 as gm
import pyref as o
from gemini_b200 import kzg, snark
from test_gpu_snark import HashTranscript
from util import R, rand_points

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols,seed", [(8, 16, 1), (8, 8, 2), (5, 7, 3), (64, 64, 4)])
def test_elastic_proof_matches_oracle_and_time_proof(ctx, rows, cols, seed):
    rng = random.Random(seed)

    def matrix():
        m = []
        for _ in range(rows):
            cs = sorted(rng.sample(range(cols), rng.randrange(1, min(4, cols) + 1)))
            m.append([(rng.randrange(1, R), c) for c in cs])
        return m

    a, b, c = matrix(), matrix(), matrix()
    z = [rng.randrange(R) for _ in range(cols)]
    w = z[cols // 2:]
    srs_le = rand_points(rows + cols + 1, 500 + seed)
    want = o.snark_new_elastic({"a": a, "b": b, "c": c, "z": z, "w": w}, srs_le, HashTranscript(), 20)
    assert want == o.snark_new_time({"a": a, "b": b, "c": c, "z": z, "w": w}, srs_le, HashTranscript())
    r1cs = snark.R1cs.from_rows(ctx, a, b, c, z, w)
    cks = kzg.CommitterKeyStream(ctx, srs_le[::-1])
    got = snark.new_elastic(ctx, r1cs, cks, HashTranscript(), 20)
    assert got == want
    ck = gm.CommitterKey(ctx, srs_le)
    assert snark.new_time(ctx, r1cs, ck, HashTranscript()) == want
