"""snark::Proof::new_time on the device (/root/reference/src/snark/time_prover.rs:19-117) - BASELINE config 4.

The reference's time prover calls the hot path (commit, two sumchecks, tensorcheck with its fold chain, batch of
commitments and one batched opening) and glues the calls together with serial O(n) loops on the CPU.  Here every
vector lives in HBM from start to end: the glue (`product_matrix_vector`, `evaluate_le`, `tensor`, `powers`,
`hadamard`, the abc_tensored sums, `linear_combination`, the division by the vanishing polynomial) are device
kernels too, and only transcript traffic (64-byte round messages, 32-byte evaluations, 96-byte commitments)
crosses PCIe.  The Fiat-Shamir transcript is any object with ``append_serializable(label, fr_or_tuple)``,
``append_g1(label, point)`` and ``get_challenge(label) -> int`` (gemini_b200.transcript.MerlinTranscript).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import field
from .context import Context, Srs
from .devvec import DeviceCsr, DeviceFr, powers, tensor
from .kzg import CommitterKey
from .sumcheck import TimeProver

R = field.R


class R1cs:
    """R1cs<F> (src/circuit.rs:45-52) with its matrices (and their transposes) in CSR form on the device."""

    def __init__(self, ctx: Context, a: DeviceCsr, b: DeviceCsr, c: DeviceCsr, at: DeviceCsr, bt: DeviceCsr, ct: DeviceCsr,
                 z: DeviceFr, w: DeviceFr):
        self.ctx, self.a, self.b, self.c, self.at, self.bt, self.ct, self.z, self.w = ctx, a, b, c, at, bt, ct, z, w

    @classmethod
    def from_rows(cls, ctx: Context, a, b, c, z: Sequence[int], w: Sequence[int]) -> "R1cs":
        n = len(z)
        mats = [DeviceCsr.from_rows(ctx, m, n) for m in (a, b, c)]
        tmats = [DeviceCsr.from_rows(ctx, m, n, transpose=True) for m in (a, b, c)]
        return cls(ctx, *mats, *tmats, DeviceFr.from_host(ctx, z), DeviceFr.from_host(ctx, w))

    @classmethod
    def dummy(cls, ctx: Context, n: int, e: int) -> "R1cs":
        """circuit::dummy_r1cs (src/circuit.rs:349-365): A = B = C = diag(1/e), z = [e; n], w = [e; n-1]."""
        d = DeviceCsr.diagonal(ctx, n, pow(e, -1, R))
        ev = np.ascontiguousarray(np.broadcast_to(field.fr_to_limbs([e]), (n, 4)))
        z = DeviceFr.from_host(ctx, ev)
        w = DeviceFr.from_host(ctx, ev[: n - 1])
        return cls(ctx, d, d, d, d, d, d, z, w)


def _prove_sumcheck(transcript, prover) -> Dict:
    """Sumcheck::prove (src/subprotocols/sumcheck/proof.rs:36-66): one native call for device provers + Merlin."""
    from .sumcheck import Sumcheck

    sc = Sumcheck.prove_transcript(prover, transcript)
    return {"messages": sc.messages, "challenges": sc.challenges, "rounds": sc.rounds, "final_foldings": [tuple(sc.final_foldings[0])]}


def tensorcheck_new_time(transcript, ck: CommitterKey, base_polynomials: Sequence[DeviceFr], body_polynomials) -> Dict:
    """TensorcheckProof::new_time (src/subprotocols/tensorcheck/mod.rs:190-275) on device vectors.
    body_polynomials: list of (list of DeviceFr, challenges)."""
    ctx = ck.ctx
    max_len = max(len(polys) for polys, _ in body_polynomials)
    batch_challenge = transcript.get_challenge(b"batch_challenge")
    batch_challenges = [pow(batch_challenge, k, R) for k in range(max_len)]
    foldings: List[DeviceFr] = []
    for polys, chals in body_polynomials:
        n = max(p.n for p in polys)
        lc = DeviceFr.zeros(ctx, n)
        for p, c in zip(polys, batch_challenges):
            lc.axpy(c, p)
        foldings += lc.fold_chain(list(chals)[:-1])
    commitments = ck.batch_commit(foldings)          # 23 independent MSMs at logsize 24: four lanes, tails overlapped
    for c in commitments:
        transcript.append_g1(b"commitment", c)
    eval_chal = transcript.get_challenge(b"evaluation-chal")
    minus = (-eval_chal) % R
    eval_chal2 = eval_chal * eval_chal % R
    base_evals = []
    for p in base_polynomials:
        at_e, at_me = p.evaluate_pm(eval_chal)
        base_evals.append([p.evaluate(eval_chal2), at_e, at_me])
    fold_evals = [list(f.evaluate_pm(eval_chal)) for f in foldings]
    for row in base_evals + fold_evals:
        for e in row:
            transcript.append_serializable(b"eval", e)
    open_chal = transcript.get_challenge(b"open-chal")
    # CommitterKey::batch_open_multi_points (src/kzg/time.rs:134-159): eta-combination, quotient by the
    # vanishing polynomial of {e^2, e, -e} as three synthetic divisions, one MSM
    polys = list(base_polynomials) + foldings
    batched = DeviceFr.zeros(ctx, max(p.n for p in polys))
    eta = 1
    for p in polys:
        batched.axpy(eta, p)
        eta = eta * open_chal % R
    q = batched
    for pt in (eval_chal2, eval_chal, minus):
        q, _ = q.div_linear(pt)
    proof = ck.commit(q) if q.n else None
    return {"base_polynomials_evaluations": base_evals, "folded_polynomials_evaluations": fold_evals,
            "evaluation_proof": proof, "folded_polynomials_commitments": commitments}


def new_time(ctx: Context, r1cs: R1cs, ck: CommitterKey, transcript, timers: Optional[Dict[str, float]] = None) -> Dict:
    """snark::Proof::new_time.  ``timers`` (optional dict) receives the wall time of the phases the reference
    instruments with start_timer! ("Commitment to w", "First sumcheck", "Second sumcheck", "Tensorcheck")."""

    def lap(name, t0):
        if timers is not None:
            ctx.synchronize()
            timers[name] = timers.get(name, 0.0) + time.perf_counter() - t0

    t0 = time.perf_counter()
    z_a, z_b, z_c = r1cs.a.matvec(r1cs.z), r1cs.b.matvec(r1cs.z), r1cs.c.matvec(r1cs.z)
    lap("matrix-vector products", t0)

    t0 = time.perf_counter()
    witness_commitment = ck.commit(r1cs.w)
    lap("Commitment to w", t0)
    transcript.append_g1(b"witness", witness_commitment)
    alpha = transcript.get_challenge(b"alpha")

    t0 = time.perf_counter()
    zc_alpha = z_c.evaluate(alpha)
    lap("zc(alpha)", t0)
    transcript.append_serializable(b"zc(alpha)", zc_alpha)

    t0 = time.perf_counter()
    first = _prove_sumcheck(transcript, TimeProver(ctx, z_a, z_b, alpha))
    lap("First sumcheck", t0)

    t0 = time.perf_counter()
    b_ch = tensor(ctx, first["challenges"])
    c_ch = powers(ctx, alpha, b_ch.n)
    a_ch = b_ch.hadamard(c_ch)
    eta = transcript.get_challenge(b"eta")
    eta2 = eta * eta % R
    abc = r1cs.at.matvec(a_ch)
    abc.axpy(eta, r1cs.bt.matvec(b_ch))
    abc.axpy(eta2, r1cs.ct.matvec(c_ch))
    lap("abc_tensored", t0)

    t0 = time.perf_counter()
    second = _prove_sumcheck(transcript, TimeProver(ctx, abc, r1cs.z, 1))
    lap("Second sumcheck", t0)

    t0 = time.perf_counter()
    tc = tensorcheck_new_time(transcript, ck, [r1cs.w], [([abc, r1cs.z], second["challenges"])])
    lap("Tensorcheck", t0)
    return {"witness_commitment": witness_commitment, "zc_alpha": zc_alpha,
            "first_sumcheck_msgs": (first["messages"], first["final_foldings"]),
            "second_sumcheck_msgs": (second["messages"], second["final_foldings"]),
            "tensorcheck_proof": tc}


# ---------------------------------------------------------------------------------------------------
# Elastic prover (config 5): snark::Proof::new_elastic with every stream resident on the device
# ---------------------------------------------------------------------------------------------------
def elastic_tensorcheck(transcript, cks, witness, body, challenges: Sequence[int], max_msm_buffer: int) -> Dict:
    """``tensorcheck`` of snark/elastic_prover.rs:109-167.  ``witness`` / ``body``: big-endian streams
    (streams.ReverseStream / LinCombStream over resident vectors, or host sequences)."""
    from .streams import as_le_device
    from .tensorcheck import FoldedPolynomialTree, evaluate_folding

    ctx = cks.ctx
    chals = list(challenges)[:-1]                                              # strip_last
    tree = FoldedPolynomialTree(ctx, body, chals)
    commitments = cks.commit_folding(tree, chals, max_msm_buffer)              # kzg/space.rs:192-223
    for c in commitments:
        transcript.append_g1(b"commitment", c)
    eval_chal = transcript.get_challenge(b"evaluation-chal")
    points = [eval_chal * eval_chal % R, eval_chal, (-eval_chal) % R]
    at_pos, at_neg = evaluate_folding(tree, points[1]), evaluate_folding(tree, points[2])
    fold_evals = [[x, y] for x, y in zip(at_pos, at_neg)]
    w_le = as_le_device(ctx, witness)
    w_pos, w_neg = w_le.evaluate_pm(points[1])
    evaluations_w = [w_le.evaluate(points[0]), w_pos, w_neg]                    # evaluate_be at the three points
    for e in evaluations_w:
        transcript.append_serializable(b"eval", e)
    for row in fold_evals:
        for e in row:
            transcript.append_serializable(b"eval", e)
    open_chal = transcript.get_challenge(b"open-chal")
    open_chals = [pow(open_chal, k, R) for k in range(len(challenges) + 1)]
    _, proof_w = cks.open_multi_points(witness, points, max_msm_buffer)
    _, proof = cks.open_folding(tree, points, open_chals[1:], max_msm_buffer)
    total = field.jacobian_to_affine(ctx.g1_sum(np.stack([field.affine_to_jacobian_limbs(proof_w), field.affine_to_jacobian_limbs(proof)])))
    return {"base_polynomials_evaluations": [evaluations_w], "folded_polynomials_evaluations": fold_evals,
            "evaluation_proof": total, "folded_polynomials_commitments": commitments}


def new_elastic(ctx: Context, r1cs: R1cs, cks, transcript, max_msm_buffer: int = 1 << 20, timers: Optional[Dict[str, float]] = None) -> Dict:
    """snark::Proof::new_elastic (snark/elastic_prover.rs:169-267).  ``cks``: kzg.CommitterKeyStream (big-endian SRS).

    The reference's streams - ``Reverse(z)``, the column-major matrix streams, ``MatrixTensor``, ``LinCombStream`` - are
    device-resident here (gemini_b200.streams): a stream is a resident little-endian vector read backwards, a
    MatrixTensor is one tensor expansion + one transposed sparse product.  What differs from ``new_time`` is what the
    reference changes too: the prover flavour (ElasticProver: rounds from the MIN length, Space -> Time hand-off) and
    the KZG side (stream commit, commit_folding, open_multi_points + open_folding instead of one batched opening).
    Nothing but transcript traffic (messages, evaluations, commitments) crosses PCIe."""
    from .streams import LinCombStream, MatrixTensor, ReverseStream
    from .sumcheck import ElasticProver

    def lap(name, t0):
        if timers is not None:
            ctx.synchronize()
            timers[name] = timers.get(name, 0.0) + time.perf_counter() - t0

    t0 = time.perf_counter()
    z_a, z_b, z_c = r1cs.a.matvec(r1cs.z), r1cs.b.matvec(r1cs.z), r1cs.c.matvec(r1cs.z)
    lap("matrix-vector products", t0)
    t0 = time.perf_counter()
    witness = ReverseStream(r1cs.w)
    witness_commitment = cks.commit(witness, max_msm_buffer)
    lap("Commitment to w", t0)
    transcript.append_g1(b"witness", witness_commitment)
    alpha = transcript.get_challenge(b"alpha")
    zc_alpha = z_c.evaluate(alpha)                                               # evaluate_be(Reverse(z_c), alpha)
    transcript.append_serializable(b"zc(alpha)", zc_alpha)
    t0 = time.perf_counter()
    first = _prove_sumcheck(transcript, ElasticProver(ctx, ReverseStream(z_a), ReverseStream(z_b), alpha))
    lap("First sumcheck", t0)
    eta = transcript.get_challenge(b"eta")
    t0 = time.perf_counter()
    b_tensors = first["challenges"]
    c_tensors = [pow(alpha, 1 << k, R) for k in range(len(b_tensors))]           # powers2(alpha, len), misc.rs:68-77
    a_tensors = [x * y % R for x, y in zip(b_tensors, c_tensors)]                # hadamard
    a_alpha = MatrixTensor(ctx, r1cs.at, a_tensors)
    b_alpha = MatrixTensor(ctx, r1cs.bt, b_tensors)
    c_alpha = MatrixTensor(ctx, r1cs.ct, c_tensors)
    lhs = LinCombStream(ctx, [a_alpha, b_alpha, c_alpha], [1, eta, eta * eta % R])
    z_stream = ReverseStream(r1cs.z)
    lhs.le()
    lap("abc_tensored", t0)
    t0 = time.perf_counter()
    second = _prove_sumcheck(transcript, ElasticProver(ctx, lhs, z_stream, 1))
    lap("Second sumcheck", t0)
    batch_challenge = transcript.get_challenge(b"batch_challenge")
    t0 = time.perf_counter()
    body = LinCombStream(ctx, [lhs, z_stream], [1, batch_challenge])
    tc = elastic_tensorcheck(transcript, cks, witness, body, second["challenges"], max_msm_buffer)
    lap("Tensorcheck", t0)
    return {"witness_commitment": witness_commitment, "zc_alpha": zc_alpha,
            "first_sumcheck_msgs": (first["messages"], first["final_foldings"]),
            "second_sumcheck_msgs": (second["messages"], second["final_foldings"]),
            "tensorcheck_proof": tc}
