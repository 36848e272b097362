#!/usr/bin/env python
"""Config 4 of BASELINE.json: `examples/snark --time-prover -i LOGSIZE` (src/examples/snark.rs:69-79) on one B200.

dummy_r1cs(n = 2^logn) (all scalars identical - the reference's default, degenerate workload), SRS resident on
the device (+ precomputed table, key-setup time), Proof::new_time with a Merlin transcript on the host.
Prints one JSON line with the prover wall time and the phases the reference instruments with start_timer!."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_elastic(ctx, logn, reps=1):
    """`examples/snark -i LOGSIZE` (the elastic prover, src/examples/snark.rs:55-67 with max_msm_buffer = 1 << 20) on one
    GPU: snark::Proof::new_elastic over device-resident streams, and its proof must equal new_time's on the same
    instance (the reference's strongest test, src/snark/tests.rs:13-58)."""
    import numpy as np

    import gemini_b200 as gm
    from gemini_b200 import snark
    from gemini_b200.transcript import MerlinTranscript

    n = 1 << logn
    t0 = time.perf_counter()
    srs = ctx.srs_generate(n, first_multiple=1)
    le = srs.read()
    srs_be = ctx.srs_load(np.ascontiguousarray(le[::-1]))      # Reverse(powers_of_g), kzg/space.rs:288-297
    del le
    srs.precompute()
    srs_be.precompute()
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    ck, cks = gm.CommitterKey(ctx, srs), gm.CommitterKeyStream(ctx, srs_be)
    e = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % gm.field.R
    r1cs = snark.R1cs.dummy(ctx, n, e)
    want = snark.new_time(ctx, r1cs, ck, MerlinTranscript())
    best = None
    for _ in range(reps):
        timers = {}
        l0 = ctx.launch_count
        ctx.synchronize()
        t0 = time.perf_counter()
        proof = snark.new_elastic(ctx, r1cs, cks, MerlinTranscript(), 1 << 20, timers)
        ctx.synchronize()
        wall = time.perf_counter() - t0
        if best is None or wall < best[0]:
            best = (wall, timers, ctx.launch_count - l0)
    wall, timers, launches = best
    res = {"metric": "snark_elastic_prover_wall_s", "value": wall, "unit": "s", "logsize": logn, "n_gpus": 1,
           "workload": "examples/snark (elastic prover, MAX_MSM_BUFFER_LOG = 20): dummy_r1cs, big-endian streams resident on the device, Merlin on host",
           "phases_s": {k: round(v, 6) for k, v in timers.items()}, "gpu_launches": launches, "srs_setup_s": setup_s,
           "elastic proof == time proof": proof == want}
    assert proof == want, "elastic proof differs from the time proof"
    srs.free()
    srs_be.free()
    return res


def run_sharded(job, logn, reps=2):
    """BASELINE config 4 as written: `examples/snark --time-prover -i LOGSIZE` end to end on N GPUs.  The committer key is
    dealt out cyclically over the ranks (dist.ShardedCommitterKey): every commitment / opening is an MSM of len / N terms
    per rank plus one all-gather of 192-byte partial sums inside the library; the Fr side of the prover (sumchecks, folds,
    quotients) runs replicated on every rank with identical transcripts.  Wall clock, max over ranks."""
    import numpy as np

    import gemini_b200 as gm
    from gemini_b200 import dist as gdist
    from gemini_b200 import snark
    from gemini_b200.transcript import MerlinTranscript

    ctx = job.ctx
    n = 1 << logn
    t0 = time.perf_counter()
    full = ctx.srs_generate(n, first_multiple=1)
    sck = gdist.ShardedCommitterKey.from_full_key(ctx, full)
    full.free()
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    e = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % gm.field.R
    r1cs = snark.R1cs.dummy(ctx, n, e)
    best = None
    for _ in range(reps):
        timers = {}
        job.barrier()
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        proof = snark.new_time(ctx, r1cs, sck, MerlinTranscript(), timers)
        ctx.synchronize()
        wall = job.max_over_ranks(time.perf_counter() - t0)
        if best is None or wall < best[0]:
            best = (wall, timers, ctx.launch_count - l0)
    wall, timers, launches = best
    # every rank must hold the same proof: compare the evaluation proof and the witness commitment across ranks
    pt = proof["tensorcheck_proof"]["evaluation_proof"]
    sig = np.array([pt[0] & ((1 << 64) - 1), pt[1] & ((1 << 64) - 1), proof["witness_commitment"][0] & ((1 << 64) - 1)], dtype=np.uint64)
    rows = ctx.comm_allgather(sig)
    same = bool((rows == rows[0]).all())
    assert same, "ranks disagree on the proof"
    res = {"metric": "snark_time_prover_wall_s", "value": wall, "unit": "s", "logsize": logn, "n_gpus": job.world,
           "workload": "examples/snark --time-prover on N GPUs: dummy_r1cs, committer key dealt out cyclically, one ncclAllGather per commitment",
           "phases_s_rank0": {k: round(v, 6) for k, v in timers.items()}, "gpu_launches_rank0": launches, "srs_setup_s": setup_s,
           "all ranks hold the same proof": same}
    sck.srs.free()
    del r1cs, proof
    import gc

    gc.collect()
    return res


def run_sharded_elastic(job, logn, reps=1):
    """BASELINE config 5 end to end on N GPUs: `examples/snark -i LOGSIZE` (the elastic prover, max_msm_buffer = 1 << 20) with
    the committer key dealt out cyclically (dist.ShardedCommitterKeyStream).  The Fr side (streams, sumchecks, fold tree,
    quotients) runs replicated on resident vectors; every commitment / opening is one MSM of len / N terms per rank + one
    all-gather.  The proof must equal the time prover's on the same key (the reference's strongest test, snark/tests.rs:13-58)."""
    import numpy as np

    import gemini_b200 as gm
    from gemini_b200 import dist as gdist
    from gemini_b200 import snark
    from gemini_b200.transcript import MerlinTranscript

    ctx = job.ctx
    n = 1 << logn
    t0 = time.perf_counter()
    full = ctx.srs_generate(n, first_multiple=1)
    sck = gdist.ShardedCommitterKey.from_full_key(ctx, full)
    full.free()
    scks = gdist.ShardedCommitterKeyStream(sck, n)
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    e = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % gm.field.R
    r1cs = snark.R1cs.dummy(ctx, n, e)
    want = snark.new_time(ctx, r1cs, sck, MerlinTranscript())
    best = None
    for _ in range(reps):
        timers = {}
        job.barrier()
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        proof = snark.new_elastic(ctx, r1cs, scks, MerlinTranscript(), 1 << 20, timers)
        ctx.synchronize()
        wall = job.max_over_ranks(time.perf_counter() - t0)
        if best is None or wall < best[0]:
            best = (wall, timers, ctx.launch_count - l0)
    wall, timers, launches = best
    same_as_time = proof == want
    pt = proof["tensorcheck_proof"]["evaluation_proof"]
    sig = np.array([pt[0] & ((1 << 64) - 1), pt[1] & ((1 << 64) - 1), proof["witness_commitment"][0] & ((1 << 64) - 1), int(same_as_time)], dtype=np.uint64)
    rows = ctx.comm_allgather(sig)
    same = bool((rows == rows[0]).all()) and bool(rows[:, 3].all())
    assert same, "ranks disagree on the elastic proof, or it differs from the time prover's"
    res = {"metric": "snark_elastic_prover_wall_s", "value": wall, "unit": "s", "logsize": logn, "n_gpus": job.world,
           "workload": "examples/snark (elastic prover, MAX_MSM_BUFFER_LOG = 20) on N GPUs: dummy_r1cs, committer key dealt out cyclically, Fr streams resident and replicated",
           "phases_s_rank0": {k: round(v, 6) for k, v in timers.items()}, "gpu_launches_rank0": launches, "srs_setup_s": setup_s,
           "elastic proof == time proof on every rank": same}
    sck.srs.free()
    del r1cs, proof, want, scks
    import gc

    gc.collect()          # the extras that follow size their MSM passes from the free HBM
    return res


def run(ctx, logn, reps, no_precompute=False):
    import gemini_b200 as gm
    from gemini_b200 import snark
    from gemini_b200.transcript import MerlinTranscript

    class _A:
        pass

    args = _A()
    args.logn, args.reps, args.no_precompute = logn, reps, no_precompute
    n = 1 << args.logn
    t0 = time.perf_counter()
    srs = ctx.srs_generate(n, first_multiple=1)
    if not args.no_precompute:
        srs.precompute()
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    ck = gm.CommitterKey(ctx, srs)
    e = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % gm.field.R
    r1cs = snark.R1cs.dummy(ctx, n, e)
    best = None
    for rep in range(args.reps):
        timers = {}
        l0 = ctx.launch_count
        ctx.synchronize()
        t0 = time.perf_counter()
        proof = snark.new_time(ctx, r1cs, ck, MerlinTranscript(), timers)
        ctx.synchronize()
        wall = time.perf_counter() - t0
        if best is None or wall < best[0]:
            best = (wall, timers, ctx.launch_count - l0)
    wall, timers, launches = best
    srs_info = srs.precompute_info()
    res = ({
        "metric": "snark_time_prover_wall_s", "value": wall, "unit": "s", "logsize": args.logn, "n_gpus": 1,
        "workload": "examples/snark --time-prover: dummy_r1cs, Merlin transcript on host, all vectors device resident",
        "phases_s": {k: round(v, 6) for k, v in timers.items()}, "gpu_launches": launches,
        "srs_setup_s": setup_s, "srs_precompute": srs_info,
        "msm_terms": 3 * n, "proof_commitments": len(proof["tensorcheck_proof"]["folded_polynomials_commitments"]) + 2,
    })
    srs.free()
    return res


def main():
    import gemini_b200 as gm

    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=24)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-precompute", action="store_true")
    ap.add_argument("--sharded", choices=["time", "elastic"], default=None,
                    help="run under torchrun on N GPUs with the committer key dealt out cyclically (configs 4 / 5 of BASELINE.json)")
    args = ap.parse_args()
    if args.sharded:
        import bench

        job = bench.Job()
        fn = run_sharded if args.sharded == "time" else run_sharded_elastic
        res = fn(job, args.logn, min(args.reps, 2))
        if job.rank == 0:
            print(json.dumps(res))
        job.close()
        return
    ctx = gm.Context(0)
    print(json.dumps(run(ctx, args.logn, args.reps, args.no_precompute)))
    ctx.close()


if __name__ == "__main__":
    main()
