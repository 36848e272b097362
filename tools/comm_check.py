#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/comm_check.py - parity of the in-library exchange at N ranks (one GPU each):
an MSM sharded by contiguous point range (gm_msm_g1_sharded, streamed variant, world-wide all-gather) against the naive
sum over ALL points computed by the oracle on rank 0, and the sharded TimeProver over gm_comm_allgather against the
single-GPU prover.  Prints "comm_check ok" on rank 0."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]


def main():
    import numpy as np
    import torch
    import torch.distributed as dist

    import gemini_b200 as gm
    import pyref as o
    from gemini_b200 import dist as gdist
    from util import rand_points, rand_scalars

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gm.Context(local)
    ctx.comm_init_torch(device=f"cuda:{local}")
    assert (ctx.comm_rank, ctx.comm_world) == (rank, world)
    rows = ctx.comm_allgather(np.array([rank, 7 * rank + 1], dtype=np.uint64))
    assert rows.tolist() == [[r, 7 * r + 1] for r in range(world)]
    n = 1000 + 37
    pts, sc = rand_points(n, 11), rand_scalars(n, 12)
    sc[5] = 0
    sc[6] = o.R - 1
    lo, hi = gdist.shard_range(n, rank, world)
    srs = ctx.srs_load(pts[lo:hi])
    total = ctx.msm_sharded(srs, sc[lo:hi])
    st = gm.msm._DeviceStream(ctx, srs, 128)
    for s0 in range(0, hi - lo, 128):
        st.push_range(s0, sc[lo + s0:min(hi, lo + s0 + 128)])
    streamed = st.finalize_sharded_raw()
    assert np.array_equal(total, streamed), "streamed sharded MSM differs"
    all_tot = ctx.comm_allgather(total)
    assert all(np.array_equal(all_tot[r], total) for r in range(world)), "ranks disagree on the total"
    if rank == 0:
        assert gm.field.jacobian_to_affine(total) == o.naive_msm(pts, sc), "sharded MSM differs from the naive sum"
    # sharded sumcheck over the library communicator
    rng = random.Random(3)
    m = 1 << 10
    f, g = [rng.randrange(o.R) for _ in range(m)], [rng.randrange(o.R) for _ in range(m - 3)]
    tw = rng.randrange(o.R)
    ch = [rng.randrange(o.R) for _ in range(16)]
    start, B, L = gdist.sumcheck_block(len(f), len(g), rank, world)
    sp = gdist.ShardedTimeProver(lambda a, b, t: gm.TimeProver(ctx, a, b, t), f[start:start + B], g[start:start + B], tw, len(f), len(g),
                                 comm=gdist.LibComm(ctx))
    it = iter(ch)
    got = gm.Sumcheck.prove(sp, lambda msg: next(it))
    it = iter(ch)
    want = o.sumcheck_prove(o.TimeProver(f, g, tw), lambda msg: next(it))
    assert got.messages == want[0] and tuple(got.final_foldings[0]) == tuple(want[2]), "sharded sumcheck differs"
    # snark::Proof::new_time with the key dealt out cyclically over the ranks == the single-GPU proof == the oracle
    from gemini_b200 import snark
    from gemini_b200.transcript import MerlinTranscript

    nn = 64
    srs_pts = rand_points(2 * nn + 5, 21)
    full = ctx.srs_load(srs_pts)
    sck = gdist.ShardedCommitterKey.from_full_key(ctx, full)
    r1 = snark.R1cs.dummy(ctx, nn, 424242)
    got_p = snark.new_time(ctx, r1, sck, MerlinTranscript())
    want_p = snark.new_time(ctx, r1, gm.CommitterKey(ctx, full), MerlinTranscript())
    assert got_p == want_p, "sharded time prover differs from the single-GPU one"
    if rank == 0:
        assert got_p == o.snark_new_time(o.dummy_r1cs(424242, nn), srs_pts, o.MerlinTranscript()), "time prover differs from the oracle"
    # strided scalars without the exchange: rank r's share of a commitment, summed by hand
    v = gm.DeviceFr.from_host(ctx, sc[:300])
    part = ctx.msm_strided_dev(sck.srs, v.ptr + 32 * rank, (300 - rank + world - 1) // world, world, sharded=False)
    tot = ctx.g1_sum(ctx.comm_allgather(part))
    assert gm.field.jacobian_to_affine(tot) == o.naive_msm(srs_pts[:300], sc[:300]) == sck.commit(v)
    ctx.comm_barrier()
    if rank == 0:
        print(f"comm_check ok: world={world} nccl={gm.lib.gm_comm_nccl_version()}")
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
