"""Sharded TimeProver with the DEVICE prover as the per-rank engine.  Two processes share cuda:0 (the
round-end GPU box has one GPU); the 64-byte exchanges go over gloo here and over NCCL in tools/bench_sumcheck.py."""
import os
import random

import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import gemini_b200 as gm
    import pyref as o
    from gemini_b200 import dist as gdist
    from util import R, rand_scalars

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = gm.Context(0)
    try:
        ok = True
        for nf, ng, twist in ((4096, 4096, 1), (3000, 4096, 987654321), (1000, 77, R - 5), (2, 2, 3)):
            f, g = rand_scalars(nf, 100 + nf), rand_scalars(ng, 200 + ng)
            chal = rand_scalars(20, 31)
            it1, it2 = iter(chal), iter(chal)
            want = o.sumcheck_prove(o.TimeProver(f, g, twist), lambda m: next(it1))
            start, B, L = gdist.sumcheck_block(nf, ng, rank, world)
            sp = gdist.ShardedTimeProver(lambda a, b, t: gm.TimeProver(ctx, a, b, t), f[start:start + B], g[start:start + B],
                                         twist, nf, ng)
            got = o.sumcheck_prove(sp, lambda m: next(it2))
            ok = ok and got[0] == want[0] and got[1] == want[1] and tuple(got[2]) == tuple(want[2])
        q.put((rank, ok))
    finally:
        ctx.close()
        dist.destroy_process_group()


def test_sharded_time_prover_device_world2():
    import torch.multiprocessing as mp

    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 33500 + random.randrange(2000)
    procs = [mpc.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_cyclic_committer_key_and_sharded_elastic_prover_on_one_gpu():
    """dist.ShardedCommitterKey / ShardedCommitterKeyStream without a communicator (world = 1 inside the library): the
    cyclic dealing is emulated by cutting BOTH shards of a world of two on this GPU and adding the two partial commitments
    with the oracle's group law; the stream key over the degenerate world-1 shard must drive snark::Proof::new_elastic to
    the proof of new_time with the plain CommitterKey (snark/tests.rs:13-58)."""
    import numpy as np

    import gemini_b200 as gm
    import pyref as o
    from gemini_b200 import dist as gdist
    from gemini_b200 import field, snark
    from gemini_b200.devvec import DeviceFr
    from gemini_b200.transcript import MerlinTranscript
    from util import rand_scalars

    ctx = gm.Context(0)
    try:
        n = 1 << 10
        full = ctx.srs_generate(n, first_multiple=1)
        ck = gm.CommitterKey(ctx, full)
        # (1) cyclic shards of a world of two, partial sums added on the host
        for m in (n, 777, 2, 1):
            v = DeviceFr.from_host(ctx, rand_scalars(m, 50 + m))
            parts = []
            for r in range(2):
                count = (n - r + 1) // 2
                shard = ctx.srs_subsample(full, r, 2, count)
                key = gdist.ShardedCommitterKey(ctx, shard, rank=r, world=2)
                local = (m - r + 1) // 2 if m > r else 0
                parts.append(field.jacobian_to_affine(ctx.msm_strided_dev(key.srs, v.ptr + 32 * r, local, 2, sharded=False)))
                shard.free()
            assert o.g1_add(parts[0], parts[1]) == ck.commit(v), m
        # (2) the elastic prover over the sharded stream key (world 1) == the time prover over the plain key
        e = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % gm.field.R
        r1cs = snark.R1cs.dummy(ctx, n, e)
        want = snark.new_time(ctx, r1cs, ck, MerlinTranscript())
        shard = ctx.srs_subsample(full, 0, 1, n)
        sck = gdist.ShardedCommitterKey(ctx, shard, rank=0, world=1)
        assert snark.new_time(ctx, r1cs, sck, MerlinTranscript()) == want
        scks = gdist.ShardedCommitterKeyStream(sck, n)
        got = snark.new_elastic(ctx, r1cs, scks, MerlinTranscript(), 1 << 20)
        assert got == want
        shard.free()
        full.free()
    finally:
        ctx.close()
