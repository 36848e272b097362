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
