"""Host-side logic that needs no GPU: marshalling, KZG scalar preparation, sharding, gloo exchange."""
import os
import random
import sys

import numpy as np
import pytest

import pyref as o
from gemini_b200 import dist as gdist
from gemini_b200 import field, kzg
from util import R, fr_random_limbs, limbs_to_ints, rand_points, rand_scalars


def test_fr_marshalling_roundtrip():
    vals = rand_scalars(50, 1) + [0, 1, R - 1]
    arr = field.fr_to_limbs(vals)
    assert arr.shape == (53, 4) and arr.dtype == np.uint64
    assert field.fr_from_limbs(arr) == vals
    assert [o.limbs_to_int(r) for r in arr] == [o.fr_to_mont(v) for v in vals]
    assert field.fr_from_limbs(field.fr_to_limbs(vals, montgomery=False), montgomery=False) == vals


def test_g1_marshalling_roundtrip():
    pts = rand_points(8, 2) + [None]
    arr = field.g1_to_limbs(pts)
    assert field.g1_from_limbs(arr) == pts
    ark = field.g1_to_ark104(pts)
    assert ark.shape == (9, 104) and ark[8, 96] == 1 and not ark[:8, 96].any()
    for p in pts:
        assert field.jacobian_to_affine(field.affine_to_jacobian_limbs(p)) == p
    # non-normalised Jacobian input
    x, y = pts[0]
    z = 12345
    jac = np.array(field._limbs((x * z * z << 384) % field.Q, 6) + field._limbs((y * z ** 3 << 384) % field.Q, 6) +
                   field._limbs((z << 384) % field.Q, 6), dtype=np.uint64)
    assert field.jacobian_to_affine(jac) == pts[0]


def test_splitmix_twin_is_canonical():
    limbs = fr_random_limbs(5000, 3)
    vals = limbs_to_ints(limbs)
    assert all(0 <= v < R for v in vals) and len(set(vals)) == 5000


class _HostVec:
    """stand-in for DeviceFr on the CPU box: the host logic around the device divisions (Newton -> monomial remainder)"""

    def __init__(self, coeffs):
        self.c = [x % R for x in coeffs]
        self.n = len(self.c)

    def div_linear(self, a):
        q, acc = [], 0
        for c in reversed(self.c):
            acc = (acc * a + c) % R
            q.append(acc)
        rem = q.pop()
        return _HostVec(q[::-1]), rem


def test_kzg_scalar_preparation_matches_oracle():
    pts = rand_scalars(3, 4)
    assert kzg.vanishing_polynomial(pts) == o.vanishing_polynomial(pts)
    z = o.vanishing_polynomial(pts)
    for n in (20, 4, 3, 2, 0):
        f = rand_scalars(n, 5)
        q, rem = kzg._divide_by_points(_HostVec(f), pts)
        assert q.c == o.poly_div(f, z)
        # remainder, highest degree first: f - q * z
        prod = [0] * max(n, 3)
        for i, a in enumerate(q.c):
            for j, b in enumerate(z):
                prod[i + j] = (prod[i + j] + a * b) % R
        want = [((f[d] if d < n else 0) - prod[d]) % R for d in range(3)]
        assert rem == want[::-1]


def test_shard_ranges_cover():
    for n in (0, 1, 7, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            rs = [gdist.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in rs) - min(e - s for s, e in rs) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 41
        bases, scalars = rand_points(n, 8), rand_scalars(n, 9)
        s, e = gdist.shard_range(n, rank, world)
        # stand-in for the device: the oracle computes this rank's partial and the combine
        partial = field.affine_to_jacobian_limbs(o.naive_msm(bases[s:e], scalars[s:e]))

        def combine(allp):
            acc = None
            for row in allp:
                acc = o.g1_add(acc, field.jacobian_to_affine(row))
            return field.affine_to_jacobian_limbs(acc)

        total = gdist.allreduce_g1(partial, combine)
        q.put((rank, field.jacobian_to_affine(total) == o.naive_msm(bases, scalars)))
    finally:
        dist.destroy_process_group()


def test_allreduce_g1_gloo_world2():
    """The N>1 exchange (all-gather of 144-byte partial sums + adds on every rank) over gloo, world_size 2."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


class _OracleLocal:
    """Stand-in for the device TimeProver in the CPU tests: the oracle plus the ``state()`` accessor."""

    def __init__(self, f, g, twist):
        self.p = o.TimeProver(f, g, twist)

    def next_message(self, vm):
        return self.p.next_message(vm)

    def fold(self, r):
        self.p.fold(r)

    def state(self):
        return self.p.f, self.p.g, self.p.twist

    def final_foldings(self):
        return self.p.final_foldings()


def _gloo_sumcheck_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for nf, ng, twist in ((64, 64, 1), (37, 64, 12345), (50, 9, R - 2), (2, 2, 7), (5, 3, 3)):
            f, g = rand_scalars(nf, 10 + nf), rand_scalars(ng, 20 + ng)
            chal = rand_scalars(16, 30)
            it1, it2 = iter(chal), iter(chal)
            want = o.sumcheck_prove(o.TimeProver(f, g, twist), lambda m: next(it1))
            start, B, L = gdist.sumcheck_block(nf, ng, rank, world)
            sp = gdist.ShardedTimeProver(_OracleLocal, f[start:start + B], g[start:start + B], twist, nf, ng)
            got = o.sumcheck_prove(sp, lambda m: next(it2))
            ok = ok and got[0] == want[0] and got[1] == want[1] and got[2] == want[2] and sp.rounds() == L
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sharded_time_prover_gloo_world2():
    """f, g split into contiguous blocks, local folds, one 64-byte all-gather per round; the messages
    and final foldings must equal the single-process TimeProver's (sumcheck/time_prover.rs:83-137)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + random.randrange(2000)
    procs = [ctx.Process(target=_gloo_sumcheck_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_sumcheck_block_layout():
    assert gdist.sumcheck_block(1 << 24, 1 << 24, 3, 8) == (3 << 21, 1 << 21, 24)
    assert gdist.sumcheck_block(17, 5, 1, 2) == (16, 16, 5)
    with pytest.raises(ValueError):
        gdist.sumcheck_block(8, 8, 0, 3)
    with pytest.raises(ValueError):
        gdist.sumcheck_block(2, 2, 0, 4)


def test_msm_planner_decisions(monkeypatch):
    """Host-only planner introspection (gm_msm_describe_plan): the decisions the measured numbers of profiles/ rest on."""
    import ctypes as C

    from gemini_b200._lib import lib

    for k in ("GM_MSM_AFFINE", "GM_MSM_C", "GM_MSM_C_PRE", "GM_AFF_WPS", "GM_AFF_KEEP"):
        monkeypatch.delenv(k, raising=False)

    def plan(n, table):
        out = (C.c_int * 8)()
        assert lib.gm_msm_describe_plan(n, 1 if table else 0, 148, out) == 0
        return dict(zip(("c", "W", "nb", "merged", "levels", "G", "warps", "L"), list(out)))

    p20 = plan(1 << 20, True)
    assert (p20["c"], p20["W"], p20["nb"], p20["merged"]) == (20, 13, 1 << 19, 1)
    assert p20["levels"] == 2 and p20["G"] == 25 and p20["warps"] % 4 == 0     # >= 64 warps per SM in the first level (slots include the 2^levels alignment padding)
    assert p20["warps"] * 32 * p20["G"] >= (13 << 20) // 2
    p24 = plan(1 << 24, True)
    assert (p24["c"], p24["W"], p24["merged"], p24["levels"], p24["G"]) == (22, 12, 1, 4, 64)
    small = plan(1 << 12, True)
    assert small["levels"] == 0 and small["c"] >= 10                           # small MSMs stay on the XYZZ path
    plain = plan(1 << 20, False)
    assert plain["merged"] == 0 and plain["W"] * plain["c"] >= 255 and plain["levels"] >= 1
    monkeypatch.setenv("GM_MSM_AFFINE", "5")
    assert plan(1 << 12, True)["levels"] == 5
    monkeypatch.setenv("GM_MSM_AFFINE", "0")
    assert plan(1 << 24, True)["levels"] == 0


@pytest.mark.parametrize("world,sizes", [(4, [1 << 20, 300, 5000, 7, 1, 64, 2, 900, 33, 12, 100]), (8, [9, 8, 7, 6, 5, 4, 3, 2, 1]), (2, [5])])
def test_sharded_batch_commit_deals_short_polynomials_out(world, sizes):
    """dist.ShardedCommitterKey.batch_commit (host logic, no GPU): polynomials longer than the replicated prefix take the
    sharded commit, in order, on every rank; the short ones are dealt out whole - longest first, snake order - and every
    rank ends up with every commitment after one all-gather per round.  Ranks run as threads over a fake context whose
    'MSM' of polynomial k is [k + 1]G and whose all-gather is a barrier."""
    import threading

    points = [o.g1_mul(o.G1_GEN, k + 1) for k in range(len(sizes))]
    barrier = threading.Barrier(world)
    board = [None] * world
    local_calls = [[] for _ in range(world)]
    sharded_calls = [[] for _ in range(world)]

    class Poly:
        def __init__(self, k, n):
            self.k, self.n, self.ptr = k, n, k          # ptr doubles as the identity of the polynomial

    class FakeSrs:
        def __len__(self):
            return 1024

    class FakeCtx:
        def __init__(self, rank):
            self.rank = rank

        def msm_dev(self, srs, ptr, n, base_offset=0):
            local_calls[self.rank].append(ptr)
            return field.affine_to_jacobian_limbs(points[ptr])

        def comm_allgather(self, row):
            board[self.rank] = np.array(row, dtype=np.uint64).copy()
            barrier.wait()
            rows = np.stack(board)
            barrier.wait()
            return rows

    class Key(gdist.ShardedCommitterKey):
        def commit(self, v):                            # the sharded path: recorded, answered without a collective
            sharded_calls[self.rank].append(v.k)
            return points[v.k]

    results = [None] * world

    def run(rank):
        key = Key(FakeCtx(rank), None, rank=rank, world=world, prefix=FakeSrs())
        results[rank] = key.batch_commit([Poly(k, n) for k, n in enumerate(sizes)])

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert all(res == points for res in results)
    long_ones = [k for k, n in enumerate(sizes) if n > 1024]
    assert all(calls == long_ones for calls in sharded_calls)
    dealt = sorted(k for calls in local_calls for k in calls)
    assert dealt == [k for k, n in enumerate(sizes) if n <= 1024]          # every short one committed exactly once
    loads = [sum(sizes[k] for k in calls) for calls in local_calls]
    assert max(len(c) for c in local_calls) - min(len(c) for c in local_calls) <= 1, loads


def test_commit_lanes_return_results_in_job_order(monkeypatch):
    """kzg._msm_on_lanes (host logic, no GPU): the lanes pull from one queue, longest job first; every job runs exactly
    once, on some lane, and the results come back in JOB order whatever lane finished first."""
    import threading
    import time

    points = [o.g1_mul(o.G1_GEN, k + 1) for k in range(9)]
    ran = []
    lock = threading.Lock()

    class FakeCtx:
        def __init__(self, name):
            self.name, self._h, self.device_id = name, 1, 0

        def synchronize(self):
            pass

        def msm_dev(self, srs, ptr, n, base_offset=0):
            time.sleep(0.001 * (n % 7))                 # lanes finish out of order
            with lock:
                ran.append((self.name, ptr, base_offset))
            return field.affine_to_jacobian_limbs(points[ptr])

    class Owner:
        def __init__(self):
            self.ctx = FakeCtx("main")
            self.made = 0

        def _helper_contexts(self, count):
            self.made = max(self.made, count)
            return [FakeCtx(f"helper{k}") for k in range(count)]

    monkeypatch.setattr(kzg, "COMMIT_LANES", 3)
    owner = Owner()
    jobs = [(k, n, 100 + k) for k, n in enumerate([5, 900, 33, 0, 12, 4096, 7, 64, 1])]
    out = kzg._msm_on_lanes(owner, object(), jobs)
    assert out == [points[k] if n else None for k, n, _ in jobs]
    assert owner.made == 2                                                   # main context + two helpers
    assert sorted(p for _, p, _ in ran) == [k for k, n, _ in jobs if n]      # empty polynomials never reach a lane
    assert all(off == 100 + p for _, p, off in ran)
    assert ran[0][1] == 5 or ran[1][1] == 5 or ran[2][1] == 5                # the longest job starts first
