#!/usr/bin/env python
"""Sharded TimeProver over NCCL: torchrun --nproc-per-node N tools/dist_sumcheck.py [--logn 24]

Every rank owns one block of 2^logn / N coefficients of f and g (device-resident, generated on the device from the
global counter stream, so the blocks are the slices of ONE global vector).  First a small instance is checked
message by message against the single-GPU TimeProver on rank 0, then the large one is timed (wall clock around
the whole Fiat-Shamir loop, max over ranks).  Rank 0 prints one JSON line."""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import gemini_b200 as gm
    from gemini_b200 import dist as gdist
    from gemini_b200 import field
    from gemini_b200.devvec import DeviceFr

    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=24)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    for k, v in (("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29633"), ("RANK", "0"), ("WORLD_SIZE", "1")):
        os.environ.setdefault(k, v)      # plain `python tools/dist_sumcheck.py` = one rank
    dist.init_process_group("nccl" if world > 1 else "gloo", device_id=torch.device("cuda", local) if world > 1 else None)
    dev = f"cuda:{local}" if world > 1 else None
    ctx = gm.Context(local)
    comm = None
    if world > 1:
        ctx.comm_init_torch(device=dev)          # the library's own NCCL communicator: gm_comm_allgather carries the round messages
        comm = gdist.LibComm(ctx)
    rng = random.Random(5)
    twist = rng.randrange(field.R)
    chal = [rng.randrange(field.R) for _ in range(64)]

    def blocks(logn, seed_f, seed_g):
        n = 1 << logn
        start, B, L = gdist.sumcheck_block(n, n, rank, world)
        return n, DeviceFr.random(ctx, B, seed_f + 4 * start), DeviceFr.random(ctx, B, seed_g + 4 * start)

    def prove(n, fb, gb):
        sp = gdist.ShardedTimeProver(lambda a, b, t: gm.TimeProver(ctx, a, b, t), fb, gb, twist, n, n, device=dev, comm=comm)
        it = iter(chal)
        return gm.Sumcheck.prove(sp, lambda m: next(it))

    # ---- parity on a small instance ----
    n, fb, gb = blocks(16, 1000, 2000)
    got = prove(n, fb, gb)
    ok = True
    if rank == 0:
        it = iter(chal)
        want = gm.Sumcheck.prove(gm.TimeProver(ctx, DeviceFr.random(ctx, n, 1000), DeviceFr.random(ctx, n, 2000), twist), lambda m: next(it))
        ok = got.messages == want.messages and tuple(got.final_foldings[0]) == tuple(want.final_foldings[0])
    # ---- timing ----
    n, fb, gb = blocks(args.logn, 3000, 4000)
    times = []
    for _ in range(args.reps + 1):
        f2, g2 = fb.clone(), gb.clone()
        dist.barrier()
        ctx.synchronize()
        t0 = time.perf_counter()
        sc = prove(n, f2, g2)
        ctx.synchronize()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times[1:])], dtype=torch.float64, device=dev or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item()) * 1e3
        print(json.dumps({"tool": "dist_sumcheck", "n_gpus": world, "n": n, "rounds": len(sc.messages), "wall_ms": ms,
                          "elements_per_s": 2 * n / (ms / 1e3), "parity_small_instance": ok,
                          "exchange": "gm_comm_allgather (ncclAllGather on the library stream) of 64 B per round + 64 B hand-off, replicated last log2(N) rounds"}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    if not ok:
        raise SystemExit("sharded messages differ from the single-GPU prover")


if __name__ == "__main__":
    main()
