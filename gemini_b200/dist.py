"""Multi-GPU sharding: one process per GPU, contiguous point ranges, one exchange.

The reference is single-process; this is the new K7 step.  sum_i s_i P_i over disjoint
index ranges is independent per rank; the only exchange is the "all-reduce" of the per-rank partial G1
accumulators.  NCCL has no curve-addition reduction operator, so the all-reduce is realised as an
all-gather of the partial accumulators followed by world_size - 1 additions done identically on every rank.

The production exchange lives INSIDE the library (``gm_comm_init`` + ``gm_msm_g1_sharded`` /
``gm_comm_allgather``, gemini_b200/csrc/comm.cu): one ncclAllGather on the library's own stream, no torch and
no host staging (:class:`LibComm`, ``Context.msm_sharded``).  :class:`TorchComm` runs the same host logic over
a torch.distributed group - the ``gloo`` CPU tests of this module, and ``allreduce_g1`` for callers that
already hold normalised partials on the host.
"""
from __future__ import annotations

from typing import Callable, Sequence, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, end) of the contiguous point range owned by ``rank`` (even split, remainder to low ranks)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_g1(partial: np.ndarray, combine: Callable[[np.ndarray], np.ndarray], group=None, device=None) -> np.ndarray:
    """partial: (18,) uint64 Jacobian point of this rank -> sum over all ranks (same on every rank).

    ``device``: torch device the collective runs on ("cuda:k" for NCCL, None/"cpu" for gloo)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return partial
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(partial).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    allp = torch.stack(out).cpu().numpy().view(np.uint64)
    return combine(allp)


def sharded_msm(ctx, srs_shard, scalars_shard, group=None, device=None) -> np.ndarray:
    """Each rank: local MSM over its shard (device), then the exchange."""
    partial = ctx.msm(srs_shard, scalars_shard)
    return allreduce_g1(partial, ctx.g1_sum, group=group, device=device)


# ---------------------------------------------------------------------------------------------------
# Sumcheck / fold across ranks ("Sumcheck / fold (Time prover)")
# ---------------------------------------------------------------------------------------------------
def _ark_log2(x: int) -> int:
    return 0 if x <= 1 else (x - 1).bit_length()


def sumcheck_block(n_f: int, n_g: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Block layout of the sharded TimeProver: the index space is padded to 2^L, L = ceil(log2(max(|f|,|g|)))
    (sumcheck/time_prover.rs:35-38), and cut into ``world`` contiguous blocks of B = 2^L / world indices.
    Returns (start, B, L); rank r owns indices [start, start + B) of f and of g (clipped to their lengths,
    the remainder reads as zero - exactly the reference's zip-to-shorter / `unwrap_or(zero)` behaviour)."""
    if world & (world - 1):
        raise ValueError("world size must be a power of two")
    L = _ark_log2(max(n_f, n_g))
    if (1 << L) < world:
        raise ValueError("vectors shorter than the world size: run the replicated prover")
    B = (1 << L) // world
    return rank * B, B, L


class TorchComm:
    """all-gather of small rows over a torch.distributed group (gloo on CPU, NCCL through torch)"""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist

        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def allgather(self, row: np.ndarray) -> np.ndarray:
        """one small uint64 row per rank -> (world, len) uint64 (same on every rank)"""
        import torch
        import torch.distributed as dist

        t = torch.from_numpy(np.ascontiguousarray(row, dtype=np.uint64).view(np.int64).copy())
        if self.device is not None:
            t = t.to(self.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return torch.stack(out).cpu().numpy().view(np.uint64)


class LibComm:
    """the library's own NCCL communicator (gm_comm_init): all-gather on the context's stream, no torch"""

    def __init__(self, ctx):
        self.ctx = ctx
        self.rank, self.world = ctx.comm_rank, ctx.comm_world

    def allgather(self, row: np.ndarray) -> np.ndarray:
        return self.ctx.comm_allgather(row)


class ShardedTimeProver:
    """trait Prover (sumcheck/prover.rs:30-45) with f and g split across ranks; messages, challenges and final
    foldings are those of the single-process TimeProver (sumcheck/time_prover.rs:42-137) on the whole vectors.

    Every rank holds one zero-padded block of B = 2^L / world consecutive coefficients and runs an ordinary
    local prover on it (``make_local(f_block, g_block, twist)``: the device TimeProver in production, the
    oracle in the CPU tests).  Folding is purely local.  The pair i of the local block is the pair
    i + rank * B_k / 2 of the whole vector in round k (B_k = B / 2^k), and the twist of round k is
    twist^(2^k), so the local sums miss the factor (twist^(2^k))^(rank * B_k) = twist^(rank * B) - the same
    in every round.  The only exchange per round is an all-gather of the 64-byte local message; every rank
    then forms  sum_r twist^(r B) (a_r, b_r).  After log2(B) rounds one coefficient is left per rank: these
    are gathered (2 x 32 B per rank) and the last log2(world) rounds run replicated on every rank.
    """

    def __init__(self, make_local: Callable, f_block: Sequence, g_block: Sequence, twist: int, n_f: int, n_g: int,
                 group=None, device=None, modulus: int = None, comm=None):
        from . import field

        self._field = field
        self.R = modulus or field.R
        self.comm = comm if comm is not None else TorchComm(group, device)
        self.rank, self.world = self.comm.rank, self.comm.world
        self.start, self.B, self.L = sumcheck_block(n_f, n_g, self.rank, self.world)
        self.make_local = make_local
        self.twist0 = twist % self.R
        self._round = 0
        self.local_rounds = _ark_log2(self.B)
        self.scales = [pow(self.twist0, r * self.B, self.R) for r in range(self.world)]
        self.local = make_local(self._pad(f_block), self._pad(g_block), self.twist0)
        self.tail = None
        if self.local_rounds == 0:
            self._start_tail()

    def _pad(self, block):
        """host blocks are zero-padded to B here; device-resident blocks (DeviceFr) must already hold B elements"""
        if hasattr(block, "ptr") and hasattr(block, "n"):
            if block.n != self.B:
                raise ValueError("a device-resident block must hold exactly 2^L / world elements (zero-pad it on the device)")
            return block
        if len(block) > self.B:
            raise ValueError("block longer than 2^L / world")
        return list(block) + [0] * (self.B - len(block))

    # -- the two exchanges ------------------------------------------------------------------------
    def _combine(self, msg: Tuple[int, int]) -> Tuple[int, int]:
        rows = self.comm.allgather(self._field.fr_to_limbs(list(msg), montgomery=False).reshape(-1))
        a = b = 0
        for r in range(self.world):
            ar, br = self._field.fr_from_limbs(rows[r].reshape(2, 4), montgomery=False)
            a = (a + self.scales[r] * ar) % self.R
            b = (b + self.scales[r] * br) % self.R
        return a, b

    def _start_tail(self) -> None:
        f, g, tw = self.local.state()
        assert len(f) == 1 and len(g) == 1
        rows = self.comm.allgather(self._field.fr_to_limbs([f[0], g[0]], montgomery=False).reshape(-1))
        vals = [self._field.fr_from_limbs(rows[r].reshape(2, 4), montgomery=False) for r in range(self.world)]
        self.tail = self.make_local([v[0] for v in vals], [v[1] for v in vals], tw)

    # -- trait Prover -----------------------------------------------------------------------------
    def next_message(self, verifier_message):
        assert self._round <= self.L, "More rounds than needed."
        if self.tail is not None:
            msg = self.tail.next_message(verifier_message)
        elif self._round == self.local_rounds:          # the fold that leaves one coefficient per rank
            self.local.fold(verifier_message)
            self._start_tail()
            msg = self.tail.next_message(None)
        else:
            msg = self._combine(self.local.next_message(verifier_message))
        if msg is not None:
            self._round += 1
        return msg

    def fold(self, r: int) -> None:
        (self.tail or self.local).fold(r)

    def rounds(self) -> int:
        return self.L

    def round(self) -> int:
        return self._round

    def final_foldings(self):
        return self.tail.final_foldings() if self.tail is not None else None


# ---------------------------------------------------------------------------------------------------
# KZG across ranks: snark::Proof::new_time on N GPUs (BASELINE config 4)
# ---------------------------------------------------------------------------------------------------
class ShardedCommitterKey:
    """``CommitterKey`` (/root/reference/src/kzg/time.rs:24-160) whose ``powers_of_g`` are dealt out CYCLICALLY to the ranks
    of the library communicator: rank r holds g^(tau^r), g^(tau^(r+W)), ...  A commitment to a vector of any length m
    then costs every rank an MSM of ceil((m - r) / W) terms - the 23 fold levels of tensorcheck halve each time, and
    contiguous ranges would leave all but the first rank idle from the third level on - plus ONE all-gather of 192-byte
    partial sums inside the library (gm_msm_g1_strided_dev with sharded = 1).

    The vectors themselves are replicated: every rank runs the prover's Fr work (sumchecks, folds, quotients) on the full
    vectors and draws the same challenges, so no collective touches them.  The methods take resident ``DeviceFr``
    polynomials and mirror ``kzg.CommitterKey``."""

    # commitments of at most this many coefficients are NOT sharded: see batch_commit
    REPLICATED_PREFIX = 1 << 18

    def __init__(self, ctx, srs_shard, rank: int = None, world: int = None, prefix=None):
        self.ctx, self.srs = ctx, srs_shard
        self.rank = ctx.comm_rank if rank is None else rank
        self.world = ctx.comm_world if world is None else world
        self.prefix = prefix            # Srs: the first REPLICATED_PREFIX points of the FULL key, on every rank (optional)

    @classmethod
    def from_full_key(cls, ctx, full_srs, precompute: bool = True) -> "ShardedCommitterKey":
        """cut this rank's shard out of a resident full key (a deployment loads the shard straight from the host:
        gm_srs_load_g1 with stride_bytes = W * 104 and the pointer advanced by r records), plus the replicated prefix"""
        rank, world = ctx.comm_rank, ctx.comm_world
        n = len(full_srs)
        count = (n - rank + world - 1) // world if n > rank else 0
        shard = ctx.srs_subsample(full_srs, rank, world, count)
        prefix = None
        if world > 1 and n > 1:
            prefix = ctx.srs_subsample(full_srs, 0, 1, min(n, cls.REPLICATED_PREFIX))
        if precompute and count:
            shard.precompute()
            if prefix is not None:
                prefix.precompute()
        return cls(ctx, shard, rank, world, prefix)

    def max_degree(self) -> int:
        return len(self.srs) * self.world - 1        # upper bound; the exact length lives with whoever cut the shards

    def commit_raw(self, v) -> np.ndarray:
        m = v.n
        local = (m - self.rank + self.world - 1) // self.world if m > self.rank else 0
        return self.ctx.msm_strided_dev(self.srs, v.ptr + 32 * self.rank, local, self.world, sharded=True)

    def commit(self, v):
        from . import field

        return field.jacobian_to_affine(self.commit_raw(v))

    def batch_commit(self, polynomials):
        """Long polynomials: one sharded MSM each, in order (every rank issues the same sequence of collectives).
        SHORT ones (at most REPLICATED_PREFIX coefficients - 18 of the 23 fold levels of tensorcheck at logsize 24) are
        not worth a collective each: a sharded commitment costs about 1.6 ms of fixed latency (the tail of a small MSM
        + the exchange) whatever its size, so they are DEALT OUT whole - longest first, snake order - each to one rank,
        which commits it alone against the replicated prefix of the key, and the 144-byte results are exchanged
        afterwards, one all-gather per round of W commitments (8 GPUs, logsize 24: 38 -> 19 ms for the 23 levels)."""
        import numpy as np

        from . import field

        polys = list(polynomials)
        out = [None] * len(polys)
        short = [i for i, p in enumerate(polys) if self.prefix is not None and 0 < p.n <= len(self.prefix)]
        for i, p in enumerate(polys):
            if i not in short and p.n:
                out[i] = self.commit(p)
        if not short:
            return out
        short.sort(key=lambda i: -polys[i].n)
        W = self.world
        rounds = [short[k:k + W] for k in range(0, len(short), W)]
        # snake: position j of round r goes to rank j (even r) or W - 1 - j (odd r)
        owner = lambda r, j: j if r % 2 == 0 else W - 1 - j
        mine = []
        for r, group in enumerate(rounds):
            j = self.rank if r % 2 == 0 else W - 1 - self.rank
            mine.append(group[j] if j < len(group) else None)
        # all of this rank's commitments first (they overlap with the other ranks' work), then the exchanges
        local = [self.ctx.msm_dev(self.prefix, polys[i].ptr, polys[i].n) if i is not None else np.zeros(18, dtype=np.uint64) for i in mine]
        for r, group in enumerate(rounds):
            rows = self.ctx.comm_allgather(local[r])
            for j, i in enumerate(group):
                out[i] = field.jacobian_to_affine(rows[owner(r, j)])
        return out

    def open(self, polynomial, evaluation_point: int):
        from . import field

        if polynomial.n == 0:
            return 0, None
        q, evaluation = polynomial.div_linear(evaluation_point % field.R)
        return evaluation, self.commit(q)

    def open_multi_points(self, polynomial, eval_points):
        from .kzg import _divide_by_points

        q, _ = _divide_by_points(polynomial, eval_points)
        return self.commit(q)


class ShardedCommitterKeyStream:
    """``CommitterKeyStream`` (/root/reference/src/kzg/space.rs:59-297) over a :class:`ShardedCommitterKey`: the elastic
    prover (snark::Proof::new_elastic, BASELINE config 5) on N GPUs.

    The reference's stream key pairs the big-endian coefficient stream with ``Reverse(powers_of_g)`` so that the coefficient
    of degree d meets g^(tau^d) (space.rs:22-55: the leading surplus bases are skipped).  The streams are resident
    little-endian vectors here, so every stream commitment IS the time commitment of the little-endian vector, and the
    cyclically dealt key serves it with one MSM of len / N terms per rank + one all-gather (``ShardedCommitterKey.commit``).
    Everything else - quotients, remainders, the eta-combination of ``open_folding`` - is ``kzg.CommitterKeyStream``'s own
    code: this class only replaces where the MSMs run.  Commitments are issued one after the other in the same order on
    every rank (each is a collective)."""

    def __init__(self, sharded_key: ShardedCommitterKey, length: int):
        self.sck, self.ctx, self._n = sharded_key, sharded_key.ctx, length

    def __len__(self) -> int:
        return self._n

    def _le(self, stream_be):
        from .streams import as_le_device

        return as_le_device(self.ctx, stream_be)

    def _commit_le(self, le, max_msm_buffer: int):
        if le.n == 0:
            return None
        assert le.n <= len(self), "polynomial longer than the SRS"
        return self.sck.commit(le)

    def commit(self, polynomial_be, step: int = 1 << 20):
        return self._commit_le(self._le(polynomial_be), step)

    def open(self, polynomial_be, alpha: int, max_msm_buffer: int):
        from . import field

        le = self._le(polynomial_be)
        if le.n == 0:
            return 0, None
        q, evaluation = le.div_linear(alpha % field.R)
        return evaluation, self._commit_le(q, max_msm_buffer)

    def open_multi_points(self, polynomial_be, points, max_msm_buffer: int):
        from .kzg import _divide_by_points

        le = self._le(polynomial_be)
        assert le.n >= len(points)
        q, rem = _divide_by_points(le, points)
        return rem, self._commit_le(q, max_msm_buffer)

    def commit_folding(self, polynomials_be, challenges, max_msm_buffer: int):
        from . import field
        from .tensorcheck import FoldedPolynomialTree

        if isinstance(polynomials_be, FoldedPolynomialTree):
            levels = polynomials_be.levels
        else:
            if len(challenges) == 0:
                return []
            levels = self._le(polynomials_be).fold_chain([c % field.R for c in challenges])
        return self.sck.batch_commit(levels)

    def open_folding(self, polynomials, points, etas, max_msm_buffer: int = 0):
        from .kzg import CommitterKeyStream

        return CommitterKeyStream.open_folding(self, polynomials, points, etas, max_msm_buffer)
