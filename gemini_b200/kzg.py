"""KZG committer keys over the device MSM.

  * :class:`CommitterKey`       - /root/reference/src/kzg/time.rs:24-160 (in-memory, little-endian coefficients)
  * :class:`CommitterKeyStream` - /root/reference/src/kzg/space.rs:59-297 (big-endian streams, chunked MSM)

Everything runs on the device: the MSMs against the resident SRS, and the O(n) scalar preparation around them -
the synthetic divisions of ``open`` / ``open_multi_points`` (``DeviceFr.div_linear``: a parallel suffix-Horner scan;
division by the vanishing polynomial of k points = k successive linear divisions), the eta-combinations
(``DeviceFr.axpy``) and the big-endian <-> little-endian reversals (``DeviceFr.reversed``).  Polynomials may be
passed as host sequences / limb arrays (uploaded once) or as resident :class:`DeviceFr` /
:class:`streams.ReverseStream` objects (nothing crosses PCIe but the 32-byte evaluations and 144-byte proofs).
"""
from __future__ import annotations

import os

from typing import List, Sequence, Tuple

import numpy as np

from . import field
from .context import Context, Srs
from .devvec import DeviceFr
from .msm import ChunkedPippenger, HashMapPippenger, VariableBaseMSM, _DeviceStream, msm_chunks  # noqa: F401

R = field.R

# Device-resident scalars need no staging buffer: ``max_msm_buffer`` is the reference's bound on HOST memory
# (ChunkedPippenger::with_size), and cutting a vector that already sits in HBM into 2^20-term pieces only costs
# throughput (measured: the 2^24-term witness commitment of the elastic prover 0.129 s in 16 chunks, 0.043 s in one).
# So a resident vector is pushed in pieces of max(max_msm_buffer, MIN_DEVICE_CHUNK) terms - one piece unless the
# vector exceeds the 2^27-term pass limit; tests lower the floor to walk chunk boundaries (the result does not depend on it).
MIN_DEVICE_CHUNK = 1 << 27
# batch_commit of resident polynomials on COMMIT_LANES lanes (see CommitterKey.batch_commit); False = the reference's sequential map
CONCURRENT_COMMITS = True
COMMIT_LANES = int(os.environ.get("GM_COMMIT_LANES", "4"))   # host threads / streams of CommitterKey.batch_commit


def vanishing_polynomial(points: Sequence[int]) -> List[int]:
    """kzg/mod.rs:262-268 (little-endian coefficients)."""
    poly = [1]
    for p in points:
        nxt = [0] * (len(poly) + 1)
        for i, c in enumerate(poly):
            nxt[i] = (nxt[i] - p * c) % R
            nxt[i + 1] = (nxt[i + 1] + c) % R
        poly = nxt
    return poly


def _powers(x: int, n: int) -> List[int]:
    out, cur = [], 1
    for _ in range(n):
        out.append(cur)
        cur = cur * x % R
    return out


def _resident(ctx: Context, polynomial) -> DeviceFr:
    """little-endian coefficients on the device (no copy when they already are)"""
    return polynomial if isinstance(polynomial, DeviceFr) else DeviceFr.from_host(ctx, polynomial)


def _divide_by_points(poly: DeviceFr, points: Sequence[int]):
    """quotient of poly by prod (X - p) as len(points) successive synthetic divisions on the device, and the
    remainder in MONOMIAL form, highest degree first (the order of the reference's state deque, space.rs:128-166).
    The k constants c_j of the divisions are the Newton form  rem = c_1 + c_2 (X - p_1) + c_3 (X - p_1)(X - p_2) ..."""
    pts = [p % R for p in points]
    k = len(pts)
    q, cs = poly, []
    for a in pts:
        if q.n == 0:
            cs.append(0)
            continue
        q, c = q.div_linear(a)
        cs.append(c)
    rem = [0] * k
    basis = [1]
    for j, c in enumerate(cs):
        for d, b in enumerate(basis):
            rem[d] = (rem[d] + c * b) % R
        if j + 1 < k:
            nb = [0] * (len(basis) + 1)
            for d, b in enumerate(basis):
                nb[d + 1] = (nb[d + 1] + b) % R
                nb[d] = (nb[d] - pts[j] * b) % R
            basis = nb
    return q, rem[::-1]


def _msm_on_lanes(owner, srs: Srs, jobs) -> List[field.Point]:
    """jobs: (device pointer of the scalars, n, base_offset) per MSM against ``srs``; results in job order.  The MSMs run
    on COMMIT_LANES lanes - the owner's context and helper contexts of the same GPU (kept on ``owner``), each with its own
    stream and scratch arena and one host thread - that pull from one queue, longest job first."""
    import threading

    ctx = owner.ctx
    lanes = [ctx] + owner._helper_contexts(min(COMMIT_LANES, len(jobs)) - 1)
    ctx.synchronize()                       # the scalars were produced on this context's stream
    order = sorted(range(len(jobs)), key=lambda i: -jobs[i][1])
    out: List = [None] * len(jobs)
    errors: List = []
    lock = threading.Lock()
    cursor = [0]

    def work(lane: int) -> None:
        try:
            c = lanes[lane]
            while True:
                with lock:
                    k = cursor[0]
                    cursor[0] += 1
                if k >= len(order):
                    return
                ptr, n, off = jobs[order[k]]
                out[order[k]] = field.jacobian_to_affine(c.msm_dev(srs, ptr, n, base_offset=off)) if n else None
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(lanes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return out


def _helper_contexts_of(owner, count: int) -> List[Context]:
    hs = [h for h in getattr(owner, "_helpers", []) if h._h]
    while len(hs) < count:
        hs.append(Context(owner.ctx.device_id))
    owner._helpers = hs
    return hs[:count]


class CommitterKey:
    """``CommitterKey<E>``: ``powers_of_g`` resident on the device."""

    def __init__(self, ctx: Context, powers_of_g):
        self.ctx = ctx
        self.srs = powers_of_g if isinstance(powers_of_g, Srs) else ctx.srs_load(powers_of_g)
        self._msm = VariableBaseMSM(ctx)

    @classmethod
    def new(cls, ctx: Context, max_degree: int, max_eval_points: int, rng, precompute: bool = False) -> "CommitterKey":
        """``CommitterKey::new`` (time.rs:49-72), G1 half: tau and g are drawn from ``rng`` (``rng.randrange``), then
        powers_of_g[i] = tau^i * g for i <= max_degree by the device fixed-base MSM.  powers_of_g2 (``max_eval_points``
        G2 elements, used by the verifier only) is outside this path; tau and g are kept for tests."""
        tau = rng.randrange(1, R)
        k = rng.randrange(1, R)
        g = ctx.srs_generate(1, first_multiple=k & ((1 << 64) - 1) or 1).points()[0]     # a random-looking G1 element
        ck = cls(ctx, ctx.srs_setup(g, tau, max_degree + 1))
        ck.tau, ck.g, ck.max_eval_points = tau, g, max_eval_points
        if precompute:
            ck.srs.precompute()
        return ck

    def max_degree(self) -> int:
        return len(self.srs) - 1

    def _commit_dev(self, v: DeviceFr) -> field.Point:
        return field.jacobian_to_affine(self.ctx.msm_dev(self.srs, v.ptr, v.n)) if v.n else None

    def commit(self, polynomial) -> field.Point:
        """time.rs:81-83: one MSM against the SRS prefix (silently truncating to the shorter)."""
        if isinstance(polynomial, DeviceFr):
            return self._commit_dev(polynomial)
        return self._msm.msm_unchecked(self.srs, polynomial)

    def commit_raw(self, polynomial) -> np.ndarray:
        if isinstance(polynomial, DeviceFr):
            return self.ctx.msm_dev(self.srs, polynomial.ptr, polynomial.n)
        return self.ctx.msm(self.srs, polynomial)

    def index_by(self, indices: Sequence[int]) -> "CommitterKey":
        """time.rs:86-95: new key with powers_of_g[i] = sum of the g_j whose index is i (identity elsewhere).
        Only psnark uses it (outside this tier's hot path): the group sums run on the device, one
        gm_g1_sum per non-trivial target."""
        pts = self.srs.read()
        n = pts.shape[0]
        groups: dict = {}
        for j, i in enumerate(list(indices)[:n]):
            groups.setdefault(i, []).append(j)
        out = np.zeros_like(pts)
        one = field.affine_to_jacobian_limbs((0, 0))[12:18]
        for i, js in groups.items():
            if len(js) == 1:
                out[i] = pts[js[0]]
                continue
            jac = np.zeros((len(js), 18), dtype=np.uint64)
            for k, j in enumerate(js):
                if pts[j].any():
                    jac[k, :12] = pts[j]
                    jac[k, 12:] = one
                else:
                    jac[k] = field.affine_to_jacobian_limbs(None)
            tot = self.ctx.g1_sum(jac)
            if tot[12:].any():
                out[i] = tot[:12]
        return CommitterKey(self.ctx, out)

    def batch_commit(self, polynomials) -> List[field.Point]:
        """time.rs:98-107 (a sequential map of ``commit`` in the reference).  Resident polynomials are committed on
        COMMIT_LANES lanes - this context and helper contexts of the same GPU, each with its own stream and scratch
        arena, driven from one host thread each: an MSM ends in a latency-bound reduction tail that uses a handful of SMs
        (a quarter of the call at 2^20 terms, 0.6 - 1.2 ms whatever the size below 2^15: DESIGN.md 4.2), and the tails of
        the short commitments then overlap the bucket accumulation of the long ones.  The lanes pull from one queue,
        longest polynomial first.  Results are in input order and identical to the sequential map."""
        polys = list(polynomials)
        if len(polys) < 2 or not all(isinstance(p, DeviceFr) for p in polys) or not CONCURRENT_COMMITS:
            return [self.commit(p) for p in polys]
        return _msm_on_lanes(self, self.srs, [(p.ptr, p.n, 0) for p in polys])

    def _helper_contexts(self, count: int) -> List[Context]:
        return _helper_contexts_of(self, count)

    def open(self, polynomial, evaluation_point: int) -> Tuple[int, field.Point]:
        """time.rs:112-131: (evaluation, proof).  The reference runs a serial Horner recurrence and builds the quotient
        with ``Vec::insert(0, ..)`` (O(n^2)); here the same quotient and remainder come from one parallel synthetic
        division on the device, then one MSM over the quotient."""
        f = _resident(self.ctx, polynomial)
        if f.n == 0:
            return 0, None
        q, evaluation = f.div_linear(evaluation_point % R)
        return evaluation, self._commit_dev(q)

    def open_multi_points(self, polynomial, eval_points: Sequence[int]) -> field.Point:
        """time.rs:134-145: commit to polynomial / vanishing_polynomial(eval_points)."""
        q, _ = _divide_by_points(_resident(self.ctx, polynomial), eval_points)
        return self._commit_dev(q)

    def batch_open_multi_points(self, polynomials, eval_points: Sequence[int], eval_chal: int) -> field.Point:
        """time.rs:149-159: eta-combination of the polynomials (misc::linear_combination), then open_multi_points."""
        polys = [_resident(self.ctx, p) for p in polynomials]
        if not polys:
            return None
        batched = DeviceFr.zeros(self.ctx, max(p.n for p in polys))
        for p, eta in zip(polys, _powers(eval_chal % R, len(polys))):
            batched.axpy(eta, p)
        return self.open_multi_points(batched, eval_points)


def _folded_len(n: int, k: int) -> int:
    return (n + (1 << k) - 1) >> k


class CommitterKeyStream:
    """``CommitterKeyStream``: the SRS in BIG-endian stream order (``Reverse(powers_of_g)``, space.rs:288-297).

    Polynomial arguments are big-endian streams: host sequences / limb arrays (the reference's order), or a
    :class:`streams.ReverseStream` / ``MatrixTensor`` / ``LinCombStream`` over resident vectors."""

    def __init__(self, ctx: Context, powers_of_g_be):
        self.ctx = ctx
        if isinstance(powers_of_g_be, Srs):
            self.srs_be = powers_of_g_be
        else:
            self.srs_be = ctx.srs_load(powers_of_g_be)

    @classmethod
    def from_committer_key(cls, ck: CommitterKey) -> "CommitterKeyStream":
        le = ck.srs.read()
        return cls(ck.ctx, np.ascontiguousarray(le[::-1]))

    def __len__(self) -> int:
        return len(self.srs_be)

    # -- helpers -------------------------------------------------------------------------------
    def _le(self, stream_be) -> DeviceFr:
        from .streams import as_le_device

        return as_le_device(self.ctx, stream_be)

    def _commit_le(self, le: DeviceFr, max_msm_buffer: int) -> field.Point:
        """sum_d le[d] * g^(tau^d) against the big-endian SRS: the coefficients are reversed on the device and pushed in
        chunks (msm_chunks / ChunkedPippenger semantics: the result does not depend on the chunking)."""
        m = le.n
        if m == 0:
            return None
        assert m <= len(self), "polynomial longer than the SRS"
        be = le.reversed()
        chunk = max(max_msm_buffer, MIN_DEVICE_CHUNK, 1)
        off = len(self) - m
        if m <= chunk:
            # one piece: the one-shot entry point (its own plan for m terms, constant-vector test included)
            out = field.jacobian_to_affine(self.ctx.msm_dev(self.srs_be, be.ptr, m, base_offset=off))
            be.free()
            return out
        st = _DeviceStream(self.ctx, self.srs_be, min(chunk, m))
        for s0 in range(0, m, chunk):
            st.push_dev(off + s0, be.ptr + 32 * s0, min(chunk, m - s0))
        out = st.finalize()
        st.free()
        be.free()
        return out

    # -- the reference's methods -----------------------------------------------------------------
    def commit(self, polynomial_be, step: int = 1 << 20) -> field.Point:
        """space.rs:169-177 -> msm_chunks (space.rs:22-55)."""
        from .streams import LinCombStream, MatrixTensor, ReverseStream

        if isinstance(polynomial_be, (ReverseStream, MatrixTensor, LinCombStream)):
            return self._commit_le(self._le(polynomial_be), step)
        return msm_chunks(self.ctx, self.srs_be, polynomial_be, step)

    def open(self, polynomial_be, alpha: int, max_msm_buffer: int) -> Tuple[int, field.Point]:
        """space.rs:95-125: (evaluation, proof); the quotient never leaves the device."""
        le = self._le(polynomial_be)
        if le.n == 0:
            return 0, None
        q, evaluation = le.div_linear(alpha % R)
        return evaluation, self._commit_le(q, max_msm_buffer)

    def open_multi_points(self, polynomial_be, points: Sequence[int], max_msm_buffer: int):
        """space.rs:128-166: returns (remainder coefficients, highest degree first, and the proof)."""
        le = self._le(polynomial_be)
        assert le.n >= len(points), "the reference reads len(points) leading coefficients (unwrap on a shorter stream panics)"
        q, rem = _divide_by_points(le, points)
        return rem, self._commit_le(q, max_msm_buffer)

    def commit_folding(self, polynomials_be, challenges: Sequence[int], max_msm_buffer: int) -> List[field.Point]:
        """space.rs:192-223.  The reference walks the FoldedPolynomialTree once and feeds one ChunkedPippenger per
        level; here the device fold chain produces every level and each is committed against the SRS range that lines
        its low-order end up with g^(tau^0).  ``polynomials_be``: big-endian stream or a tensorcheck.FoldedPolynomialTree."""
        from .tensorcheck import FoldedPolynomialTree

        if isinstance(polynomials_be, FoldedPolynomialTree):
            levels = polynomials_be.levels
        else:
            if len(challenges) == 0:
                return []
            levels = self._le(polynomials_be).fold_chain([c % R for c in challenges])
        n_levels = max(len(levels), 1)
        chunk = max(max_msm_buffer // n_levels, MIN_DEVICE_CHUNK, 1)
        if CONCURRENT_COMMITS and len(levels) >= 2 and all(0 < lvl.n <= chunk for lvl in levels):
            # every level is one resident piece: the independent MSMs share the GPU on several lanes (batch_commit's scheme)
            be = [lvl.reversed() for lvl in levels]
            out = _msm_on_lanes(self, self.srs_be, [(v.ptr, v.n, len(self) - v.n) for v in be])
            for v in be:
                v.free()
            return out
        return [self._commit_le(lvl, max_msm_buffer // n_levels) for lvl in levels]

    def _helper_contexts(self, count: int) -> List[Context]:
        return _helper_contexts_of(self, count)

    def open_folding(self, polynomials, points: Sequence[int], etas: Sequence[int], max_msm_buffer: int = 0):
        """space.rs:229-285 -> (remainders per level, evaluation proof).  ``polynomials``: tensorcheck.FoldedPolynomialTree.

        The reference streams every level through a division by the vanishing polynomial of ``points`` and feeds
        eta_i * quotient coefficients into one HashMapPippenger (all levels pair the coefficient of degree d with the
        same base g^(tau^d), so the map merges them).  Here: k synthetic divisions per level on the device
        (k = len(points)), the eta-combination of the quotients as one resident vector, ONE MSM."""
        ctx = self.ctx
        remainders, batched = [], None
        for i, lvl in enumerate(polynomials.levels):
            q, rem = _divide_by_points(lvl, points)
            remainders.append(rem)
            if q.n:
                if batched is None:
                    batched = DeviceFr.zeros(ctx, max(l.n for l in polynomials.levels))
                batched.axpy(etas[i] % R, q)
        if batched is None:
            return remainders, None
        return remainders, self._commit_le(batched, max_msm_buffer)
