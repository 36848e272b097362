"""KZG committer keys over the device MSM.

  * :class:`CommitterKey`       - /root/reference/src/kzg/time.rs:24-160 (in-memory, little-endian coefficients)
  * :class:`CommitterKeyStream` - /root/reference/src/kzg/space.rs:59-297 (big-endian streams, chunked MSM)

The MSMs (the hot path) run on the GPU against the device-resident SRS.  The O(n) scalar preparation
around them (Horner quotients, division by the vanishing polynomial, linear combinations) is host-side
bookkeeping in this round - SURVEY.md 8(f) rank 1 lists it as the next tier to move onto the device.
"""
from __future__ import annotations

from collections import deque
from typing import List, Sequence, Tuple

import numpy as np

from . import field
from .context import Context, Srs, as_fr_array
from .msm import ChunkedPippenger, HashMapPippenger, VariableBaseMSM, _DeviceStream, msm_chunks

R = field.R


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


def _poly_div(f: Sequence[int], z: Sequence[int]) -> List[int]:
    """DensePolynomial::div by a monic divisor, quotient only (time.rs:142-143)."""
    f = [x % R for x in f]
    dz = len(z) - 1
    if len(f) <= dz:
        return []
    q = [0] * (len(f) - dz)
    for i in range(len(f) - 1, dz - 1, -1):
        c = f[i]
        q[i - dz] = c
        if c:
            for j, zc in enumerate(z):
                f[i - dz + j] = (f[i - dz + j] - c * zc) % R
    return q


def _linear_combination(polys: Sequence[Sequence[int]], coeffs: Sequence[int]) -> List[int]:
    """misc::linear_combination (src/misc.rs:37-48)."""
    n = max((len(p) for p in polys), default=0)
    out = [0] * n
    for p, c in zip(polys, coeffs):
        for i, v in enumerate(p):
            out[i] = (out[i] + c * v) % R
    return out


def _powers(x: int, n: int) -> List[int]:
    out, cur = [], 1
    for _ in range(n):
        out.append(cur)
        cur = cur * x % R
    return out


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

    def commit(self, polynomial) -> field.Point:
        """time.rs:81-83: one MSM against the SRS prefix (silently truncating to the shorter)."""
        return self._msm.msm_unchecked(self.srs, polynomial)

    def commit_raw(self, polynomial) -> np.ndarray:
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
        """time.rs:98-107."""
        return [self.commit(p) for p in polynomials]

    def open(self, polynomial: Sequence[int], evaluation_point: int) -> Tuple[int, field.Point]:
        """time.rs:112-131: synthetic division (Horner) then an MSM over the quotient."""
        quotient_rev = []
        prev = 0
        for c in reversed(list(polynomial)):
            coeff = (c + prev * evaluation_point) % R
            quotient_rev.append(coeff)
            prev = coeff
        if not quotient_rev:
            return 0, None
        quotient = quotient_rev[::-1]
        return quotient[0], self._msm.msm_unchecked(self.srs, quotient[1:])

    def open_multi_points(self, polynomial: Sequence[int], eval_points: Sequence[int]) -> field.Point:
        """time.rs:134-145."""
        return self.commit(_poly_div(polynomial, vanishing_polynomial(eval_points)))

    def batch_open_multi_points(self, polynomials, eval_points: Sequence[int], eval_chal: int) -> field.Point:
        """time.rs:149-159."""
        etas = _powers(eval_chal, len(polynomials))
        return self.open_multi_points(_linear_combination(polynomials, etas), eval_points)


def _folded_len(n: int, k: int) -> int:
    return (n + (1 << k) - 1) >> k


class CommitterKeyStream:
    """``CommitterKeyStream``: the SRS in BIG-endian stream order (``Reverse(powers_of_g)``, space.rs:288-297)."""

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

    def commit(self, polynomial_be, step: int = 1 << 20) -> field.Point:
        """space.rs:169-177 -> msm_chunks (space.rs:22-55)."""
        return msm_chunks(self.ctx, self.srs_be, polynomial_be, step)

    def open(self, polynomial_be: Sequence[int], alpha: int, max_msm_buffer: int) -> Tuple[int, field.Point]:
        """space.rs:95-125: quotient coefficients stream into the chunked MSM."""
        poly = [x % R for x in polynomial_be]
        n, off = len(poly), len(self) - len(poly)
        st = _DeviceStream(self.ctx, self.srs_be, max(max_msm_buffer, 1))
        cap = max(max_msm_buffer, 1)
        prev, buf, start = 0, [], 0
        for i, scalar in enumerate(poly):
            buf.append(prev)
            prev = (prev * alpha + scalar) % R
            if len(buf) == cap:
                st.push_range(off + start, buf)
                start, buf = i + 1, []
        if buf:
            st.push_range(off + start, buf)
        return prev, st.finalize()

    def open_multi_points(self, polynomial_be: Sequence[int], points: Sequence[int], max_msm_buffer: int):
        """space.rs:128-166: returns (remainder, proof)."""
        zeros = vanishing_polynomial(points)
        deg = len(zeros) - 1
        poly = [x % R for x in polynomial_be]
        off = len(self) - len(poly) + deg
        it = iter(poly)
        state = deque(next(it) for _ in range(len(points)))
        cap = max(max_msm_buffer, 1)
        st = _DeviceStream(self.ctx, self.srs_be, cap)
        buf, start, i = [], 0, 0
        for coeff in it:
            qc = state.popleft()
            state.append(coeff)
            for k in range(len(points)):
                state[k] = (state[k] - zeros[deg - k - 1] * qc) % R
            buf.append(qc)
            i += 1
            if len(buf) == cap:
                st.push_range(off + start, buf)
                start, buf = i, []
        if buf:
            st.push_range(off + start, buf)
        return list(state), st.finalize()

    def commit_folding(self, polynomials_be, challenges: Sequence[int], max_msm_buffer: int) -> List[field.Point]:
        """space.rs:192-223.  The reference walks the FoldedPolynomialTree once and feeds one
        ChunkedPippenger per level; here every level is produced by the device fold chain and committed
        against the SRS range that lines its low-order end up with g^(tau^0)."""
        f_le = as_fr_array(polynomials_be)[::-1].copy()
        k = len(challenges)
        if k == 0:
            return []
        levels = self.ctx.fr_fold_chain(f_le, challenges)
        out = []
        for lvl in levels:
            m = lvl.shape[0]
            st = _DeviceStream(self.ctx, self.srs_be, max(m, 1))
            st.push_range(len(self) - m, np.ascontiguousarray(lvl[::-1]))
            out.append(st.finalize())
        return out

    def open_folding(self, polynomials, points: Sequence[int], etas: Sequence[int], max_msm_buffer: int = 0):
        """space.rs:229-285 -> (remainders per level, evaluation proof).  ``polynomials``: tensorcheck.FoldedPolynomialTree.

        The reference streams every level through a division by the vanishing polynomial of ``points`` and feeds
        eta_i * quotient coefficients into one HashMapPippenger (all levels pair the coefficient of degree d with the
        same base g^(tau^d), so the map merges them).  Here: k synthetic divisions per level on the device
        (k = len(points)), the eta-combination of the quotients as one resident vector, ONE MSM."""
        from .devvec import DeviceFr

        ctx = self.ctx
        pts = [p % R for p in points]
        k = len(pts)
        remainders, batched = [], None
        for i, lvl in enumerate(polynomials.levels):
            q, cs = lvl, []
            for a in pts:
                if q.n == 0:
                    cs.append(0)
                    continue
                q, c = q.div_linear(a)
                cs.append(c)
            # remainder in Newton form c_1 + c_2 (X - a_1) + c_3 (X - a_1)(X - a_2) ... -> monomial coefficients
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
            remainders.append(rem[::-1])                      # deque order of the reference: highest degree first
            if q.n:
                if batched is None:
                    batched = DeviceFr.zeros(ctx, max(l.n for l in polynomials.levels))
                batched.axpy(etas[i] % R, q)
        if batched is None:
            return remainders, None
        m = batched.n
        st = _DeviceStream(ctx, self.srs_be, m)
        st.push_range(len(self) - m, np.ascontiguousarray(batched.limbs()[::-1]))
        return remainders, st.finalize()

