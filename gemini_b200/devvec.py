"""Device-resident Fr vectors: the prover's polynomials stay in HBM between the hot calls."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import field
from ._lib import check, lib
from .context import Context, _ptr, as_fr_array


def _fr1(x: int) -> np.ndarray:
    return field.fr_to_limbs([x])


class DeviceFr:
    """n Fr elements (Montgomery limbs) in device memory; a view does not own its storage."""

    def __init__(self, ctx: Context, n: int, ptr: Optional[int] = None):
        self.ctx = ctx
        self.n = n
        self.owned = ptr is None
        self.ptr = ctx.dev_alloc(max(n, 1) * 32) if ptr is None else ptr

    # -- construction ------------------------------------------------------------------------
    @classmethod
    def from_host(cls, ctx: Context, values) -> "DeviceFr":
        arr = as_fr_array(values)
        v = cls(ctx, arr.shape[0])
        if arr.shape[0]:
            ctx.dev_upload(v.ptr, arr)
        return v

    @classmethod
    def zeros(cls, ctx: Context, n: int) -> "DeviceFr":
        v = cls(ctx, n)
        check(lib.gm_dev_memset(ctx._h, C.c_void_p(v.ptr), 0, n * 32))
        return v

    @classmethod
    def random(cls, ctx: Context, n: int, seed: int) -> "DeviceFr":
        v = cls(ctx, n)
        ctx.fr_random_dev(v.ptr, n, seed)
        return v

    def view(self, offset: int, n: int) -> "DeviceFr":
        assert offset + n <= self.n
        return DeviceFr(self.ctx, n, self.ptr + 32 * offset)

    def reversed(self) -> "DeviceFr":
        """new vector with out[i] = self[n-1-i] (big-endian stream order <-> little-endian), on the device"""
        v = DeviceFr(self.ctx, self.n)
        check(lib.gm_fr_reverse_dev(self.ctx._h, C.c_void_p(self.ptr), self.n, C.c_void_p(v.ptr)))
        return v

    def reverse_(self) -> "DeviceFr":
        check(lib.gm_fr_reverse_dev(self.ctx._h, C.c_void_p(self.ptr), self.n, C.c_void_p(self.ptr)))
        return self

    def clone(self) -> "DeviceFr":
        v = DeviceFr(self.ctx, self.n)
        check(lib.gm_dev_copy(self.ctx._h, C.c_void_p(v.ptr), C.c_void_p(self.ptr), self.n * 32))
        return v

    # -- host access -------------------------------------------------------------------------
    def limbs(self) -> np.ndarray:
        if self.n == 0:
            return np.empty((0, 4), dtype=np.uint64)
        return self.ctx.dev_download(self.ptr, self.n * 32).reshape(self.n, 4)

    def to_ints(self) -> List[int]:
        return field.fr_from_limbs(self.limbs())

    def __len__(self) -> int:
        return self.n

    # -- the vector helpers (C ABI gm_fr_*_dev) ------------------------------------------------
    def evaluate_pm(self, x: int):
        """(f(x), f(-x)) in one pass over the vector (misc::evaluate_le)."""
        out = np.empty(8, dtype=np.uint64)
        check(lib.gm_fr_eval_dev(self.ctx._h, C.c_void_p(self.ptr), self.n, _ptr(_fr1(x)), _ptr(out)))
        e, o = field.fr_from_limbs(out)
        return (e + o) % field.R, (e - o) % field.R

    def evaluate(self, x: int) -> int:
        return self.evaluate_pm(x)[0]

    def axpy(self, c: int, x: "DeviceFr", n: Optional[int] = None) -> "DeviceFr":
        """self[i] += c * x[i] for i < n (default len(x))."""
        n = x.n if n is None else n
        assert n <= self.n and n <= x.n
        check(lib.gm_fr_axpy_dev(self.ctx._h, C.c_void_p(self.ptr), C.c_void_p(x.ptr), n, _ptr(_fr1(c))))
        return self

    def hadamard(self, other: "DeviceFr") -> "DeviceFr":
        assert self.n == other.n
        out = DeviceFr(self.ctx, self.n)
        check(lib.gm_fr_hadamard_dev(self.ctx._h, C.c_void_p(self.ptr), C.c_void_p(other.ptr), self.n, C.c_void_p(out.ptr)))
        return out

    def div_linear(self, a: int):
        """(quotient, remainder) of self / (X - a)."""
        q = DeviceFr(self.ctx, max(self.n - 1, 0))
        rem = np.empty(4, dtype=np.uint64)
        check(lib.gm_fr_div_linear_dev(self.ctx._h, C.c_void_p(self.ptr), self.n, _ptr(_fr1(a)), C.c_void_p(q.ptr), _ptr(rem)))
        return q, field.fr_from_limbs(rem)[0]

    def fold_chain(self, challenges: Sequence[int]) -> List["DeviceFr"]:
        """All fold levels (views into one allocation kept alive by the first element)."""
        k = len(challenges)
        tot = int(lib.gm_fr_fold_chain_len(self.n, k))
        store = DeviceFr(self.ctx, tot)
        ch = field.fr_to_limbs(challenges)
        check(lib.gm_fr_fold_chain_dev(self.ctx._h, C.c_void_p(self.ptr), self.n, _ptr(ch), k, C.c_void_p(store.ptr)))
        levels, off, n = [], 0, self.n
        for _ in range(k):
            n = (n + 1) // 2
            lv = store.view(off, n)
            lv._keepalive = store
            levels.append(lv)
            off += n
        return levels

    def free(self) -> None:
        if self.owned and self.ptr:
            self.ctx.dev_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


def powers(ctx: Context, x: int, n: int) -> DeviceFr:
    """misc::powers on the device."""
    out = DeviceFr(ctx, n)
    check(lib.gm_fr_powers_dev(ctx._h, _ptr(_fr1(x)), n, C.c_void_p(out.ptr)))
    return out


def tensor(ctx: Context, elements: Sequence[int]) -> DeviceFr:
    """misc::tensor on the device: 2^k products of subsets of the challenges."""
    assert len(elements) > 0
    out = DeviceFr(ctx, 1 << len(elements))
    rho = field.fr_to_limbs(elements)
    check(lib.gm_fr_tensor_dev(ctx._h, _ptr(rho), len(elements), C.c_void_p(out.ptr)))
    return out


class DeviceCsr:
    """Sparse matrix in CSR form on the device (u32 rowptr / col, Fr values)."""

    def __init__(self, ctx: Context, rowptr: np.ndarray, col: np.ndarray, vals: np.ndarray, ncols: int):
        self.ctx = ctx
        self.nrows = rowptr.shape[0] - 1
        self.ncols = ncols
        self._bufs = []
        for arr in (rowptr.astype(np.uint32), col.astype(np.uint32), np.ascontiguousarray(vals, dtype=np.uint64)):
            p = ctx.dev_alloc(max(arr.nbytes, 16))
            if arr.nbytes:
                ctx.dev_upload(p, np.ascontiguousarray(arr))
            self._bufs.append(p)

    @classmethod
    def from_rows(cls, ctx: Context, rows, ncols: int, transpose: bool = False) -> "DeviceCsr":
        """rows: list of [(value:int, column:int)] as in R1cs (src/circuit.rs:45-52)."""
        entries = [(i, c, v) for i, row in enumerate(rows) for v, c in row]
        nrows = len(rows)
        if transpose:
            entries = [(c, i, v) for i, c, v in entries]
            nrows, ncols = ncols, nrows
        entries.sort(key=lambda t: (t[0], t[1]))
        rowptr = np.zeros(nrows + 1, dtype=np.uint32)
        for r, _, _ in entries:
            rowptr[r + 1] += 1
        rowptr = np.cumsum(rowptr, dtype=np.uint64).astype(np.uint32)
        col = np.array([c for _, c, _ in entries], dtype=np.uint32)
        vals = field.fr_to_limbs([v for _, _, v in entries])
        return cls(ctx, rowptr, col, vals, ncols)

    @classmethod
    def diagonal(cls, ctx: Context, n: int, value: int) -> "DeviceCsr":
        rowptr = np.arange(n + 1, dtype=np.uint32)
        col = np.arange(n, dtype=np.uint32)
        vals = np.ascontiguousarray(np.broadcast_to(field.fr_to_limbs([value]), (n, 4)))
        return cls(ctx, rowptr, col, vals, n)

    def matvec(self, x: DeviceFr) -> DeviceFr:
        assert x.n >= self.ncols
        y = DeviceFr(self.ctx, self.nrows)
        check(lib.gm_fr_spmv_dev(self.ctx._h, C.c_void_p(self._bufs[0]), C.c_void_p(self._bufs[1]), C.c_void_p(self._bufs[2]),
                                 self.nrows, C.c_void_p(x.ptr), C.c_void_p(y.ptr)))
        return y

    def free(self) -> None:
        for p in self._bufs:
            self.ctx.dev_free(p)
        self._bufs = []
