"""Host-side marshalling between Python integers and the in-memory form of the C ABI.

Field elements cross the boundary as little-endian u64 limbs in Montgomery form, exactly as
arkworks' ``Fp<MontBackend, N>`` holds them (ark-ff 0.4.2, /root/reference/Cargo.lock:62-64):
Fr = 4 limbs (R = 2^256), Fq = 6 limbs (R = 2^384).  A G1 affine point is x | y (12 limbs),
the identity is x = y = 0; a projective result is Jacobian X | Y | Z (18 limbs).
This module only converts representations - it performs no group or polynomial arithmetic.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

Q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_M64 = (1 << 64) - 1
_FR_RINV = pow(1 << 256, -1, R)
_FQ_RINV = pow(1 << 384, -1, Q)

Point = Optional[Tuple[int, int]]


def _limbs(x: int, n: int) -> List[int]:
    return [(x >> (64 * i)) & _M64 for i in range(n)]


def _int(limbs) -> int:
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v


def fr_to_limbs(values: Iterable[int], montgomery: bool = True) -> np.ndarray:
    """ints -> (n, 4) uint64 (Montgomery form unless ``montgomery=False`` = BigInt<4>)."""
    vals = list(values)
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v %= R
        if montgomery:
            v = (v << 256) % R
        out[i] = _limbs(v, 4)
    return out


def fr_from_limbs(arr, montgomery: bool = True) -> List[int]:
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    out = []
    for row in a:
        v = _int(row)
        out.append(v * _FR_RINV % R if montgomery else v)
    return out


def fr_one_limbs() -> np.ndarray:
    return fr_to_limbs([1])[0]


def g1_to_limbs(points: Sequence[Point]) -> np.ndarray:
    """[(x, y) | None] -> (n, 12) uint64, identity = all-zero record."""
    out = np.zeros((len(points), 12), dtype=np.uint64)
    for i, p in enumerate(points):
        if p is None:
            continue
        out[i, :6] = _limbs((p[0] << 384) % Q, 6)
        out[i, 6:] = _limbs((p[1] << 384) % Q, 6)
    return out


def g1_from_limbs(arr) -> List[Point]:
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 12)
    out: List[Point] = []
    for row in a:
        x, y = _int(row[:6]), _int(row[6:])
        out.append(None if x == 0 and y == 0 else (x * _FQ_RINV % Q, y * _FQ_RINV % Q))
    return out


def g1_to_ark104(points: Sequence[Point]) -> np.ndarray:
    """arkworks' in-memory ``Affine`` records: x | y | infinity flag, 104 bytes each."""
    packed = g1_to_limbs(points)
    out = np.zeros((len(points), 104), dtype=np.uint8)
    out[:, :96] = packed.view(np.uint8).reshape(len(points), 96)
    for i, p in enumerate(points):
        if p is None:
            out[i, 96] = 1
    return out


def jacobian_to_affine(limbs18) -> Point:
    """Canonicalise a Jacobian X|Y|Z result (Montgomery limbs) to affine ints (SURVEY 8d parity rule)."""
    a = np.asarray(limbs18, dtype=np.uint64).reshape(18)
    x, y, z = (_int(a[0:6]) * _FQ_RINV % Q, _int(a[6:12]) * _FQ_RINV % Q, _int(a[12:18]) * _FQ_RINV % Q)
    if z == 0:
        return None
    zi = pow(z, -1, Q)
    zi2 = zi * zi % Q
    return (x * zi2 % Q, y * zi2 * zi % Q)


def affine_to_jacobian_limbs(p: Point) -> np.ndarray:
    out = np.zeros(18, dtype=np.uint64)
    one = (1 << 384) % Q
    if p is None:
        out[0:6] = _limbs(one, 6)
        out[6:12] = _limbs(one, 6)
        return out
    out[0:6] = _limbs((p[0] << 384) % Q, 6)
    out[6:12] = _limbs((p[1] << 384) % Q, 6)
    out[12:18] = _limbs(one, 6)
    return out
