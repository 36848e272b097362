"""Device-resident stream adaptors of the elastic prover.

The reference's space-efficient prover never holds a polynomial: everything is an ``Iterable`` that can be re-streamed
BIG-endian (highest-degree coefficient first) - ``Reverse(z)``, the column-major matrix streams, ``MatrixTensor``
(/root/reference/src/snark/streams.rs:13-102), ``LinCombStream`` (/root/reference/src/subprotocols/tensorcheck/streams.rs:42-132).
On a B200 a 2^28-element Fr vector is 8 GiB of 180 GB of HBM, so the adaptors here keep the vector they describe
RESIDENT in little-endian order and the "stream" is that vector read backwards:

  * :class:`ReverseStream`  - the big-endian view of a resident little-endian :class:`DeviceFr`; every consumer
    (CommitterKeyStream, SpaceProver / ElasticProver, FoldedPolynomialTree) takes it without touching the host
  * :class:`MatrixTensor`   - per column of the matrix, sum of value * tensor(v)[row]: one expansion of the tensor
    (k_fr_tensor) and one sparse matrix-vector product with the transposed matrix (k_fr_spmv)
  * :class:`LinCombStream`  - sum_i coeff_i * stream_i with the streams aligned at their LOW-degree end: axpy passes

``iter()`` / ``to_ints_be()`` yield the reference's element order for callers (and tests) that want the stream."""
from __future__ import annotations

from typing import List, Sequence

from . import field
from .context import Context
from .devvec import DeviceCsr, DeviceFr, tensor


class ReverseStream:
    """Big-endian stream over a resident little-endian vector (``Reverse(&[F])``, iterable/mod.rs of ark-std)."""

    def __init__(self, le: DeviceFr):
        self.le = le
        self.ctx = le.ctx

    def __len__(self) -> int:
        return self.le.n

    def to_device_be(self) -> DeviceFr:
        """materialise the stream order on the device (one reversed copy, no host traffic)"""
        return self.le.reversed()

    def to_ints_be(self) -> List[int]:
        return self.le.to_ints()[::-1]

    def iter(self):
        return iter(self.to_ints_be())


def as_le_device(ctx: Context, stream_be) -> DeviceFr:
    """resident little-endian vector of a big-endian stream given as ReverseStream / MatrixTensor / LinCombStream
    (no copy), or as a host sequence / limb array (one upload + one on-device reversal)"""
    if isinstance(stream_be, ReverseStream):
        return stream_be.le
    if isinstance(stream_be, (MatrixTensor, LinCombStream)):
        return stream_be.le()
    if isinstance(stream_be, DeviceFr):
        raise TypeError("a bare DeviceFr is little-endian: wrap it in ReverseStream to use it as a big-endian stream")
    v = DeviceFr.from_host(ctx, stream_be)
    return v.reverse_()


class MatrixTensor:
    """``MatrixTensor`` (snark/streams.rs:13-102): for every column (streamed last column first) the sum over its
    entries of value * tensor(v)[row index].  ``matrix_t`` is the TRANSPOSED matrix in CSR form (= the reference's
    column-major stream, circuit.rs:179-205), ``v`` the tensor factors (challenges)."""

    def __init__(self, ctx: Context, matrix_t: DeviceCsr, v: Sequence[int]):
        self.ctx, self.matrix_t, self.v = ctx, matrix_t, [x % field.R for x in v]
        self._le = None

    def __len__(self) -> int:
        return self.matrix_t.nrows          # one element per column of the matrix

    def le(self) -> DeviceFr:
        if self._le is None:
            expanded = tensor(self.ctx, self.v)                  # tensor(v)[index], 2^len(v) elements
            assert expanded.n >= self.matrix_t.ncols, "tensor shorter than the row count of the matrix"
            self._le = self.matrix_t.matvec(expanded)
            expanded.free()
        return self._le

    def to_ints_be(self) -> List[int]:
        return self.le().to_ints()[::-1]

    def iter(self):
        return iter(self.to_ints_be())


class LinCombStream:
    """``LinCombStream`` (tensorcheck/streams.rs:42-132): sum_i coeffs[i] * streams[i]; big-endian streams of unequal
    length are aligned at their low-degree end, i.e. the resident little-endian vectors are added index by index."""

    def __init__(self, ctx: Context, streams: Sequence, coeffs: Sequence[int]):
        self.ctx = ctx
        self.streams = list(streams)
        self.coeffs = [c % field.R for c in coeffs]
        self._le = None

    def __len__(self) -> int:
        return max((len(s) for s in self.streams), default=0)

    def le(self) -> DeviceFr:
        if self._le is None:
            acc = DeviceFr.zeros(self.ctx, len(self))
            for s, c in zip(self.streams, self.coeffs):
                acc.axpy(c, as_le_device(self.ctx, s))
            self._le = acc
        return self._le

    def to_ints_be(self) -> List[int]:
        return self.le().to_ints()[::-1]

    def iter(self):
        return iter(self.to_ints_be())
