"""Merlin transcript (STROBE-128 over Keccak-f[1600]) with the GeminiTranscript shorthands.

Host-side Fiat-Shamir of the reference: /root/reference/src/transcript.rs:8-34 on top of the `merlin` 3.0.0 crate
(Cargo.lock:606-608, not vendored): restated from the published Merlin / STROBE specifications and pinned by
Merlin's own known-answer vector in tests/test_transcript.py.  It stays on the host, like in the reference - per
round it hashes 64 bytes.  Encodings follow ark-serialize 0.4 `serialize_uncompressed`:
  Fr            32 bytes, little-endian canonical integer
  RoundMsg      a | b (sumcheck/prover.rs:10)
  G1 (ark-bls12-381): 96 bytes, zcash layout - x | y big-endian, flag bits in the top of byte 0 (infinity = 0x40)
`get_challenge` = 64 PRF bytes -> Fr::from_random_bytes (first 32 bytes little-endian, top bit cleared, retry
when >= r).  The byte-level agreement with a compiled reference cannot be checked in this image (no Rust
toolchain); the layouts above are the ark-serialize / ark-ff 0.4.2 behaviour as documented upstream.
"""
from __future__ import annotations

from . import field

_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
    0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
    0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M = (1 << 64) - 1


def _rol(x: int, n: int) -> int:
    return ((x << n) | (x >> (64 - n))) & _M if n else x


def keccak_f1600(state: bytearray) -> None:
    a = [[int.from_bytes(state[8 * (x + 5 * y): 8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y): 8 * (x + 5 * y) + 8] = a[x][y].to_bytes(8, "little")


_R = 166
_FLAG_I, _FLAG_A, _FLAG_C, _FLAG_T, _FLAG_M, _FLAG_K = 1, 2, 4, 8, 16, 32


class Strobe128:
    def __init__(self, protocol_label: bytes):
        st = bytearray(200)
        st[0:6] = bytes([1, _R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.state, self.pos, self.pos_begin, self.cur_flags = st, 0, 0, 0
        self.meta_ad(protocol_label, False)

    def _run_f(self) -> None:
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[_R + 1] ^= 0x80
        keccak_f1600(self.state)
        self.pos = self.pos_begin = 0

    def _absorb(self, data: bytes) -> None:
        for byte in data:
            self.state[self.pos] ^= byte
            self.pos += 1
            if self.pos == _R:
                self._run_f()

    def _squeeze(self, n: int) -> bytes:
        out = bytearray(n)
        for i in range(n):
            out[i] = self.state[self.pos]
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == _R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags: int, more: bool) -> None:
        if more:
            assert self.cur_flags == flags
            return
        assert flags & _FLAG_T == 0
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if flags & (_FLAG_C | _FLAG_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data: bytes, more: bool) -> None:
        self._begin_op(_FLAG_M | _FLAG_A, more)
        self._absorb(data)

    def ad(self, data: bytes, more: bool) -> None:
        self._begin_op(_FLAG_A, more)
        self._absorb(data)

    def prf(self, n: int, more: bool) -> bytes:
        self._begin_op(_FLAG_I | _FLAG_A | _FLAG_C, more)
        return self._squeeze(n)


class MerlinTranscript:
    """merlin::Transcript + GeminiTranscript (src/transcript.rs)."""

    def __init__(self, label: bytes = b"GEMINI-v0", g1_encoding: str = "zcash"):  # PROTOCOL_NAME, src/lib.rs:74
        self.strobe = Strobe128(b"Merlin v1.0")
        self.g1_encoding = g1_encoding
        self.append_message(b"dom-sep", label)

    def append_message(self, label: bytes, message: bytes) -> None:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n, False)

    # -- GeminiTranscript ----------------------------------------------------------------------
    def serialize(self, obj) -> bytes:
        if isinstance(obj, int):
            return (obj % field.R).to_bytes(32, "little")
        if isinstance(obj, (tuple, list)):
            return b"".join(self.serialize(o) for o in obj)
        raise TypeError(type(obj))

    def _g1(self, p) -> bytes:
        if self.g1_encoding == "zcash":
            if p is None:
                return bytes([0x40]) + bytes(95)
            return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")
        if p is None:  # ark-serialize default SW layout: x | y little-endian, infinity flag in the last byte
            return bytes(95) + bytes([0x40])
        return p[0].to_bytes(48, "little") + p[1].to_bytes(48, "little")

    def append_g1(self, label: bytes, point) -> None:
        self.append_message(label, self._g1(point))

    def append_serializable(self, label: bytes, obj) -> None:
        self.append_message(label, self.serialize(obj))

    def get_challenge(self, label: bytes) -> int:
        while True:
            b = self.challenge_bytes(label, 64)
            v = int.from_bytes(b[:32], "little") & ((1 << 255) - 1)
            if v < field.R:
                return v
