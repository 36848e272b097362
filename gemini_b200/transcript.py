"""Merlin transcript (STROBE-128 over Keccak-f[1600]) with the GeminiTranscript shorthands.

Host-side Fiat-Shamir of the reference: /root/reference/src/transcript.rs:8-34 on top of the `merlin` 3.0.0 crate
(Cargo.lock:606-608, not vendored).  The transcript itself is native code inside libgemini_b200
(gemini_b200/csrc/transcript.cu: ``gm_transcript_*``), restated from the published Merlin / STROBE specifications and
pinned by Merlin's own known-answer vector and by the pure-Python restatement kept with the test oracle
(tests/test_transcript.py); this class only serialises.  It stays on the host, like in the reference - per round it
hashes 64 bytes - and ``Sumcheck::prove`` runs as ONE native call (``gm_sumcheck_prove``), so the round loop never
returns to Python.  Encodings follow ark-serialize 0.4 `serialize_uncompressed`:
  Fr            32 bytes, little-endian canonical integer
  RoundMsg      a | b (sumcheck/prover.rs:10)
  G1 (ark-bls12-381): 96 bytes, zcash layout - x | y big-endian, flag bits in the top of byte 0 (infinity = 0x40)
`get_challenge` = 64 PRF bytes -> Fr::from_random_bytes (first 32 bytes little-endian, top bit cleared, retry
when >= r).  The byte-level agreement with a compiled reference cannot be checked in this image (no Rust
toolchain); the layouts above are the ark-serialize / ark-ff 0.4.2 behaviour as documented upstream.
"""
from __future__ import annotations

import ctypes as C

from . import field
from ._lib import check, lib


class MerlinTranscript:
    """merlin::Transcript + GeminiTranscript (src/transcript.rs) over the native ``gm_transcript``."""

    def __init__(self, label: bytes = b"GEMINI-v0", g1_encoding: str = "zcash"):  # PROTOCOL_NAME, src/lib.rs:74
        h = C.c_void_p()
        check(lib.gm_transcript_new(label, len(label), C.byref(h)))
        self._h = h
        self.g1_encoding = g1_encoding

    def clone(self) -> "MerlinTranscript":
        t = MerlinTranscript.__new__(MerlinTranscript)
        h = C.c_void_p()
        check(lib.gm_transcript_clone(self._h, C.byref(h)))
        t._h, t.g1_encoding = h, self.g1_encoding
        return t

    def __del__(self):
        try:
            if self._h:
                lib.gm_transcript_free(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass

    def append_message(self, label: bytes, message: bytes) -> None:
        check(lib.gm_transcript_append_message(self._h, label, len(label), message, len(message)))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        out = C.create_string_buffer(n)
        check(lib.gm_transcript_challenge_bytes(self._h, label, len(label), out, n))
        return out.raw

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
        out = (C.c_uint64 * 4)()
        check(lib.gm_transcript_get_challenge_fr(self._h, label, len(label), out))
        return field.fr_from_limbs(list(out))[0]
