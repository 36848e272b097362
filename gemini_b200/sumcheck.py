"""Sumcheck provers on the device: trait ``Prover`` (/root/reference/src/subprotocols/sumcheck/prover.rs:30-45).

  * :class:`TimeProver`        - sumcheck/time_prover.rs:42-137
  * :class:`HerringTimeProver` - herring/time_prover.rs:44-137 over ``FModule`` (herring/module.rs:127-146)
  * :class:`SpaceProver`       - sumcheck/space_prover.rs:38-266.  B200-first: a 2^28-element Fr stream is
    8 GiB and fits in HBM many times over, so the "space" prover keeps the folded vectors resident and
    folds once per round instead of re-streaming and re-folding the whole input every round; it keeps
    the reference's observable behaviour (big-endian inputs, rounds from the MIN length, stream
    alignment, final foldings taken from the head of the stream).
  * :class:`ElasticProver`     - sumcheck/elastic_prover.rs:29-79
  * :class:`Sumcheck`          - the Fiat-Shamir driver, sumcheck/proof.rs:36-66
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import field
from ._lib import check, lib
from .context import Context, _ptr, as_fr_array

GEMINI_TIME, HERRING_F, GEMINI_SPACE = 0, 1, 2
INPUT_DEVICE, INPUT_BIG_ENDIAN = 1, 2


def ark_log2(x: int) -> int:
    return 0 if x <= 1 else (x - 1).bit_length()


def fold_polynomial(ctx: Context, f, r: int) -> List[int]:
    """misc::fold_polynomial (src/misc.rs:52-56) / herring split_fold."""
    return field.fr_from_limbs(ctx.fr_fold(f, r))


def _fr_input(x):
    """(pointer argument, length, is_device) of one prover input: DeviceFr, torch tensor (device or pinned host),
    (n,4) uint64 limb array, or a sequence of ints"""
    if hasattr(x, "ptr") and hasattr(x, "n"):          # gemini_b200.devvec.DeviceFr
        return C.c_void_p(x.ptr), x.n, True, x
    if hasattr(x, "data_ptr"):
        return _ptr(x), x.numel() * x.element_size() // 32, bool(x.is_cuda), x
    arr = as_fr_array(x)
    return _ptr(arr), arr.shape[0], False, arr


class _DeviceProver:
    def __init__(self, ctx: Context, f, g, twist: int, flavour: int, big_endian: bool = False):
        self.ctx = ctx
        pf, nf, f_dev, keep_f = _fr_input(f)
        pg, ng, g_dev, keep_g = _fr_input(g)
        if f_dev != g_dev:
            raise ValueError("f and g must both be host buffers or both be device buffers")
        tw = field.fr_to_limbs([twist])
        h = C.c_void_p()
        flags = (INPUT_DEVICE if f_dev else 0) | (INPUT_BIG_ENDIAN if big_endian else 0)
        check(lib.gm_sumcheck_new_ex(ctx._h, pf, nf, pg, ng, _ptr(tw), flavour, flags, C.byref(h)))
        del keep_f, keep_g       # the prover copied its inputs (Witness::new clones, time_prover.rs:26-32)
        self._h = h

    # -- trait Prover ------------------------------------------------------------------------
    def next_message_raw(self, verifier_message) -> Optional[np.ndarray]:
        out = np.empty(8, dtype=np.uint64)
        has = C.c_int(0)
        ch = None if verifier_message is None else as_fr_array(
            [verifier_message] if isinstance(verifier_message, int) else verifier_message)
        check(lib.gm_sumcheck_next_message(self._h, _ptr(ch), _ptr(out), C.byref(has)))
        return out if has.value else None

    def next_message(self, verifier_message: Optional[int]) -> Optional[Tuple[int, int]]:
        raw = self.next_message_raw(verifier_message)
        if raw is None:
            return None
        a, b = field.fr_from_limbs(raw)
        return (a, b)

    def fold(self, r: int) -> None:
        check(lib.gm_sumcheck_fold(self._h, _ptr(field.fr_to_limbs([r]))))

    def rounds(self) -> int:
        return int(lib.gm_sumcheck_rounds(self._h))

    def round(self) -> int:
        return int(lib.gm_sumcheck_round(self._h))

    @property
    def tot_rounds(self) -> int:
        return self.rounds()

    def final_foldings(self) -> Optional[Tuple[int, int]]:
        out = np.empty(8, dtype=np.uint64)
        has = C.c_int(0)
        check(lib.gm_sumcheck_final_foldings(self._h, _ptr(out), C.byref(has)))
        if not has.value:
            return None
        a, b = field.fr_from_limbs(out)
        return (a, b)

    # -- timing (CUDA events on the handle's own stream) ----------------------------------------
    def timer_start(self) -> None:
        check(lib.gm_sumcheck_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        check(lib.gm_sumcheck_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    # -- inspection --------------------------------------------------------------------------
    def lengths(self) -> Tuple[int, int]:
        nf, ng = C.c_size_t(0), C.c_size_t(0)
        check(lib.gm_sumcheck_read_state(self._h, None, C.byref(nf), None, C.byref(ng), None))
        return int(nf.value), int(ng.value)

    def state(self):
        """(f, g, twist) of the current round as Python ints (tests, elastic hand-off)."""
        nf, ng = self.lengths()
        f = np.empty((nf, 4), dtype=np.uint64)
        g = np.empty((ng, 4), dtype=np.uint64)
        tw = np.empty(4, dtype=np.uint64)
        a, b = C.c_size_t(0), C.c_size_t(0)
        check(lib.gm_sumcheck_read_state(self._h, _ptr(f), C.byref(a), _ptr(g), C.byref(b), _ptr(tw)))
        return field.fr_from_limbs(f), field.fr_from_limbs(g), field.fr_from_limbs(tw)[0]

    def free(self) -> None:
        if self._h:
            lib.gm_sumcheck_free(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class TimeProver(_DeviceProver):
    """sumcheck/time_prover.rs: rounds = ceil(log2(max(|f|,|g|))), twist in fold AND message."""

    def __init__(self, ctx: Context, f, g, twist: int = 1):
        super().__init__(ctx, f, g, twist, GEMINI_TIME)


class HerringTimeProver(_DeviceProver):
    """herring/time_prover.rs with FModule: rounds from the MIN length, twist only in ``fold``."""

    def __init__(self, ctx: Context, f, g, twist: int = 1):
        super().__init__(ctx, f, g, twist, HERRING_F)


class SpaceProver(_DeviceProver):
    """sumcheck/space_prover.rs: inputs are BIG-endian streams (highest-degree coefficient first).

    ``f_be`` / ``g_be``: host sequences or limb arrays in stream order (uploaded once, reversed on the device), or
    :class:`streams.ReverseStream` / ``MatrixTensor`` / ``LinCombStream`` over resident vectors (no host traffic)."""

    def __init__(self, ctx: Context, f_be, g_be, twist: int = 1):
        from .streams import LinCombStream, MatrixTensor, ReverseStream, as_le_device

        stream_types = (ReverseStream, MatrixTensor, LinCombStream)
        if isinstance(f_be, stream_types) or isinstance(g_be, stream_types):
            # resident little-endian vectors behind both streams: handed over as they are
            super().__init__(ctx, as_le_device(ctx, f_be), as_le_device(ctx, g_be), twist, GEMINI_SPACE)
        else:
            super().__init__(ctx, f_be, g_be, twist, GEMINI_SPACE, big_endian=True)

    def _check_alignment(self, fold_first: bool) -> None:
        """The reference aligns the two folded streams (space_prover.rs:142-174) and asserts equal pair
        counts (:173); inputs on which it panics are rejected here instead of answered differently."""
        nf, ng = self.lengths()
        if fold_first:
            nf, ng = (nf + 1) // 2, (ng + 1) // 2
        if nf == ng:
            return
        if nf > ng:
            nf -= nf - ng + (ng % 2)
        else:
            ng -= ng - nf + (nf % 2)
        assert nf >= 1 and ng >= 1, "stream exhausted during alignment (reference panics: unwrap on None)"
        f_pairs = (nf - 2 + nf % 2) // 2
        g_pairs = (ng - 2 + ng % 2) // 2
        assert f_pairs == g_pairs, "assert_eq!(f_pairs, g_pairs) fails in the reference"

    def next_message_raw(self, verifier_message):
        if self.is_space and self.round() < self.rounds():
            self._check_alignment(verifier_message is not None)
        return super().next_message_raw(verifier_message)

    def check_alignment_all_rounds(self) -> None:
        """the alignment asserts of every remaining round, from the lengths alone (before a native prove loop)"""
        if not self.is_space:
            return
        nf, ng = self.lengths()
        for k in range(self.round(), self.rounds()):
            if k > self.round():       # every message after the first folds by the previous challenge first
                nf, ng = (nf + 1) // 2, (ng + 1) // 2
            a, b = nf, ng
            if a != b:
                if a > b:
                    a -= a - b + (b % 2)
                else:
                    b -= b - a + (a % 2)
                assert a >= 1 and b >= 1, "stream exhausted during alignment (reference panics: unwrap on None)"
                assert (a - 2 + a % 2) // 2 == (b - 2 + b % 2) // 2, "assert_eq!(f_pairs, g_pairs) fails in the reference"

    is_space = True

    def final_foldings(self):
        """head of the folded big-endian streams (space_prover.rs:260-266): gm_sumcheck_final_foldings of the SPACE
        flavour reads the LAST coefficient of the resident little-endian vectors - 64 bytes, not the vectors"""
        nf, ng = self.lengths()
        if self.round() != self.rounds() or nf == 0 or ng == 0:
            return None
        return super().final_foldings()

    def into_time(self) -> None:
        """From<&SpaceProver> for TimeProver (space_prover.rs:269-307): the folded vectors are already resident in
        little-endian order; only the semantics of final_foldings change, round counters and twist are kept."""
        check(lib.gm_sumcheck_set_flavour(self._h, GEMINI_TIME))
        self.is_space = False


class ElasticProver:
    """sumcheck/elastic_prover.rs:29-79.  ``next_message`` always delegates to the wrapped prover (so,
    as in the reference, the Space->Time switch only happens through an explicit ``fold`` call)."""

    def __init__(self, ctx: Context, f_be, g_be, twist: int = 1, threshold: int = 22):
        self.p = SpaceProver(ctx, f_be, g_be, twist)
        self.threshold = threshold  # SPACE_TIME_THRESHOLD, src/lib.rs:76

    @property
    def is_space(self) -> bool:
        return self.p.is_space

    def fold(self, r: int) -> None:
        if self.p.is_space and self.p.rounds() - self.p.round() < self.threshold:
            self.p.into_time()
        self.p.fold(r)

    def next_message(self, verifier_message):
        return self.p.next_message(verifier_message)

    def rounds(self) -> int:
        return self.p.rounds()

    def round(self) -> int:
        return self.p.round()

    @property
    def tot_rounds(self) -> int:
        return self.p.rounds()

    def final_foldings(self):
        if self.p.is_space:
            return self.p.final_foldings()
        return _DeviceProver.final_foldings(self.p)

    def free(self) -> None:
        self.p.free()


class Sumcheck:
    """proof.rs:13-66.  The transcript is abstracted as ``challenge_fn(message) -> challenge``."""

    def __init__(self, messages, challenges, rounds, final_foldings):
        self.messages = messages
        self.challenges = challenges
        self.rounds = rounds
        self.final_foldings = final_foldings

    @classmethod
    def prove_transcript(cls, prover, transcript) -> "Sumcheck":
        """Sumcheck::prove (proof.rs:36-66) against a transcript.  With a device prover and the native Merlin transcript
        the whole Fiat-Shamir loop is ONE call into the library (gm_sumcheck_prove): message kernel, 64-byte D2H,
        transcript append, challenge, next launch - without returning to Python between rounds.  Any other
        prover / transcript pair runs the same loop here."""
        from .transcript import MerlinTranscript

        inner = prover.p if isinstance(prover, ElasticProver) else prover
        if isinstance(inner, _DeviceProver) and isinstance(transcript, MerlinTranscript):
            if isinstance(inner, SpaceProver):
                inner.check_alignment_all_rounds()
            cap = inner.rounds() + 1
            msgs = np.empty((cap, 8), dtype=np.uint64)
            chals = np.empty((cap, 4), dtype=np.uint64)
            fin = np.empty(8, dtype=np.uint64)
            k = C.c_size_t(0)
            check(lib.gm_sumcheck_prove(inner._h, transcript._h, _ptr(msgs), _ptr(chals), cap, C.byref(k), _ptr(fin)))
            k = int(k.value)
            flat = field.fr_from_limbs(msgs[:k].reshape(-1, 4))
            messages = [(flat[2 * i], flat[2 * i + 1]) for i in range(k)]
            return cls(messages, field.fr_from_limbs(chals[:k]), inner.rounds(), [tuple(field.fr_from_limbs(fin.reshape(2, 4)))])
        messages, challenges = [], []
        vm = None
        while True:
            msg = prover.next_message(vm)
            if msg is None:
                break
            transcript.append_serializable(b"evaluations", msg)
            vm = transcript.get_challenge(b"challenge")
            messages.append(msg)
            challenges.append(vm)
        ff = prover.final_foldings()
        transcript.append_serializable(b"final-folding", ff[0])
        transcript.append_serializable(b"final-folding", ff[1])
        return cls(messages, challenges, prover.rounds(), [ff])

    @classmethod
    def prove(cls, prover, challenge_fn: Callable[[Tuple[int, int]], int]) -> "Sumcheck":
        messages, challenges = [], []
        vm = None
        while True:
            msg = prover.next_message(vm)
            if msg is None:
                break
            ch = challenge_fn(msg) % field.R
            vm = ch
            messages.append(msg)
            challenges.append(ch)
        return cls(messages, challenges, prover.rounds(), [prover.final_foldings()])

    @classmethod
    def prove_batch(cls, provers, coefficient_fn: Callable[[], int], challenge_fn) -> "Sumcheck":
        """proof.rs:69-122: random linear combination of the messages of several provers.  The transcript is
        abstracted as ``coefficient_fn() -> batch-sumcheck challenge`` and ``challenge_fn(message)``.  The
        reference drives the provers from rayon workers (proof.rs:85); handles here are independent and the
        device kernels of different provers are queued back to back on the context's stream."""
        R = field.R
        rounds = max((p.rounds() for p in provers), default=0) + 1
        coefficients = [coefficient_fn() % R for _ in provers]
        messages, challenges = [], []
        vm = None
        for _ in range(rounds):
            a_tot = b_tot = 0
            for p, c in zip(provers, coefficients):
                msg = p.next_message(vm)
                if msg is None:
                    ff = p.final_foldings()
                    assert ff is not None, "If next_message is None, we expect final foldings to be available"
                    msg = (ff[0] * ff[1] % R, 0)
                a_tot = (a_tot + msg[0] * c) % R
                b_tot = (b_tot + msg[1] * c) % R
            message = (a_tot, b_tot)
            ch = challenge_fn(message) % R
            vm = ch
            messages.append(message)
            challenges.append(ch)
        return cls(messages, challenges, rounds, [p.final_foldings() for p in provers])

    @classmethod
    def new_time(cls, ctx: Context, challenge_fn, f, g, twist: int) -> "Sumcheck":
        return cls.prove(TimeProver(ctx, f, g, twist), challenge_fn)

    @classmethod
    def new_space(cls, ctx: Context, challenge_fn, f_be, g_be, twist: int) -> "Sumcheck":
        return cls.prove(SpaceProver(ctx, f_be, g_be, twist), challenge_fn)

    @classmethod
    def new_elastic(cls, ctx: Context, challenge_fn, f_be, g_be, twist: int) -> "Sumcheck":
        return cls.prove(ElasticProver(ctx, f_be, g_be, twist), challenge_fn)
