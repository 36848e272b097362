"""tensorcheck::foldings_polynomial (/root/reference/src/subprotocols/tensorcheck/mod.rs:124-133)."""
from __future__ import annotations

from typing import List, Sequence

from . import field
from .context import Context


def foldings_polynomial(ctx: Context, polynomial, challenges: Sequence[int], raw: bool = False) -> List:
    """All fold-in-half polynomials f^(1) .. f^(k-1) (the LAST challenge is stripped, mod.rs:128),
    every level kept - one device pass per level, one D2H for the whole chain."""
    chals = list(challenges)[:-1] if len(challenges) else []
    if not chals:
        return []
    levels = ctx.fr_fold_chain(polynomial, chals)
    return levels if raw else [field.fr_from_limbs(l) for l in levels]


# ---------------------------------------------------------------------------------------------------
# Elastic-prover side: FoldedPolynomialTree and what consumes it
# ---------------------------------------------------------------------------------------------------
SPACE_TIME_THRESHOLD = 22  # /root/reference/src/lib.rs


class FoldedPolynomialTree:
    """``FoldedPolynomialTree`` (/root/reference/src/subprotocols/sumcheck/streams.rs:13-138).

    The reference re-streams the BIG-endian input and folds it through a stack machine every time the tree is
    walked.  B200-first: the fold levels f^(1) .. f^(k) are produced once by the device fold chain and stay
    resident in HBM (little-endian ``DeviceFr``), which is what every consumer (commit_folding, evaluate_folding,
    open_folding, transcribe_foldings) actually needs.  ``iter()`` still yields the reference's (level, coefficient)
    sequence for callers that want the stream."""

    def __init__(self, ctx: Context, coefficients_be, challenges: Sequence[int], _levels=None, _n=None):
        from .context import as_fr_array
        from .devvec import DeviceFr

        self.ctx = ctx
        self.challenges = [c % field.R for c in challenges]
        if _levels is not None:
            self.levels, self.n = _levels, _n
            return
        from .streams import as_le_device

        # big-endian stream: host coefficients (one upload, reversed on the device) or a ReverseStream / MatrixTensor /
        # LinCombStream over resident vectors (no host traffic)
        base = as_le_device(ctx, coefficients_be)
        self.n = base.n
        self.levels = base.fold_chain(self.challenges) if self.challenges else []

    @classmethod
    def from_le_device(cls, ctx: Context, poly_le, challenges: Sequence[int]) -> "FoldedPolynomialTree":
        """device-resident little-endian coefficients (the time prover's vectors) - no host copy at all"""
        ch = [c % field.R for c in challenges]
        return cls(ctx, None, ch, _levels=poly_le.fold_chain(ch) if ch else [], _n=poly_le.n)

    def depth(self) -> int:
        return len(self.challenges)

    def __len__(self) -> int:
        return self.n

    def level_le(self, i: int):
        """f^(i), 1 <= i <= depth, little-endian DeviceFr of ceil(n / 2^i) coefficients"""
        return self.levels[i - 1]

    def iter(self):
        """(level, coefficient) for level >= 1 in the order of FoldedPolynomialTreeIter (streams.rs:112-138): a level-j
        coefficient appears as soon as the two level-(j-1) coefficients below it have been produced."""
        lv = [l.to_ints()[::-1] for l in self.levels]          # big-endian per level
        pos = [0] * len(lv)
        depth = self.depth()
        # the stack machine emits, after base element t (counted from the padded start), every level j with 2^j | t+1
        chunk = 1 << depth
        pad = (chunk - self.n % chunk) % chunk
        for t in range(pad, pad + self.n):
            j = 1
            while j <= depth and (t + 1) % (1 << j) == 0:
                yield (j, lv[j - 1][pos[j - 1]])
                pos[j - 1] += 1
                j += 1


def evaluate_folding(polynomials: FoldedPolynomialTree, x: int) -> List[int]:
    """tensorcheck/mod.rs:73-88: f^(j)(x) for j = 1 .. depth, one device Horner pass per level."""
    return [lvl.evaluate(x) for lvl in polynomials.levels]


def transcribe_foldings(foldings: FoldedPolynomialTree, threshold_level: int) -> List[List[int]]:
    """tensorcheck/mod.rs:136-158: the foldings above ``threshold_level`` as little-endian vectors."""
    return [foldings.levels[j].to_ints() for j in range(threshold_level, foldings.depth())]


def partially_foldtree(ctx: Context, stream_be, challenges: Sequence[int]):
    """tensorcheck/mod.rs:160-178 -> (tree over the first ``threshold_level`` challenges, transcribed upper foldings)."""
    full = FoldedPolynomialTree(ctx, stream_be, challenges)
    depth = full.depth()
    threshold = depth - SPACE_TIME_THRESHOLD if depth > SPACE_TIME_THRESHOLD else depth
    transcribed = transcribe_foldings(full, threshold)
    partial = FoldedPolynomialTree(ctx, None, list(challenges)[:threshold], _levels=full.levels[:threshold], _n=full.n)
    return partial, transcribed
