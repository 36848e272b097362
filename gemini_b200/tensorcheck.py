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
