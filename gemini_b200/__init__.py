"""gemini_b200 - B200-native (sm_100a) implementation of the Gemini prover hot path.

The package is a thin host-side mirror of the reference's interfaces for that path
(arkworks-rs/gemini: ``ark_ec::VariableBaseMSM``, ``kzg::CommitterKey{,Stream}``,
``subprotocols::sumcheck``, ``tensorcheck::foldings_polynomial``, herring's fold) on top of
the C ABI in ``include/gemini_b200.h`` / ``libgemini_b200.so``.  All arithmetic runs in
hand-written CUDA kernels; importing the package without the compiled library, or creating
a :class:`Context` without a CUDA device, fails loudly - there is no CPU fallback.
"""
from ._lib import GeminiError, lib, lib_path  # noqa: F401
from .context import Context, Srs  # noqa: F401
from . import field  # noqa: F401
from .msm import VariableBaseMSM, ChunkedPippenger, HashMapPippenger, msm_chunks  # noqa: F401
from .kzg import CommitterKey, CommitterKeyStream  # noqa: F401
from .sumcheck import TimeProver, HerringTimeProver, SpaceProver, ElasticProver, Sumcheck, fold_polynomial  # noqa: F401
from .tensorcheck import (foldings_polynomial, FoldedPolynomialTree, evaluate_folding, transcribe_foldings,  # noqa: F401
                          partially_foldtree)
from . import dist  # noqa: F401
from .devvec import DeviceFr, DeviceCsr  # noqa: F401
from .transcript import MerlinTranscript  # noqa: F401
from . import snark  # noqa: F401

__all__ = [
    "Context", "Srs", "GeminiError", "VariableBaseMSM", "ChunkedPippenger", "HashMapPippenger", "msm_chunks",
    "CommitterKey", "CommitterKeyStream", "TimeProver", "HerringTimeProver", "SpaceProver", "ElasticProver",
    "Sumcheck", "fold_polynomial", "foldings_polynomial", "FoldedPolynomialTree", "evaluate_folding", "transcribe_foldings",
    "partially_foldtree", "dist", "field", "DeviceFr", "DeviceCsr", "MerlinTranscript", "snark",
]
