"""ctypes binding of libgemini_b200.so (the C ABI declared in include/gemini_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GEMINI_B200_LIB selects another in-tree build of the same ABI (experimental kernels built with `make EXP=...`)
lib_path = os.environ.get("GEMINI_B200_LIB") or os.path.join(_HERE, "libgemini_b200.so")


class GeminiError(RuntimeError):
    """A non-zero return code of the C ABI."""

    def __init__(self, code: int, message: str):
        super().__init__(f"gemini_b200 error {code}: {message}")
        self.code = code


GM_OK, GM_ERR_CUDA, GM_ERR_ARG, GM_ERR_LENGTH, GM_ERR_STATE, GM_ERR_OOM = range(6)

if not os.path.exists(lib_path):
    raise ImportError(
        f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C gemini_b200/csrc`).  gemini_b200 has no CPU fallback."
    )

lib = C.CDLL(lib_path)

_vp, _sz, _i, _l, _u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_long, C.c_uint64
_pp = C.POINTER(C.c_void_p)
_psz = C.POINTER(C.c_size_t)
_pi = C.POINTER(C.c_int)

# name -> (restype, argtypes); mirrors include/gemini_b200.h one to one
SIGNATURES = {
    "gm_init": (_i, [_i, _pp]),
    "gm_shutdown": (_i, [_vp]),
    "gm_last_error": (C.c_char_p, []),
    "gm_abi_version": (_i, []),
    "gm_msm_describe_plan": (_i, [_sz, _i, _i, _vp]),
    "gm_launch_count": (_u64, [_vp]),
    "gm_last_device_ms": (C.c_float, [_vp, _i]),
    "gm_device_synchronize": (_i, [_vp]),
    "gm_timer_start": (_i, [_vp]),
    "gm_timer_stop": (_i, [_vp, C.POINTER(C.c_float)]),
    "gm_l2_flush": (_i, [_vp]),
    "gm_srs_load_g1": (_i, [_vp, _vp, _sz, _sz, _l, _pp]),
    "gm_srs_generate_g1": (_i, [_vp, _sz, _u64, _pp]),
    "gm_srs_setup_g1": (_i, [_vp, _vp, _vp, _sz, _pp]),
    "gm_srs_fill_g1": (_i, [_vp, _vp, _sz, _pp]),
    "gm_srs_precompute": (_i, [_vp, _vp, _sz]),
    "gm_srs_precompute_info": (_i, [_vp, _pi, _pi]),
    "gm_srs_len": (_sz, [_vp]),
    "gm_srs_read": (_i, [_vp, _vp, _sz, _sz, _vp]),
    "gm_srs_free": (_i, [_vp]),
    "gm_msm_g1": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _vp]),
    "gm_msm_g1_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _vp]),
    "gm_msm_g1_checked": (_i, [_vp, _vp, _sz, _sz, _vp, _sz, _vp, _psz]),
    "gm_msm_g1_hostbases": (_i, [_vp, _vp, _sz, _l, _vp, _sz, _i, _vp]),
    "gm_msm_stream_new": (_i, [_vp, _vp, _sz, _pp]),
    "gm_msm_stream_push": (_i, [_vp, _vp, _sz, _l, _sz, _vp, _sz, _i]),
    "gm_msm_stream_push_dev": (_i, [_vp, _sz, _vp, _sz, _i]),
    "gm_msm_stream_finalize": (_i, [_vp, _vp]),
    "gm_msm_stream_free": (_i, [_vp]),
    "gm_g1_sum": (_i, [_vp, _vp, _sz, _vp]),
    "gm_comm_unique_id": (_i, [_vp]),
    "gm_comm_init": (_i, [_vp, _vp, _i, _i]),
    "gm_comm_rank": (_i, [_vp]),
    "gm_comm_world": (_i, [_vp]),
    "gm_comm_nccl_version": (_i, []),
    "gm_comm_barrier": (_i, [_vp]),
    "gm_comm_allgather": (_i, [_vp, _vp, _sz, _vp]),
    "gm_msm_g1_sharded": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _vp]),
    "gm_msm_g1_sharded_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _i, _vp]),
    "gm_msm_stream_finalize_sharded": (_i, [_vp, _vp]),
    "gm_msm_g1_strided_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _sz, _i, _i, _vp]),
    "gm_srs_subsample": (_i, [_vp, _vp, _sz, _sz, _sz, _pp]),
    "gm_fr_fold": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "gm_fr_fold_dev": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "gm_fr_fold_chain": (_i, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "gm_fr_fold_chain_len": (_sz, [_sz, _sz]),
    "gm_sumcheck_new": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _i, _pp]),
    "gm_sumcheck_new_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _i, _pp]),
    "gm_sumcheck_new_ex": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _i, _i, _pp]),
    "gm_sumcheck_set_flavour": (_i, [_vp, _i]),
    "gm_sumcheck_next_message": (_i, [_vp, _vp, _vp, _pi]),
    "gm_sumcheck_fold": (_i, [_vp, _vp]),
    "gm_sumcheck_rounds": (_sz, [_vp]),
    "gm_sumcheck_round": (_sz, [_vp]),
    "gm_sumcheck_set_rounds": (_i, [_vp, _sz, _sz]),
    "gm_sumcheck_final_foldings": (_i, [_vp, _vp, _pi]),
    "gm_sumcheck_timer_start": (_i, [_vp]),
    "gm_sumcheck_timer_stop": (_i, [_vp, C.POINTER(C.c_float)]),
    "gm_sumcheck_state_dev": (_i, [_vp, _pp, _psz, _pp, _psz]),
    "gm_sumcheck_read_state": (_i, [_vp, _vp, _psz, _vp, _psz, _vp]),
    "gm_sumcheck_free": (_i, [_vp]),
    "gm_transcript_new": (_i, [C.c_char_p, _sz, _pp]),
    "gm_transcript_clone": (_i, [_vp, _pp]),
    "gm_transcript_free": (_i, [_vp]),
    "gm_transcript_append_message": (_i, [_vp, C.c_char_p, _sz, C.c_char_p, _sz]),
    "gm_transcript_challenge_bytes": (_i, [_vp, C.c_char_p, _sz, _vp, _sz]),
    "gm_transcript_append_fr": (_i, [_vp, C.c_char_p, _sz, _vp, _sz]),
    "gm_transcript_get_challenge_fr": (_i, [_vp, C.c_char_p, _sz, _vp]),
    "gm_sumcheck_prove": (_i, [_vp, _vp, _vp, _vp, _sz, _psz, _vp]),
    "gm_dev_alloc": (_i, [_vp, _sz, _pp]),
    "gm_dev_free": (_i, [_vp, _vp]),
    "gm_dev_upload": (_i, [_vp, _vp, _vp, _sz]),
    "gm_dev_download": (_i, [_vp, _vp, _vp, _sz]),
    "gm_fr_random_dev": (_i, [_vp, _vp, _sz, _u64]),
    "gm_dev_memset": (_i, [_vp, _vp, _i, _sz]),
    "gm_dev_copy": (_i, [_vp, _vp, _vp, _sz]),
    "gm_fr_reverse_dev": (_i, [_vp, _vp, _sz, _vp]),
    "gm_fr_powers_dev": (_i, [_vp, _vp, _sz, _vp]),
    "gm_fr_eval_dev": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "gm_fr_tensor_dev": (_i, [_vp, _vp, _sz, _vp]),
    "gm_fr_hadamard_dev": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "gm_fr_axpy_dev": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "gm_fr_spmv_dev": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "gm_fr_div_linear_dev": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "gm_fr_fold_chain_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "gm_selftest_field": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz]),
    "gm_selftest_curve": (_i, [_vp, _i, _vp, _vp, _vp, _sz]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


def check(code: int) -> None:
    if code != GM_OK:
        raise GeminiError(code, lib.gm_last_error().decode("utf-8", "replace"))
