"""Handle / context lifetimes of the C ABI (include/gemini_b200.h "lifetime"): the context is reference counted, so
gm_srs_free / gm_sumcheck_free / gm_msm_stream_free are valid AFTER gm_shutdown - the order Rust `Drop` and Python
`__del__` produce.  Round 1's smoke() died of exactly this (use-after-free of the context, exit code 139), so every
order runs in a child process and the exit code is the assertion."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = """
import sys, ctypes as C
sys.path[:0] = [%r, %r]
import numpy as np
import gemini_b200 as gm
from gemini_b200._lib import lib, check
import pyref as o
ctx = gm.Context(0)
pts = [o.g1_mul(o.G1_GEN, k + 2) for k in range(8)]
scal = list(range(3, 11))
want = o.naive_msm(pts, scal)
""" % (ROOT, os.path.join(ROOT, "oracle"))


def _run(body: str):
    code = PRELUDE + textwrap.dedent(body) + "\nprint('child ok')\n"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "child ok" in out.stdout, f"rc={out.returncode}\n{out.stdout[-1500:]}\n{out.stderr[-3000:]}"


@pytest.mark.gpu
def test_handles_freed_after_shutdown():
    _run("""
        ck = gm.CommitterKey(ctx, pts)
        assert ck.commit(scal) == want
        prover = gm.TimeProver(ctx, [1, 2, 3, 4, 5], [5, 4, 3, 2, 1], 7)
        assert prover.next_message(None) is not None
        st = gm.msm._DeviceStream(ctx, ck.srs, 4)
        st.push_range(0, scal[:4])
        ctx.close()                      # gm_shutdown first ...
        ck.srs.free()                    # ... then every kind of handle
        prover.free()
        st.free()
    """)


@pytest.mark.gpu
def test_interpreter_teardown_order():
    # nothing is freed explicitly: __del__ of the handles and of the context run in whatever order the interpreter picks
    _run("""
        ck = gm.CommitterKey(ctx, pts)
        assert ck.commit(scal) == want
        prover = gm.TimeProver(ctx, [1, 2, 3, 4], [4, 3, 2, 1], 1)
        prover.next_message(None)
        vec = gm.DeviceFr.from_host(ctx, [1, 2, 3])
    """)


@pytest.mark.gpu
def test_calls_after_shutdown_fail_cleanly():
    _run("""
        ck = gm.CommitterKey(ctx, pts)
        h = ctx._h
        srs_h = ck.srs._h
        out = np.zeros(18, dtype=np.uint64)
        arr = gm.field.fr_to_limbs(scal)
        check(lib.gm_msm_g1(h, srs_h, 0, arr.ctypes.data, 8, 0, out.ctypes.data))
        lib.gm_shutdown(h)               # the SRS handle keeps the context struct alive
        rc = lib.gm_msm_g1(h, srs_h, 0, arr.ctypes.data, 8, 0, out.ctypes.data)
        assert rc == 4, rc               # GM_ERR_STATE, not a crash
        assert b"shut down" in lib.gm_last_error()
        ctx._h = C.c_void_p(None)        # already shut down by hand
        ck.srs.free()
    """)


@pytest.mark.gpu
def test_smoke_entry_point_exits_cleanly():
    code = "import sys; sys.path.insert(0, %r); import __graft_entry__ as g; g.smoke()" % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "smoke ok" in out.stdout, f"rc={out.returncode}\n{out.stdout[-1500:]}\n{out.stderr[-3000:]}"


def test_null_handles_are_accepted():
    # no device needed: freeing NULL is a no-op for every handle type
    import gemini_b200  # noqa: F401
    from gemini_b200._lib import lib

    assert lib.gm_shutdown(None) == 0
    assert lib.gm_srs_free(None) == 0
    assert lib.gm_sumcheck_free(None) == 0
    assert lib.gm_msm_stream_free(None) == 0
