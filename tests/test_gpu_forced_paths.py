"""The long-vector code paths at SHORT, ragged sizes.

Two paths of the library only switch on above a size threshold: the staged (cp.async) sumcheck round kernel
(`k_sc_staged`, more than 2^15 pairs) and the constant-scalar-vector MSM (`msm_common`, 2^16 terms and more).  The
thresholds are read once per process from the environment, so this test re-runs the sumcheck, fold and MSM parity suites
in a child process with both thresholds lowered to a few elements: every ragged length pair of those suites (93 / 16,
9000 / 8191, odd tails, empty inputs, all-equal scalars, identity and duplicate bases ...) then goes through the
long-vector kernels and is compared with the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_long_vector_paths_at_small_ragged_sizes():
    env = dict(os.environ, GM_SC_STAGED_MIN_LOG="2", GM_CONST_SCALAR_MIN="2")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_gpu_sumcheck.py"), os.path.join(ROOT, "tests", "test_gpu_msm.py"),
           os.path.join(ROOT, "tests", "test_gpu_golden.py"), os.path.join(ROOT, "tests", "test_gpu_stream.py")]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert " passed" in res.stdout
