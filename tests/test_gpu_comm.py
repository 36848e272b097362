"""The library's own multi-GPU exchange (gemini_b200/csrc/comm.cu): gm_comm_init + gm_msm_g1_sharded.

  * world = 1: the communicator is created, the sharded entry points go through the same ncclAllGather path and must
    return what the plain MSM returns (runs on the single-GPU box of the driver);
  * world = 2: two processes, one GPU each (skipped when the box has one GPU): the MSM split by contiguous point range
    equals the naive sum over ALL points, and the streamed variant agrees."""
import os
import subprocess
import sys

import numpy as np
import pytest

import pyref as o
from util import rand_points, rand_scalars

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_world_of_one():
    code = f"""
import sys
sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'oracle')!r}, {os.path.join(ROOT, 'tests')!r}]
import numpy as np
import gemini_b200 as gm
import pyref as o
from util import rand_points, rand_scalars
ctx = gm.Context(0)
ctx.comm_init(ctx.comm_unique_id(), 0, 1)
assert (ctx.comm_rank, ctx.comm_world) == (0, 1)
ctx.comm_barrier()
row = np.arange(8, dtype=np.uint64)
assert np.array_equal(ctx.comm_allgather(row), row.reshape(1, 8))
pts, sc = rand_points(200, 1), rand_scalars(200, 2)
srs = ctx.srs_load(pts)
a = ctx.msm_sharded(srs, sc)
b = ctx.msm(srs, sc)
assert np.array_equal(a, b) and gm.field.jacobian_to_affine(a) == o.naive_msm(pts, sc)
st = gm.msm._DeviceStream(ctx, srs, 64)
for s0 in range(0, 200, 64):
    st.push_range(s0, sc[s0:s0 + 64])
assert np.array_equal(st.finalize_sharded_raw(), b)
print('child ok', gm.lib.gm_comm_nccl_version())
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "child ok" in out.stdout, f"rc={out.returncode}\n{out.stdout[-1500:]}\n{out.stderr[-3000:]}"


def _gpu_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_two_ranks_sharded_msm():
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = os.path.join(ROOT, "tools", "comm_check.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29741", script], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "comm_check ok" in out.stdout, f"rc={out.returncode}\n{out.stdout[-1500:]}\n{out.stderr[-3000:]}"
