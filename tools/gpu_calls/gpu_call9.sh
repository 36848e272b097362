#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu9.log 2>&1
grep -E "passed|failed|Error|error" gpurun_out/pytest_gpu9.log | head -20
python - <<'PY' > gpurun_out/srs_setup9.txt 2>&1
import sys, time, random
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import gemini_b200 as gm
ctx = gm.Context(0)
for logn in (16, 20, 24):
    t0 = time.perf_counter()
    ck = gm.CommitterKey.new(ctx, (1 << logn) - 1, 3, random.Random(logn))
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print(f"CommitterKey::new logn={logn}: {dt*1e3:.1f} ms wall ({(1<<logn)/dt:.3e} points/s)")
    ck.srs.free()
ctx.close()
PY
cat gpurun_out/srs_setup9.txt
