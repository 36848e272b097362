#!/bin/bash
mkdir -p gpurun_out

( time timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_msm_affine.py tests/test_gpu_golden.py tests/test_gpu_stream.py tests/test_gpu_arith.py -m gpu -x -q ) > gpurun_out/pytest_gpu3.log 2>&1
tail -4 gpurun_out/pytest_gpu3.log
for a in 1 2 3 4; do
  export GM_MSM_AFFINE=$a
  timeout 300 python bench.py --steps 8 --no-cpu > gpurun_out/bench3_n20_aff_$a.json 2> gpurun_out/bench3_n20_aff_$a.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench3_n20_aff_$a.json").read().strip().splitlines()[-1])
    print("n20 affine=$a", round(d["ms_per_step"],3), "ms", "%.3e"%d["value"], d["phases_ms"], d["gpu_launches"])
except Exception as e:
    print("n20 affine=$a FAILED", e)
PY
done
for a in 5; do
  export GM_MSM_AFFINE=$a
  timeout 400 python bench.py --steps 4 --no-cpu --logn 24 > gpurun_out/bench3_n24_aff_$a.json 2> gpurun_out/bench3_n24_aff_$a.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench3_n24_aff_$a.json").read().strip().splitlines()[-1])
    print("n24 affine=$a", round(d["ms_per_step"],3), "ms", "%.3e"%d["value"], d["phases_ms"], d["gpu_launches"])
except Exception as e:
    print("n24 affine=$a FAILED", e)
PY
done
export GM_MSM_AFFINE=3
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches3_n20.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu3_n20.log 2>&1
