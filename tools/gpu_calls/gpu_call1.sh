#!/bin/bash
# One GPU call: parity tests, MSM bench with 0..4 affine levels, launch lists (MSM, sumcheck).
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for a in auto 0 1 2 3 4; do
  if [ "$a" = auto ]; then unset GM_MSM_AFFINE; else export GM_MSM_AFFINE=$a; fi
  timeout 300 python bench.py --steps 8 --no-cpu > gpurun_out/bench_n20_aff_$a.json 2> gpurun_out/bench_n20_aff_$a.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n20_aff_$a.json").read().strip().splitlines()[-1])
    print("n20 affine=$a", round(d["ms_per_step"],3), "ms", "%.3e"%d["value"], d["phases_ms"], d["gpu_launches"])
except Exception as e:
    print("n20 affine=$a FAILED", e)
PY
done
for a in auto 0 3 5; do
  if [ "$a" = auto ]; then unset GM_MSM_AFFINE; else export GM_MSM_AFFINE=$a; fi
  timeout 400 python bench.py --steps 4 --no-cpu --logn 24 > gpurun_out/bench_n24_aff_$a.json 2> gpurun_out/bench_n24_aff_$a.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n24_aff_$a.json").read().strip().splitlines()[-1])
    print("n24 affine=$a", round(d["ms_per_step"],3), "ms", "%.3e"%d["value"], d["phases_ms"], d["gpu_launches"])
except Exception as e:
    print("n24 affine=$a FAILED", e)
PY
done
unset GM_MSM_AFFINE
timeout 300 python bench.py --steps 4 --no-cpu --no-precompute > gpurun_out/bench_n20_noprecompute.json 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_n20.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_n20.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_sc_|k_fr_fold' -c 120 --csv --log-file gpurun_out/launches_sumcheck.csv python tools/bench_sumcheck.py --reps 2 > gpurun_out/ncu_sumcheck.log 2>&1
timeout 200 python tools/bench_sumcheck.py --reps 3 > gpurun_out/sumcheck.json 2>&1
cat gpurun_out/sumcheck.json
