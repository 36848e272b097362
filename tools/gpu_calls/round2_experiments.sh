#!/bin/bash
# Round-2 A/B of the two changes that were validated on the host only in round 1.  Build the experimental library
# HERE first (no GPU needed):
#   make -C gemini_b200/csrc EXP="-DGM_FAST_INV -DGM_LAZY_SUMCHECK" OUT=../libgemini_b200_exp.so OBJ=../../build/obj_exp
# then:  gpurun --timeout 1500 -- 'bash tools/gpu_calls/round2_experiments.sh'
mkdir -p gpurun_out
export GEMINI_B200_LIB=$PWD/gemini_b200/libgemini_b200_exp.so
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_exp.log 2>&1; grep -E "passed|failed" gpurun_out/r2_pytest_exp.log
for a in 2 3 4; do
  GM_MSM_AFFINE=$a timeout 300 python bench.py --steps 8 --no-cpu > gpurun_out/r2_bench_n20_exp_aff$a.json 2>&1
done
timeout 300 python bench.py --steps 4 --no-cpu --logn 24 > gpurun_out/r2_bench_n24_exp.json 2>&1
timeout 200 python tools/bench_sumcheck.py --reps 3 > gpurun_out/r2_sumcheck_exp.json 2>&1
unset GEMINI_B200_LIB
timeout 300 python bench.py --steps 8 --no-cpu > gpurun_out/r2_bench_n20_base.json 2>&1
timeout 200 python tools/bench_sumcheck.py --reps 3 > gpurun_out/r2_sumcheck_base.json 2>&1
tail -n 3 gpurun_out/r2_sumcheck_exp.json gpurun_out/r2_sumcheck_base.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"], 3), "%.3e" % d["value"], d["phases_ms"])
    except Exception as e:
        print(f, "FAILED", e)
PY
