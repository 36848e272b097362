#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu4.log 2>&1
tail -4 gpurun_out/pytest_gpu4.log
run() {  # name logn steps
  timeout 300 python bench.py --steps $3 --no-cpu --logn $2 > gpurun_out/bench4_$1.json 2> gpurun_out/bench4_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench4_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["ms_per_step"],3), "ms", "%.3e"%d["value"], {k: round(v,3) for k,v in d["phases_ms"].items()}, d["gpu_launches"])
except Exception as e:
    print("$1 FAILED", e)
PY
}
export GM_MSM_AFFINE=2
for wps in 64 32 16; do for hv in 1 2; do
  export GM_AFF_WPS=$wps GM_AFF_HALVES=$hv
  run n20_L2_wps${wps}_h${hv} 20 8
done; done
export GM_MSM_AFFINE=3 GM_AFF_WPS=32 GM_AFF_HALVES=2
run n20_L3_wps32_h2 20 8
export GM_MSM_AFFINE=4
for wps in 64 32; do for hv in 1 2; do
  export GM_AFF_WPS=$wps GM_AFF_HALVES=$hv
  run n24_L4_wps${wps}_h${hv} 24 4
done; done
unset GM_MSM_AFFINE GM_AFF_WPS GM_AFF_HALVES
timeout 200 python tools/bench_sumcheck.py --reps 3 > gpurun_out/sumcheck4.json 2>&1
cat gpurun_out/sumcheck4.json
