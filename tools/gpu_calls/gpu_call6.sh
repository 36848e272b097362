#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_msm_affine.py tests/test_gpu_golden.py tests/test_gpu_stream.py tests/test_gpu_snark.py -m gpu -x -q ) > gpurun_out/pytest_gpu6.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu6.log
run() {  # name logn steps
  timeout 300 python bench.py --steps $3 --no-cpu --logn $2 > gpurun_out/bench6_$1.json 2> gpurun_out/bench6_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench6_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["ms_per_step"],3), "ms", "%.3e"%d["value"], {k: round(v,3) for k,v in d["phases_ms"].items()}, d["gpu_launches"])
except Exception as e:
    print("$1 FAILED", e)
PY
}
for pad in 0 1; do for fg in 0 32; do
  export GM_TABLE_PAD=$pad GM_L2_FETCH=$fg
  run n20_pad${pad}_fg${fg} 20 8
  run n24_pad${pad}_fg${fg} 24 3
done; done
export GM_TABLE_PAD=1 GM_L2_FETCH=32
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches6_n24.csv python bench.py --steps 1 --warmup 1 --no-cpu --logn 24 > gpurun_out/ncu6_n24.log 2>&1
export GM_TABLE_PAD=0 GM_L2_FETCH=32
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches6_n24_nopad.csv python bench.py --steps 1 --warmup 1 --no-cpu --logn 24 > gpurun_out/ncu6_n24b.log 2>&1
timeout 200 python tools/bench_sumcheck.py --reps 3 > gpurun_out/sumcheck6.json 2>&1
cat gpurun_out/sumcheck6.json
