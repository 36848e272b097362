#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu7.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu7.log
run() {  # name logn steps extra
  timeout 300 python bench.py --steps $3 --no-cpu --logn $2 $4 > gpurun_out/bench7_$1.json 2> gpurun_out/bench7_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench7_$1.json").read().strip().splitlines()[-1])
    print("$1", round(d["ms_per_step"],3), "ms", "%.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], {k: round(v,3) for k,v in d["phases_ms"].items()}, d["gpu_launches"])
except Exception as e:
    print("$1 FAILED", e)
PY
}
run n20 20 8
run n24 24 3
run n22 22 4
run n20_equal 20 6 "--scalars equal"
run n20_noprecompute 20 6 --no-precompute
run n24_noprecompute 24 3 --no-precompute
timeout 300 python tools/bench_snark.py > gpurun_out/snark7.json 2>&1; tail -3 gpurun_out/snark7.json
timeout 300 python tools/bench_stream.py > gpurun_out/stream7.json 2>&1; tail -3 gpurun_out/stream7.json
