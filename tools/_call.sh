#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_msm.py tests/test_gpu_msm_affine.py tests/test_gpu_snark.py tests/test_gpu_stream.py tests/test_gpu_dist.py tests/test_gpu_comm.py tests/test_gpu_elastic.py -m gpu -x -q > $O/q6_pytest.log 2>&1; tail -n 4 $O/q6_pytest.log
python tools/bench_snark.py --logn 24 --reps 3 2>&1 | tail -1 > $O/q6_snark.json; python -c "
import json; d=json.load(open('$O/q6_snark.json')); print(d['value'], d['phases_s'])"
for n in 20 24; do python bench.py --logn $n --no-extras --no-cpu --steps 6 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'])"; done
