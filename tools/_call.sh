#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests/test_gpu_dist.py tests/test_gpu_elastic.py -m gpu -x -q > $O/q15_pytest.log 2>&1; tail -n 5 $O/q15_pytest.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3) > $O/q15_bench_2gpu.json 2> $O/q15_bench_2gpu.err
tail -n 6 $O/q15_bench_2gpu.err; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/q15_bench_2gpu.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'])
for k,v in d['extra'].items(): print(k, json.dumps(v)[:700])
PY
