#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_msm_affine.py tests/test_gpu_golden.py tests/test_gpu_stream.py -m gpu -x -q ) > gpurun_out/pytest_gpu5.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu5.log
timeout 300 python bench.py --steps 8 > gpurun_out/bench5_n20.json 2> gpurun_out/bench5_n20.err; cat gpurun_out/bench5_n20.json
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches5_n20.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu5_n20.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches5_n24.csv python bench.py --steps 1 --warmup 1 --no-cpu --logn 24 > gpurun_out/ncu5_n24.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_aff_finish|k_aff_prepare|k_accumulate' -s 6 -c 5 -o gpurun_out/r01_ncu_affine_n20 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu5_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
