#!/bin/bash
# 8-GPU evidence for BASELINE configs 4-5: weak-scaling MSM at 2^20 and 2^24 per GPU, streamed MSM 2^28 total
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench10_n20_g$N.json 2> gpurun_out/bench10_n20_g$N.err; tail -1 gpurun_out/bench10_n20_g$N.json | cut -c1-260
timeout 400 $TR bench.py --gpus $N --steps 3 --warmup 3 --logn 24 > gpurun_out/bench10_n24_g$N.json 2> gpurun_out/bench10_n24_g$N.err; tail -1 gpurun_out/bench10_n24_g$N.json | cut -c1-260
timeout 500 $TR tools/bench_stream.py --logn 25 --reps 2 > gpurun_out/stream10_n28_g$N.json 2> gpurun_out/stream10_n28_g$N.err; tail -1 gpurun_out/stream10_n28_g$N.json | cut -c1-400; tail -2 gpurun_out/stream10_n28_g$N.err | cut -c1-300
timeout 300 $TR tools/dist_sumcheck.py --logn 27 --reps 2 > gpurun_out/dist_sumcheck10_g$N.json 2> gpurun_out/dist_sumcheck10_g$N.err; tail -1 gpurun_out/dist_sumcheck10_g$N.json | cut -c1-300
