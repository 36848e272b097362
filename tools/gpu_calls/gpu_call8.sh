#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 400 $TR bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench8_n20_g2.json 2> gpurun_out/bench8_n20_g2.err; tail -1 gpurun_out/bench8_n20_g2.json | cut -c1-400
timeout 400 $TR tools/dist_sumcheck.py --logn 24 > gpurun_out/dist_sumcheck8_g2.json 2> gpurun_out/dist_sumcheck8_g2.err; tail -1 gpurun_out/dist_sumcheck8_g2.json; tail -3 gpurun_out/dist_sumcheck8_g2.err
timeout 200 python tools/dist_sumcheck.py --logn 24 > gpurun_out/dist_sumcheck8_g1.json 2> gpurun_out/dist_sumcheck8_g1.err; tail -1 gpurun_out/dist_sumcheck8_g1.json; tail -3 gpurun_out/dist_sumcheck8_g1.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench8_ref_g2.json 2> gpurun_out/bench8_ref_g2.err; tail -1 gpurun_out/bench8_ref_g2.json | cut -c1-300
